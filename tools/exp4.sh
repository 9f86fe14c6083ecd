cd $GRAFT_REPO_ROOT
CS=gym_mapf_b200/csrc
B=1048576
for t in v6 v6a v6b v6c v6ab; do TIME_GRAPH=1 TIME_B=$B MAPF_B200_LIB=$CS/libmapf_b200_$t.so timeout 120 python tools/time_step.py $t 2>&1 | tail -1; done
for ns in 2 4; do for bps in 1 2; do
TIME_STREAMS=$ns TIME_GRAPH=1 TIME_B=$B MAPF_BLOCKS_PER_SM=$bps MAPF_B200_LIB=$CS/libmapf_b200_v6.so timeout 120 python tools/time_step.py streams${ns}_bps$bps 2>&1 | tail -1
done; done
TIME_STREAMS=2 TIME_GRAPH=1 TIME_B=$B MAPF_B200_LIB=$CS/libmapf_b200_v6.so timeout 120 python tools/time_step.py streams2_full 2>&1 | tail -1
B=8388608
for t in v6 v6a v6b v6c; do TIME_GRAPH=1 TIME_B=$B MAPF_B200_LIB=$CS/libmapf_b200_$t.so timeout 120 python tools/time_step.py $t 2>&1 | tail -1; done
