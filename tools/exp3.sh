cd $GRAFT_REPO_ROOT
CS=gym_mapf_b200/csrc
for B in 1048576 8388608; do
 for t in tune v4 v5 v5r7; do TIME_GRAPH=1 TIME_B=$B MAPF_B200_LIB=$CS/libmapf_b200_$t.so timeout 120 python tools/time_step.py $t 2>&1 | tail -1; done
done
