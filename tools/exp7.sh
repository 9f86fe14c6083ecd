cd $GRAFT_REPO_ROOT
CS=gym_mapf_b200/csrc
for B in 1048576 8388608; do
 TIME_GRAPH=1 TIME_B=$B timeout 120 python tools/time_step.py ship 2>&1 | tail -1
 TIME_GRAPH=1 TIME_B=$B MAPF_B200_LIB=$CS/libmapf_b200_mb1.so timeout 120 python tools/time_step.py mb1 2>&1 | tail -1
done
