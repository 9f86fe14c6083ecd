cd $GRAFT_REPO_ROOT
CS=gym_mapf_b200/csrc
for B in 1048576 8388608; do
 for t in tune nowait r7; do TIME_GRAPH=1 TIME_B=$B MAPF_B200_LIB=$CS/libmapf_b200_$t.so timeout 120 python tools/time_step.py $t 2>&1 | tail -1; done
 for bps in 1 2; do TIME_GRAPH=1 TIME_B=$B MAPF_BLOCKS_PER_SM=$bps MAPF_B200_LIB=$CS/libmapf_b200_tune.so timeout 120 python tools/time_step.py bps$bps 2>&1 | tail -1; done
 TIME_GRAPH=1 TIME_B=$B MAPF_BLOCKS_PER_SM=1 MAPF_B200_LIB=$CS/libmapf_b200_nowait.so timeout 120 python tools/time_step.py nowait_bps1 2>&1 | tail -1
done
MAPF_B200_LIB=$CS/libmapf_b200_trace.so timeout 120 python tools/trace_step.py 12 > gpurun_out/r02_step_timeline.txt 2>gpurun_out/trace.err; tail -3 gpurun_out/trace.err; head -40 gpurun_out/r02_step_timeline.txt
