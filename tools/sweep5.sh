#!/bin/bash
cd "$(dirname "$0")/.."
CS=gym-mapf_b200/csrc
lib=$CS/libmapf_b200_lb4p.so
for B in 2048 65536 262144 524288 1048576 2097152; do
    env TIME_GRAPH=1 TIME_B=$B MAPF_THREADS=256 MAPF_B200_LIB=$lib MAPF_STEP_EPT=2 timeout 120 python tools/time_step.py lb4p 2>&1 | tail -1
done
