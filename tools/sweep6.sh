#!/bin/bash
cd "$(dirname "$0")/.."
CS=gym-mapf_b200/csrc
lib=$CS/libmapf_b200_lb4p.so
for pdl in 0 1; do for g in 0 1; do for B in 2048 1048576; do
    env MAPF_PDL=$pdl TIME_GRAPH=$g TIME_B=$B MAPF_THREADS=256 MAPF_B200_LIB=$lib MAPF_STEP_EPT=2 timeout 120 python tools/time_step.py pdl$pdl 2>&1 | tail -1
done; done; done
