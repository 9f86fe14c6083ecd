#!/bin/bash
cd "$(dirname "$0")/.."
CS=gym-mapf_b200/csrc
lib=$CS/libmapf_b200_wide.so
for B in 65536 262144 524288 1048576 2097152 4194304 8388608; do
  for ept in 1 2; do
    env TIME_B=$B MAPF_B200_LIB=$lib MAPF_STEP_EPT=$ept timeout 120 python tools/time_step.py wide 2>&1 | tail -1
  done
done
