#!/bin/bash
# ablations of k_step<4>: which resource bounds it? (1: no move-table gather, 3: no stores, 4: no loads, ab13: no gather + no Philox)
cd "$(dirname "$0")/.."
for t in ab1 ab3 ab4 ab13; do for B in 1048576 8388608; do
  TIME_GRAPH=1 TIME_B=$B MAPF_B200_LIB=gym_mapf_b200/csrc/libmapf_b200_$t.so timeout 120 python tools/time_step.py $t 2>&1 | tail -1
done; done
for B in 1048576 8388608; do TIME_GRAPH=1 TIME_B=$B MAPF_STEP_EPT=1 timeout 120 python tools/time_step.py ept1 2>&1 | tail -1; done
