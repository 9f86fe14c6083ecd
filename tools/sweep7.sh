#!/bin/bash
# CTA size vs per-launch staging traffic (one 33 KB image per CTA): 256 x4, 512 x2, 1024 x1 per SM
cd "$(dirname "$0")/.."
CS=gym_mapf_b200/csrc
for B in 1048576 8388608; do
for thr in 256 512; do
    env TIME_GRAPH=1 TIME_B=$B MAPF_THREADS=$thr timeout 120 python tools/time_step.py base 2>&1 | tail -1
done
for thr in 256 512 1024; do
    env TIME_GRAPH=1 TIME_B=$B MAPF_THREADS=$thr MAPF_B200_LIB=$CS/libmapf_b200_t1024.so timeout 120 python tools/time_step.py t1024 2>&1 | tail -1
done
done
