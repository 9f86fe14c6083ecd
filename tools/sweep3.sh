#!/bin/bash
cd "$(dirname "$0")/.."
CS=gym-mapf_b200/csrc
for tag in wide lb4 lb5 lb6 lb8; do
  lib=$CS/libmapf_b200_$tag.so
  for ept in 1 2; do
    env TIME_GRAPH=1 TIME_B=1048576 MAPF_THREADS=256 MAPF_B200_LIB=$lib MAPF_STEP_EPT=$ept timeout 120 python tools/time_step.py $tag 2>&1 | tail -1
    env TIME_GRAPH=0 TIME_B=8388608 MAPF_THREADS=256 MAPF_B200_LIB=$lib MAPF_STEP_EPT=$ept timeout 120 python tools/time_step.py $tag 2>&1 | tail -1
  done
done
env TIME_GRAPH=1 TIME_B=1048576 MAPF_THREADS=128 MAPF_B200_LIB=$CS/libmapf_b200_lb5.so MAPF_STEP_EPT=2 timeout 120 python tools/time_step.py lb5 2>&1 | tail -1
env TIME_GRAPH=1 TIME_B=1048576 MAPF_THREADS=128 MAPF_B200_LIB=$CS/libmapf_b200_lb6.so MAPF_STEP_EPT=1 timeout 120 python tools/time_step.py lb6 2>&1 | tail -1
