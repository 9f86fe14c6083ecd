#!/bin/bash
# tuning sweep of the step kernel (run on the GPU box): variants x CTA size x CTAs/SM x envs-per-thread
cd "$(dirname "$0")/.."
CS=gym-mapf_b200/csrc
for lib in $CS/libmapf_b200_*.so; do
  tag=$(basename $lib .so | sed 's/libmapf_b200_//')
  for ept in 1 2; do
    for thr in 128 256 512; do
      for bps in 0 2 4; do
        env MAPF_B200_LIB=$lib MAPF_STEP_EPT=$ept MAPF_THREADS=$thr $( [ $bps -gt 0 ] && echo MAPF_BLOCKS_PER_SM=$bps ) \
          timeout 120 python tools/time_step.py $tag 2>&1 | tail -1
      done
    done
  done
done
