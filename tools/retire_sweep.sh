#!/bin/bash
# frames in flight x retire policy (RT_B200_KEEP_DIV) on one GPU; world=8 renders rank 0's shard of an 8-GPU frame
mkdir -p gpurun_out
OUT=gpurun_out/retire_sweep.txt
: > $OUT
for kd in 1 4 8; do
  echo "=== KEEP_DIV=$kd world=8" | tee -a $OUT
  RT_B200_KEEP_DIV=$kd RT_PIPE_SHARE="1 2" RT_PIPE_M="8 12 16" RT_PIPE_STEPS=96 timeout 300 python tools/pipe_probe.py c3 8 2>&1 | tail -6 | tee -a $OUT
done
for kd in 1 4; do
  echo "=== KEEP_DIV=$kd world=1" | tee -a $OUT
  RT_B200_KEEP_DIV=$kd RT_PIPE_SHARE="4 2" RT_PIPE_M="3 4 6" timeout 300 python tools/pipe_probe.py c3 1 2>&1 | tail -6 | tee -a $OUT
done
