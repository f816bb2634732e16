#!/bin/bash
# latency mode of the LM driver: per-call times of the single-pair flow for option sets given as arguments
mkdir -p gpurun_out
for o in "" "$@"; do
  NID_LM_TIMES=1 NID_OPTS=$o timeout 300 python tools/time_single.py 480 640 4 16 > gpurun_out/lat_one.log 2>&1
  echo "opts=$o: $(grep -o 'solve(10) *[0-9.]* ms' gpurun_out/lat_one.log) $(grep -o 'eval c+J *[0-9.]* ms' gpurun_out/lat_one.log) $(grep -o 'prepare *[0-9.]* ms' gpurun_out/lat_one.log) | $(grep 'lm rounds' gpurun_out/lat_one.log | tail -1)"
done
