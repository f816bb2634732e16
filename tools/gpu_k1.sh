#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "warp_sample or kernel1 or more_jobs" 2>&1 | tail -5
for v in base "$@"; do
  if [ "$v" = base ]; then unset NID_B200_LIB; else export NID_B200_LIB=$PWD/build/libvar_$v.so; fi
  timeout 120 python tools/time_k1.py 96 30 2>&1 | tail -1
done
unset NID_B200_LIB
timeout 120 python tools/time_k1.py 96 30 f64 2>&1 | tail -1
