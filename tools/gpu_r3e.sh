#!/bin/bash
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in base old; do
  if [ "$v" = base ]; then unset NID_B200_LIB; else export NID_B200_LIB=$PWD/build/libvar_$v.so; fi
  for g in "480 640 4 16" "480 640 16 10" "960 1280 4 32" "480 640 1 8" "480 640 8 10"; do echo "== $v: $(timeout 300 python tools/time_config.py $g 48 10 2>&1 | grep 'sorted want_jac=1')"; done
  timeout 100 python tools/time_single.py 2>&1 | tail -1
done
