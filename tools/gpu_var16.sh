#!/bin/bash
# time the reference's default geometry (16x16 cells, 10 bins, 96 pairs) for build variants: tools/gpu_var16.sh v1 v2 ...
for v in "$@"; do
  if [ "$v" = base ]; then unset NID_B200_LIB; else export NID_B200_LIB=$PWD/build/libvar_$v.so; fi
  echo "== $v: $(timeout 300 python tools/time_config.py 480 640 16 10 96 20 2>&1 | grep 'path=sorted want_jac=1')"
done
