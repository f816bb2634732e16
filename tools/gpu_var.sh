#!/bin/bash
# time C2 evaluation (96 pairs) for build variants: tools/gpu_var.sh v1 v2 ...   ("base" = the in-tree library)
for v in "$@"; do
  if [ "$v" = base ]; then unset NID_B200_LIB; else export NID_B200_LIB=$PWD/build/libvar_$v.so; fi
  echo "== $v: $(timeout 300 python tools/time_config.py 480 640 4 16 96 20 2>&1 | grep 'path=sorted want_jac=1')"
done
