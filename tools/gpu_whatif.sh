#!/bin/bash
# what-if visit: variants in build/libvar_*.so given as arguments (plus "base"), then NID_OPTS combos after "--"
mkdir -p gpurun_out
run() {  # $1 label
  out=$(timeout 300 python bench.py --steps 20 --warmup 3 --cpu-budget 0.2 --solves 0 2>&1 | tail -1)
  echo "$out" | python -c "
import sys, json
s = sys.stdin.read()
try:
    d = json.loads(s)
    r = d['roofline']; sh = r['kernel_share_of_step']
    ms = d['ms_per_step']
    print('%-28s value %7.0f e2e %7.0f ms/step %.4f  ' % ('$1', d['value'], d['e2e']['value'], ms) + ' '.join('%s=%.3fms' % (k, v * ms) for k, v in sh.items()))
except Exception as e:
    print('$1 FAILED', s[-300:])
"
}
libs=(); opts=(); sep=0
for a in "$@"; do if [ "$a" = "--" ]; then sep=1; elif [ $sep = 0 ]; then libs+=("$a"); else opts+=("$a"); fi; done
for v in "${libs[@]}"; do
  if [ "$v" = base ]; then unset NID_B200_LIB; else export NID_B200_LIB=$PWD/build/libvar_$v.so; fi
  run "lib=$v"
done
unset NID_B200_LIB
for o in "${opts[@]}"; do NID_OPTS=$o run "opts=$o"; done
