#!/bin/bash
# usage: tools/bench_variants.sh base var1 var2 ...   (build/libvar_<name>.so); prints value / e2e / shares
for v in "$@"; do
  if [ "$v" = base ]; then unset NID_B200_LIB; else export NID_B200_LIB=$PWD/build/libvar_$v.so; fi
  out=$(timeout 300 python bench.py --steps 10 --warmup 3 --cpu-budget 0.4 2>&1 | tail -1)
  echo "== $v"
  echo "$out" | python -c "
import sys, json
s = sys.stdin.read()
try:
    d = json.loads(s)
    sh = d['roofline']['kernel_share_of_step']
    print('value %.0f e2e %.0f ms/step %.4f shares ' % (d['value'], d['e2e']['value'], d['ms_per_step']) + ' '.join('%s=%.2f' % (k, v) for k, v in sh.items()))
except Exception as e:
    print('FAILED', s[-300:])
"
done
