#!/bin/bash
# prints the warp/sample probe line of bench.py for each library variant given (base = the in-tree library)
for v in "$@"; do
  if [ "$v" = base ]; then unset NID_B200_LIB; else export NID_B200_LIB=$PWD/build/libvar_$v.so; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --cpu-budget 0.1 --solves 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); q=d['roofline']['warp_sample_probe']
print('$v', 'value %.0f' % d['value'], 'probe ms %.4f achieved %.0f GB/s frac %.3f (moved %.3f)' % (q['ms_per_launch'], q['achieved'], q['frac'], q['frac_moved']))"
done
