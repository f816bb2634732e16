#!/bin/bash
for n in "$@"; do
  out=$(timeout 300 python bench.py --steps 10 --warmup 3 --cpu-budget 0.4 --pairs $n 2>&1 | tail -1)
  echo "== pairs $n"
  echo "$out" | python -c "
import sys, json
s = sys.stdin.read()
try:
    d = json.loads(s)
    sh = d['roofline']['kernel_share_of_step']
    print('value %.0f e2e %.0f ms/step %.4f us/eval %.2f shares ' % (d['value'], d['e2e']['value'], d['ms_per_step'], 1e3*d['ms_per_step']/d['config']['evals_per_step']) + ' '.join('%s=%.2f' % (k, v) for k, v in sh.items()))
except Exception as e:
    print('FAILED', s[-300:])
"
done
