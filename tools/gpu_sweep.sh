#!/bin/bash
# sweep of developer knobs: each argument is "pairs:NID_OPTS" (e.g. "24:task_px=32")
tag=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  pairs=${spec%%:*}; opt=${spec#*:}
  echo "== pairs=$pairs NID_OPTS=$opt"
  NID_OPTS=$opt timeout 300 python bench.py --steps 20 --warmup 3 --cpu-budget 0.2 --pairs $pairs 2> gpurun_out/${tag}_sweep.err | tee -a gpurun_out/${tag}_sweep.jsonl | python -c "
import sys, json
s = sys.stdin.read()
try:
    d = json.loads(s.strip().splitlines()[-1])
    sh = d['roofline']['kernel_share_of_step']
    print('value %.0f e2e %.0f us/eval %.2f shares ' % (d['value'], d['e2e']['value'], 1e3*d['ms_per_step']/d['config']['evals_per_step']) + ' '.join('%s=%.2f' % (k, v) for k, v in sh.items()))
except Exception as e:
    print('FAILED', s[-300:])
"
  tail -2 gpurun_out/${tag}_sweep.err
done
