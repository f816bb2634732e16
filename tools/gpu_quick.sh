#!/bin/bash
# quick GPU visit: parity tests + a short bench (+ optional sweep of NID_OPTS given as arguments)
tag=${1:-q}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${tag}_pytest.log
for opt in "" "$@"; do
  echo "== NID_OPTS=$opt"
  NID_OPTS=$opt timeout 300 python bench.py --steps 20 --warmup 3 --cpu-budget 0.4 2> gpurun_out/${tag}_bench.err | tee -a gpurun_out/${tag}_bench.jsonl | python -c "
import sys, json
s = sys.stdin.read()
try:
    d = json.loads(s.strip().splitlines()[-1])
    sh = d['roofline']['kernel_share_of_step']
    print('value %.0f e2e %.0f ms/step %.4f shares ' % (d['value'], d['e2e']['value'], d['ms_per_step']) + ' '.join('%s=%.2f' % (k, v) for k, v in sh.items()))
except Exception as e:
    print('FAILED', s[-300:])
"
  tail -3 gpurun_out/${tag}_bench.err
done
