#!/bin/bash
tag=${1:-r02w}
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
bash tools/gpu_var.sh base
timeout 300 python tools/time_config.py 480 640 16 10 96 10 2>&1 | grep "sorted want_jac=1"
timeout 900 python bench.py --steps 20 --warmup 3 --cpu-budget 6 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "clocks", "pair_setup", "pose_solves", "c4", "c5", "reference_api_shim"):
    print(k, json.dumps(d.get(k))[:700])
print("e2e", d["e2e"]["value"])
PY
