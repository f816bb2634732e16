#!/bin/bash
tag=${1:-r02j}
mkdir -p gpurun_out
bash tools/gpu_k1.sh g1 g2 g1b16 2>&1 | tee gpurun_out/${tag}_k1.log | grep -E "lib|passed|failed"
timeout 900 python bench.py --steps 20 --warmup 3 --cpu-budget 6 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/${tag}_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02j_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "gpu_launches", "clocks", "pair_setup", "pose_solves", "c4", "c5", "old_gpu_path", "cpu_baseline"):
    print(k, json.dumps(d.get(k))[:900])
print("e2e", d["e2e"]["value"]); r = d["roofline"]; print({k: r[k] for k in ("achieved", "frac", "kernel_ms_per_launch", "fp64", "kernel_share_of_step")}); print(r["warp_sample_probe"])
PY
