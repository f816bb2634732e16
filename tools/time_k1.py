#!/usr/bin/env python
"""Kernel 1 alone (nid_warp_sample_jobs) on 96 pair slots of C2 shape, raw 16-bit depth in, float4 out: ms per launch
and GB/s by SURVEY 8(d)'s 24 B/px. Developer tool (NID_B200_LIB selects a build variant)."""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nid = importlib.import_module("nid-pose-estimation_b200")
synth = importlib.import_module("nid-pose-estimation_b200.synth")
from oracle import binding as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
u16 = (sys.argv[3] != "f64") if len(sys.argv) > 3 else True
base = [synth.make_pair(1000 + i, 480, 640) for i in range(3)]
ctx = nid.Context(480, 640, 4, 16, n_pairs=n, max_jobs=n)
if u16:
    keep = ctx.set_pairs_u16(0, np.stack([base[i % 3].depth0_u16 for i in range(n)]), np.stack([base[i % 3].im0 for i in range(n)]),
                             np.stack([base[i % 3].im1 for i in range(n)]), np.stack([base[i % 3].T_wc0 for i in range(n)]),
                             np.stack([base[i % 3].intr for i in range(n)]))
    ctx.sync()
else:
    for i in range(n):
        p = base[i % 3]
        ctx.set_pair(i, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
pose0 = [orc.reference_perturbation(p.T_wc1) for p in base]
rng = np.random.default_rng(1)
jp = np.arange(n, dtype=np.int32)
def poses():
    return np.stack([orc.se3_to_mat16(orc.se3_mul(orc.se3_exp(rng.uniform(-1, 1, 6) * 3e-3), pose0[i % 3])) for i in range(n)])
P = [poses() for _ in range(4)]
for k in range(3):
    ctx.warp_sample_jobs(P[k], jp, fetch=False)
ctx.event_record(0)
for k in range(steps):
    ctx.warp_sample_jobs(P[k % 4], jp, fetch=False)
ctx.event_record(1)
ms = ctx.event_elapsed_ms() / steps
gbs = 480 * 640 * 24 * n / (ms * 1e-3) / 1e9
print(json.dumps({"lib": os.environ.get("NID_B200_LIB", "base"), "depth": "u16" if u16 else "f64", "ms_per_launch": ms,
                  "GBps_algorithmic_24B": gbs, "frac_of_6540": gbs / 6540.2}))
