#!/usr/bin/env python
"""C5 shape: hard-binned NID (NID_standard_property) of many poses of one 640x480 pair, cell=16, 8 bins: evals/s."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nid = importlib.import_module("nid-pose-estimation_b200")
synth = importlib.import_module("nid-pose-estimation_b200.synth")
from oracle import binding as orc
jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
p = synth.make_pair(1000, 480, 640)
gt = orc.se3_from_mat16(synth.mat16_inverse(p.T_wc1))
ctx = nid.Context(480, 640, 16, 8, n_pairs=1, max_jobs=jobs)
ctx.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
rng = np.random.default_rng(0)
poses = np.stack([orc.se3_to_mat16(orc.se3_mul(orc.se3_exp(rng.uniform(-0.05, 0.05, 6)), gt)) for _ in range(jobs)])
jp = np.zeros(jobs, dtype=np.int32)
tot, cells = ctx.hard_eval_jobs(poses, jp)
to, _ = orc.hard_nid(p.im0, p.depth0, p.im1, p.T_wc0, poses[3], p.intr, 16, 8)
assert abs(tot[3] - to) <= 1e-12 * abs(to), (tot[3], to)
t0 = time.perf_counter()
for _ in range(5):
    ctx.hard_eval_jobs(poses, jp)
dt = (time.perf_counter() - t0) / 5
print(f"hard-binned NID, 640x480, 16x16 cells, 8 bins: {jobs / dt:.0f} evals/s ({dt / jobs * 1e6:.1f} us/eval), 64^2 x 6 sweep = {24576 * dt / jobs:.2f} s on one GPU")
