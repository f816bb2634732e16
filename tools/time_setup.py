#!/usr/bin/env python
"""Pair set-up rate (nid_set_pairs_u16 + nid_prepare_pairs) at C2 geometry from pinned host buffers; developer tool."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
nid = importlib.import_module("nid-pose-estimation_b200")
synth = importlib.import_module("nid-pose-estimation_b200.synth")
from oracle import binding as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cell, bins = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4, 16)
base = [synth.make_pair(1000 + i, 480, 640) for i in range(4)]
N = 480 * 640
d16 = torch.empty((n, N), dtype=torch.uint16).pin_memory()
im0 = torch.empty((n, N), dtype=torch.uint8).pin_memory()
im1 = torch.empty((n, N), dtype=torch.uint8).pin_memory()
for i in range(n):
    p = base[i % 4]
    d16[i] = torch.from_numpy(p.depth0_u16.reshape(-1).copy())
    im0[i] = torch.from_numpy(p.im0.reshape(-1).copy())
    im1[i] = torch.from_numpy(p.im1.reshape(-1).copy())
T = np.stack([base[i % 4].T_wc0 for i in range(n)])
K = np.stack([base[i % 4].intr for i in range(n)])
init = np.stack([orc.se3_to_mat16(orc.reference_perturbation(base[i % 4].T_wc1)) for i in range(n)])
ctx = nid.Context(480, 640, cell, bins, n_pairs=n, max_jobs=min(n, 128))
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    keep = ctx.set_pairs_u16(0, d16.numpy(), im0.numpy(), im1.numpy(), T, K)
    t1 = time.perf_counter()
    ctx.sync()
    t2 = time.perf_counter()
    nc, href = ctx.prepare_pairs(0, init)
    t3 = time.perf_counter()
    print(f"rep {rep}: n={n} cell={cell} bins={bins}  set_pairs_u16 issue {1e3*(t1-t0):.1f} ms, done {1e3*(t2-t0):.1f} ms; "
          f"prepare_pairs {1e3*(t3-t2):.1f} ms  => {n/(t3-t0):.0f} pairs/s", flush=True)
p7 = np.stack([orc.reference_perturbation(base[i % 4].T_wc1) for i in range(min(n, 128))])
jp = np.arange(min(n, 128), dtype=np.int32)
ctx.solve_jobs(p7, jp)
t0 = time.perf_counter()
out, st = ctx.solve_jobs(p7, jp)
t1 = time.perf_counter()
print(f"solve_jobs {len(jp)} problems: {1e3*(t1-t0):.1f} ms => {len(jp)/(t1-t0):.0f} solves/s, iters {st[:,0].mean():.1f}")
