#!/usr/bin/env python
"""Latency of the single-pair flow the reference's NID_pose_estimation runs per frame pair: set-up, one evaluation, one LM solve."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nid = importlib.import_module("nid-pose-estimation_b200")
synth = importlib.import_module("nid-pose-estimation_b200.synth")
from oracle import binding as orc
rows, cols, cell, bins = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (480, 640, 4, 16)
p = synth.make_pair(1000, rows, cols)
pose0 = orc.reference_perturbation(p.T_wc1)
M0 = orc.se3_to_mat16(pose0)
ctx = nid.Context(rows, cols, cell, bins, n_pairs=1, max_jobs=1)
for kv in filter(None, os.environ.get("NID_OPTS", "").split(",")):
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
def t(f, n=20):
    f(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(n): f()
    ctx.sync()
    return (time.perf_counter() - t0) / n * 1e3
print(f"{rows}x{cols} cell={cell} bins={bins} opts={os.environ.get('NID_OPTS', '')}")
if os.environ.get("NID_SOLVE_ONLY"):  # under ncu: one set-up and one solve
    ctx.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr); ctx.prepare(0, M0); ctx.solve(0, pose0); sys.exit(0)
print("set_pair   %.3f ms" % t(lambda: ctx.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)))
print("prepare    %.3f ms" % t(lambda: ctx.prepare(0, M0)))
print("eval cost  %.3f ms" % t(lambda: ctx.eval(0, M0, False), 100))
print("eval c+J   %.3f ms" % t(lambda: ctx.eval(0, M0, True), 100))
print("solve(10)  %.3f ms" % t(lambda: ctx.solve(0, pose0), 10))
# warm per-kernel times (CUDA events between the launches) of one evaluation with 1 and with 4 jobs in flight
for nj in (1, 4):
    c2 = nid.Context(rows, cols, cell, bins, n_pairs=1, max_jobs=nj)
    c2.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr); c2.prepare(0, M0)
    poses = np.stack([orc.se3_to_mat16(orc.se3_mul(orc.se3_exp(1e-3 * (k + 1) * np.ones(6)), pose0)) for k in range(nj)])
    c2.set_option("time_kernels", 1)
    for _ in range(30): c2.eval_jobs(poses, np.zeros(nj, np.int32), True)
    kt = c2.kernel_times()
    print(f"{nj} job(s): " + "  ".join("%s %.1f us" % (k, 1e3 * ms / max(n, 1)) for k, (ms, n) in kt.items()))
    c2.close()
