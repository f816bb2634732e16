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
print("set_pair   %.3f ms" % t(lambda: ctx.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)))
print("prepare    %.3f ms" % t(lambda: ctx.prepare(0, M0)))
print("eval cost  %.3f ms" % t(lambda: ctx.eval(0, M0, False), 100))
print("eval c+J   %.3f ms" % t(lambda: ctx.eval(0, M0, True), 100))
print("solve(10)  %.3f ms" % t(lambda: ctx.solve(0, pose0), 10))
