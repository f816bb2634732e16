#!/usr/bin/env python
"""Device-time evals/s of one geometry on both evaluation paths: python tools/time_config.py rows cols cell bins [pairs] [steps]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nid = importlib.import_module("nid-pose-estimation_b200")
synth = importlib.import_module("nid-pose-estimation_b200.synth")
from oracle import binding as orc
rows, cols, cell, bins = (int(v) for v in sys.argv[1:5])
pairs = int(sys.argv[5]) if len(sys.argv) > 5 else 24
steps = int(sys.argv[6]) if len(sys.argv) > 6 else 10
P = [synth.make_pair(1000 + i, rows, cols) for i in range(min(pairs, 4))]
for path in (1, 2):
    try:
        ctx = nid.Context(rows, cols, cell, bins, n_pairs=pairs, max_jobs=pairs)
        ctx.set_option("path", path)
        for kv in filter(None, os.environ.get("NID_OPTS", "").split(",")):
            k, v = kv.split("=")
            ctx.set_option(k, int(v))
    except Exception as e:
        print("path", path, "unavailable:", e)
        continue
    pose0 = []
    for s in range(pairs):
        p = P[s % len(P)]
        p0 = orc.reference_perturbation(p.T_wc1)
        ctx.set_pair(s, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
        ctx.prepare(s, orc.se3_to_mat16(p0))
        pose0.append(orc.se3_to_mat16(p0))
    poses = np.stack(pose0)
    jp = np.arange(pairs, dtype=np.int32)
    for want_jac in (True, False):
        for _ in range(3):
            ctx.stage_jobs(poses, jp); ctx.eval_staged(pairs, want_jac)
        ctx.sync()
        ctx.event_record(0)
        for _ in range(steps):
            ctx.stage_jobs(poses, jp); ctx.eval_staged(pairs, want_jac)
        ctx.event_record(1)
        ms = ctx.event_elapsed_ms()
        ctx.set_option("time_kernels", 1)
        for _ in range(3):
            ctx.stage_jobs(poses, jp); ctx.eval_staged(pairs, want_jac)
        ctx.sync()
        kt = ctx.kernel_times()
        ctx.set_option("time_kernels", 0)
        split = " ".join(f"{k}={v[0] / max(v[1], 1) / pairs * 1e3:.2f}us" for k, v in kt.items() if v[1])
        print(f"{rows}x{cols} cell={cell} bins={bins} path={'natural' if path == 1 else 'sorted'} want_jac={int(want_jac)}: "
              f"{pairs * steps / ms * 1e3:.0f} evals/s ({ms / steps / pairs * 1e3:.1f} us/eval)  per eval: {split}")
    ctx.close()
