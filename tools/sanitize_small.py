"""Small evaluations of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nid = importlib.import_module("nid-pose-estimation_b200")
synth = importlib.import_module("nid-pose-estimation_b200.synth")
from oracle import binding as orc
for (rows, cols, cell, bins) in ((120, 160, 2, 16), (120, 160, 4, 10), (96, 128, 1, 9), (120, 160, 2, 40)):
    p = synth.make_pair(1000, rows, cols)
    pose0 = orc.reference_perturbation(p.T_wc1)
    M0 = orc.se3_to_mat16(pose0)
    ctx = nid.Context(rows, cols, cell, bins, n_pairs=2, max_jobs=12)
    for s in range(2):
        ctx.set_pair(s, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
        ctx.prepare(s, M0)
    jp = np.arange(12, dtype=np.int32) % 2
    poses = np.stack([M0] * 12)
    ctx.eval_jobs(poses, jp, True)
    ctx.eval(0, M0, False)
    ctx.solve_jobs(np.stack([pose0] * 12), jp, 3)
    ctx.solve(0, pose0, 2)
    ctx.warp_sample_jobs(poses, jp)
    ctx.hard_eval_jobs(poses, jp)
    ctx.set_target(1, p.im0)
    ctx.prepare(1, M0)
    ctx.eval(1, M0, True)
    ctx.close()
    print("ok", rows, cols, cell, bins)
