#!/usr/bin/env python
"""Phase times inside k_assemble (build variant -DNID_ASM_TRACE prints %globaltimer deltas of one cell): a few warm evaluations of one job."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nid = importlib.import_module("nid-pose-estimation_b200")
synth = importlib.import_module("nid-pose-estimation_b200.synth")
from oracle import binding as orc
rows, cols, cell, bins = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (480, 640, 4, 16)
p = synth.make_pair(1000, rows, cols)
M0 = orc.se3_to_mat16(orc.reference_perturbation(p.T_wc1))
ctx = nid.Context(rows, cols, cell, bins, n_pairs=1, max_jobs=1)
ctx.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr); ctx.prepare(0, M0)
for _ in range(6): ctx.eval(0, M0, True)
