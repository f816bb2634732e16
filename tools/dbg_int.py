import importlib, sys, dataclasses
import numpy as np
sys.path.insert(0, '.')
nid = importlib.import_module("nid-pose-estimation_b200")
synth = importlib.import_module("nid-pose-estimation_b200.synth")
from oracle import binding as orc
p = synth.make_pair(1000, 120, 160)
d=p.depth0.copy(); d[0,:]=0; d[:,0]=0
p = dataclasses.replace(p, im1=p.im0.copy(), depth0=d)
pose_id = orc.se3_from_mat16(synth.mat16_inverse(p.T_wc0))
P = orc.Problem(p.im0, p.depth0, p.im1, p.T_wc0, p.intr, 2, 16, threads=4)
P.set_quirks(0, 1)
M = orc.se3_to_mat16(pose_id)
nco, hrefo = P.prepare(pose_id)
Hto, Hjo, erro, Jo = P.eval(pose_id, True)
px = P.pixels(pose_id)
print("oracle J", Jo)
print("frac u", np.unique(np.round(px[:,0]-np.floor(px[:,0]), 6), return_counts=True))
for path in (1, 2):
    ctx = nid.Context(p.rows, p.cols, 2, 16)
    ctx.set_option("path", path)
    ctx.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
    nc, href = ctx.prepare(0, M)
    Ht, Hj, J = ctx.eval(0, M, True)
    print("path", path, "nc eq", np.array_equal(nc, nco), "Ht", np.max(np.abs(Ht-Hto)/Hto), "Hj", np.max(np.abs(Hj-Hjo)/Hjo))
    print(J)
    got = ctx.warp_sample_f64(0, M)
    m = ~np.isnan(px[:,0])
    print("vj", np.sum(got[m,6]), np.sum(px[m,6]), "rows of gy diff", np.unique(np.nonzero(np.abs(got[:,4]-px[:,4])>1e-9)[0]//160)[:10], np.unique(np.nonzero(np.abs(got[:,4]-px[:,4])>1e-9)[0]%160)[:10]);print("u eq", np.array_equal(got[m,0], px[m,0]), "ic maxdiff", np.nanmax(np.abs(got[m,2]-px[m,2])), "gx maxdiff", np.nanmax(np.abs(got[m,3]-px[m,3])), np.nanmax(np.abs(got[m,4]-px[m,4])))
