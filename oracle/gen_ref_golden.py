"""Generates tests/golden/ref_gpu_*.npz: outputs of the REFERENCE'S OWN CUDA code (unmodified
CudaPoints3d.cu + g2o/g2o/core/computeH.cu, and CudaComputeHref.cu with its cudaMemset fixed, compiled by
oracle/Makefile into oracle/_ref/) on seeded synthetic pairs: one case per knot table the reference ships
(6, 8, 10, 12, 14 bins, computeH.cu:99-112). The whole chain is the reference's: Calculate3Dpoint ->
CudaComputeHref -> CudaComputeH; the oracle only supplies the poses. Must run on a GPU box:

    gpurun -- python oracle/gen_ref_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/

The CPU test tests/test_golden.py re-creates the same inputs from the seed and checks the CPU oracle
against these vectors; that is the pin the oracle stands on (the reference ships no test vectors).
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [  # name, seed, cell, bins, pose perturbations (xi applied on the left of the initial pose)
    ("c4_b10", 1000, 4, 10),
    ("c8_b8", 1001, 8, 8),
    ("c16_b10", 1002, 16, 10),
    ("c4_b14", 1003, 4, 14),
    ("c1_b8", 1004, 1, 8),
    ("c4_b6", 1005, 4, 6),
    ("c4_b12", 1006, 4, 12),
]
XIS = np.array([[0, 0, 0, 0, 0, 0], [0.002, -0.001, 0.0015, 0.004, -0.003, 0.002], [-0.004, 0.003, -0.002, -0.01, 0.006, 0.004]])
SUB = 97  # stride of the points3d subsample kept in the fixture


def main():
    from oracle import binding as orc
    from oracle import ref_gpu
    synth = importlib.import_module("nid-pose-estimation_b200.synth")
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, seed, cell, bins in CASES:
        p = synth.make_pair(seed, ref_gpu.SAFE_ROWS, ref_gpu.SAFE_COLS)
        pose0 = orc.reference_perturbation(p.T_wc1)
        P = orc.Problem(p.im0, p.depth0, p.im1, p.T_wc0, p.intr, cell, bins)
        P.set_quirks(1, 1)
        nc, href = P.prepare(pose0)
        bv, bi = P.ref_weights()
        R = ref_gpu.RefGpu(p, cell, bins)
        pts = R.points3d()
        # a2 by the reference itself (CudaComputeHref.cu), at the initial pose
        r_nc, r_bv, r_bi, r_href = R.compute_href(orc.se3_to_mat16(pose0))
        inb = ~np.isnan(r_bv[0::4])
        assert np.array_equal(r_nc, nc), (r_nc, nc)
        print(f"{name}: reference CudaComputeHref vs oracle  n_c equal, max rel dHref "
              f"{np.nanmax(np.abs(r_href - href) / np.abs(href)):.3e}, in-bounds pixels {int(inb.sum())}", flush=True)
        # a6-a9 on the reference's own prepare output (Href as the driver passes it on, NaN for inactive cells)
        # (pixels that were out of bounds at the initial pose carry NaN weights in the reference's GPU arrays, which
        # would poison the joint histogram if they come in bounds later (SURVEY B-6); the CPU edge keeps zero rows for
        # them, types_six_dof_expmap.cpp:639-653, and that is the convention pinned here)
        # The same pixels carry `bs_index = NAN` converted to int (CudaComputeHref.cu:82,126; SURVEY B-5: INT_MIN after
        # constant folding), which CalculateProKernel uses unguarded as a histogram index (computeH.cu:237-243): index 0
        # is substituted so that the unmodified kernel stays inside its buffers; with zero weights it adds nothing.
        bad_index = sorted(set(int(x) for x in np.unique(r_bi[~inb])))
        print(f"{name}: bs_index written by the reference for pixels without a sample: {bad_index[:4]}", flush=True)
        R.set_prepare(r_nc, np.where(np.isnan(r_bv), 0.0, r_bv), np.where(inb, r_bi, 0), r_href)
        Hts, Hjs, ders, poses = [], [], [], []
        for xi in XIS:
            pose = orc.se3_mul(orc.se3_exp(xi), pose0)
            M = orc.se3_to_mat16(pose)
            Ht, Hj, der = R.compute_h(M, True)
            Ht2, Hj2, _ = R.compute_h(M, False)
            assert np.array_equal(Ht, Ht2, equal_nan=True) or np.allclose(Ht, Ht2, rtol=1e-12, equal_nan=True)
            Hts.append(Ht); Hjs.append(Hj); ders.append(der); poses.append(M)
        R.close()
        np.savez_compressed(os.path.join(out_dir, f"ref_gpu_{name}.npz"), seed=seed, rows=p.rows, cols=p.cols, cell=cell,
                            bins=bins, pose0=pose0, xis=XIS, poses=np.array(poses), Ht=np.array(Hts), Hj=np.array(Hjs),
                            der=np.array(ders), n_c=nc, href=href, pts_sub=pts.reshape(-1, 3)[::SUB],
                            ref_n_c=r_nc, ref_href=r_href, ref_bs_value_sub=r_bv.reshape(-1, 4)[::SUB],
                            ref_bs_index_sub=r_bi[::SUB], ref_inb_count=int(inb.sum()),
                            ref_bs_index_sum=int(r_bi[inb].astype(np.int64).sum()),
                            ref_bs_value_colsum=r_bv.reshape(-1, 4)[inb].sum(axis=0),
                            pts_nan=int(np.isnan(pts).sum()), im0_sum=int(p.im0.astype(np.int64).sum()),
                            im1_sum=int(p.im1.astype(np.int64).sum()), d16_sum=int(p.depth0_u16.astype(np.int64).sum()))
        # quick report against the oracle
        Ho = [P.eval(orc.se3_mul(orc.se3_exp(xi), pose0), True) for xi in XIS]
        dj = max(np.nanmax(np.abs(ders[k] - Ho[k][3]) / np.nanmax(np.abs(Ho[k][3]), axis=1, keepdims=True)) for k in range(len(XIS)))
        dh = max(np.nanmax(np.abs(Hjs[k] - Ho[k][1]) / Ho[k][1]) for k in range(len(XIS)))
        print(f"{name}: reference CUDA vs oracle  max rel dHj {dh:.3e}  max rel dJ {dj:.3e}", flush=True)


if __name__ == "__main__":
    main()
