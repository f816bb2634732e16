#!/usr/bin/env python
"""Writes one synthetic frame pair in the ETH-CVG layout the reference binaries read
(NID_pose_estimation.cpp:84-113): rgb/NNNN.png (8-bit, 3 channels), depth/NNNN.png (uint16, metres*5000),
groundtruth.txt (TUM lines `ts tx ty tz qx qy qz qw`, pose of frame k on line k) and a config YAML with the
keys of config_eth_cvg.yaml.

    python tools/gen_synth.py OUT_DIR [--seed 1000] [--rows 480] [--cols 640] [--id0 0] [--id1 1] [--pgm]

--pgm writes binary PGM instead of PNG (the decoder-free fallback apps/nid_io.hpp also reads).
The colour frames carry three distinct channels (a fixed tint per channel) so that the reference's
BGR-as-RGB grayscale quirk is exercised; `gray_of_rgb` below is the conversion the binaries apply.
"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gray_of_rgb(rgb: np.ndarray, shift: int = 14) -> np.ndarray:
    """imread(UNCHANGED) + cvtColor(CV_RGB2GRAY) on BGR data, OpenCV's 8-bit fixed-point kernel
    (shift 14: OpenCV 2.4/3.x, the reference's dependency; shift 15: OpenCV >= 3.4.3 / 4.x)."""
    R, G, B = (rgb[..., i].astype(np.uint32) for i in range(3))
    if shift == 15:
        return ((B * 9798 + G * 19235 + R * 3735 + 16384) >> 15).astype(np.uint8)
    return ((B * 4899 + G * 9617 + R * 1868 + 8192) >> 14).astype(np.uint8)


def tint(gray: np.ndarray) -> np.ndarray:
    g = gray.astype(np.int32)
    rgb = np.stack([np.clip(g + 7, 0, 255), g, np.clip(g - 5, 0, 255)], axis=-1).astype(np.uint8)
    return rgb


def quat_from_R(R):
    tr = np.trace(R)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        return np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
    q = np.zeros(4)
    q[i] = 0.25 * s
    q[3] = (R[k, j] - R[j, k]) / s
    q[j] = (R[j, i] + R[i, j]) / s
    q[k] = (R[k, i] + R[i, k]) / s
    return q


def write_pgm(path, a):
    a = np.ascontiguousarray(a)
    maxv = 65535 if a.dtype == np.uint16 else 255
    with open(path, "wb") as f:
        f.write(f"P5\n{a.shape[1]} {a.shape[0]}\n{maxv}\n".encode())
        f.write(a.astype(">u2").tobytes() if a.dtype == np.uint16 else a.tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--rows", type=int, default=480)
    ap.add_argument("--cols", type=int, default=640)
    ap.add_argument("--id0", type=int, default=0)
    ap.add_argument("--id1", type=int, default=1)
    ap.add_argument("--pgm", action="store_true")
    ap.add_argument("--cell", type=int, default=None)
    ap.add_argument("--bins", type=int, default=None)
    args = ap.parse_args()
    synth = importlib.import_module("nid-pose-estimation_b200.synth")
    p = synth.make_pair(args.seed, args.rows, args.cols)
    out = os.path.abspath(args.out)
    os.makedirs(os.path.join(out, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(out, "depth"), exist_ok=True)
    ids = [f"{args.id0:04d}", f"{args.id1:04d}"]
    frames = [tint(p.im0), tint(p.im1)]
    if args.pgm:
        for i, fr in zip(ids, frames):
            write_pgm(os.path.join(out, "rgb", i + ".pgm"), gray_of_rgb(fr))
        write_pgm(os.path.join(out, "depth", ids[0] + ".pgm"), p.depth0_u16)
    else:
        import cv2
        for i, fr in zip(ids, frames):
            cv2.imwrite(os.path.join(out, "rgb", i + ".png"), fr[..., ::-1])  # cv2 wants BGR in memory
        cv2.imwrite(os.path.join(out, "depth", ids[0] + ".png"), p.depth0_u16)
    n = max(args.id0, args.id1) + 1
    lines = []
    for k in range(n):
        T = (p.T_wc0 if k == args.id0 else p.T_wc1 if k == args.id1 else np.eye(4).T.reshape(16)).reshape(4, 4).T
        q = quat_from_R(T[:3, :3])
        lines.append(" ".join([str(k)] + [repr(float(v)) for v in (*T[:3, 3], *q)]))
    open(os.path.join(out, "groundtruth.txt"), "w").write("\n".join(lines) + "\n")
    cfg = ["%YAML:1.0", f"image0_id: '{ids[0]}'", f"image1_id: '{ids[1]}'", "image0_type: rgb", "image1_type: rgb", "",
           "use_groundtruth: '1'", "", "dataset: eth_cvg", "", f"im_address: {out}/", "", "depth_factor: 5000.0", "",
           f"fx: {float(p.intr[0])!r}", f"fy: {float(p.intr[1])!r}", f"cx: {float(p.intr[2])!r}", f"cy: {float(p.intr[3])!r}", "",
           "#make sure you have GPU and CUDA. 1: use GPU, 0: use CPU", "use_gpu: 1"]
    if args.cell:
        cfg.append(f"cell: {args.cell}")
    if args.bins:
        cfg.append(f"bin_num: {args.bins}")
    open(os.path.join(out, "config.yaml"), "w").write("\n".join(cfg) + "\n")
    # what the binaries will see after the grayscale conversion (for tests)
    np.savez(os.path.join(out, "expected_inputs.npz"), im0=gray_of_rgb(frames[0]), im1=gray_of_rgb(frames[1]),
             depth_u16=p.depth0_u16, T_wc0=p.T_wc0, T_wc1=p.T_wc1, intr=p.intr)
    print(out)


if __name__ == "__main__":
    main()
