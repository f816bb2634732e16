"""The two config-driven binaries (apps/) and their input side (f2 dataset ingest, f3 outputs).

CPU: apps/nid_io.hpp against OpenCV itself (python cv2 performs the reference's imread + cvtColor calls).
GPU: the binaries on a synthetic ETH-CVG-layout dataset against the C-ABI driven from Python and the oracle."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")


@pytest.fixture(scope="module")
def apps_built():
    if not os.path.exists(os.path.join(ROOT, "nid-pose-estimation_b200", "libnid_b200.so")):
        pytest.skip("libnid_b200.so not built")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "apps")], stdout=subprocess.DEVNULL)
    return BIN


def _gen(tmp_path, *extra):
    out = str(tmp_path / "ds")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_synth.py"), out, *extra], stdout=subprocess.DEVNULL)
    return out


def _check(apps_built, mode, path, *extra):
    return subprocess.check_output([os.path.join(apps_built, "nid_io_check"), mode, path, *extra]).decode()


def _gray14(rgb):
    """OpenCV 2.4/3.x CV_RGB2GRAY applied to BGR data (the reference's dependency and call)."""
    R, G, B = (rgb[..., i].astype(np.uint32) for i in range(3))
    return ((B * 4899 + G * 9617 + R * 1868 + 8192) >> 14).astype(np.uint8)


def _sums(a):
    a = np.ascontiguousarray(a).reshape(-1).astype(np.uint64)
    w = (np.arange(a.size, dtype=np.uint64) % 1009 + 1)
    return int(a.sum()), int((a * w).sum())


def test_png_decoder_and_gray_quirk_match_opencv(apps_built, tmp_path):
    cv2 = pytest.importorskip("cv2")
    ds = _gen(tmp_path, "--rows", "96", "--cols", "130")
    # 8-bit colour frame: samples in file (RGB) order, then the reference's gray conversion
    path = os.path.join(ds, "rgb", "0000.png")
    bgr = cv2.imread(path, cv2.IMREAD_UNCHANGED)                # NID_pose_estimation.cpp:91
    gray = cv2.cvtColor(bgr, cv2.COLOR_RGB2GRAY)                # :93, on BGR data
    tok = _check(apps_built, "image", path, "15").split()      # python's cv2 is 4.x: the 15-bit flavour of the kernel
    assert [int(t) for t in tok[:4]] == [96, 130, 3, 8]
    assert (int(tok[4]), int(tok[5])) == _sums(bgr[..., ::-1])
    assert (int(tok[6]), int(tok[7])) == _sums(gray)
    tok = _check(apps_built, "image", path).split()            # default: OpenCV 3, the reference's dependency
    assert (int(tok[6]), int(tok[7])) == _sums(_gray14(bgr[..., ::-1]))
    # 16-bit depth
    path = os.path.join(ds, "depth", "0000.png")
    d = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert d.dtype == np.uint16
    tok = _check(apps_built, "image", path).split()
    assert [int(t) for t in tok[:4]] == [96, 130, 1, 16]
    assert (int(tok[4]), int(tok[5])) == _sums(d)
    # an RGBA frame with every PNG filter type in play (cv2 picks filters adaptively on noisy content)
    rng = np.random.default_rng(3)
    rgba = rng.integers(0, 256, size=(37, 53, 4), dtype=np.uint8)
    rgba[:, :20] = (rgba[:, :20] // 32) * 32
    p2 = str(tmp_path / "rgba.png")
    cv2.imwrite(p2, rgba)
    back = cv2.imread(p2, cv2.IMREAD_UNCHANGED)
    tok = _check(apps_built, "image", p2, "15").split()
    assert [int(t) for t in tok[:4]] == [37, 53, 4, 8]
    assert (int(tok[4]), int(tok[5])) == _sums(back[..., [2, 1, 0, 3]])
    assert (int(tok[6]), int(tok[7])) == _sums(cv2.cvtColor(back, cv2.COLOR_RGBA2GRAY))  # 4-channel flavour of the same kernel


def test_pgm_fallback_config_and_groundtruth(apps_built, tmp_path):
    ds = _gen(tmp_path, "--rows", "64", "--cols", "80", "--pgm", "--id0", "2", "--id1", "5")
    e = np.load(os.path.join(ds, "expected_inputs.npz"))
    tok = _check(apps_built, "image", os.path.join(ds, "rgb", "0002.pgm")).split()
    assert [int(t) for t in tok[:4]] == [64, 80, 1, 8] and (int(tok[4]), int(tok[5])) == _sums(e["im0"])
    tok = _check(apps_built, "image", os.path.join(ds, "depth", "0002.pgm")).split()
    assert [int(t) for t in tok[:4]] == [64, 80, 1, 16] and (int(tok[4]), int(tok[5])) == _sums(e["depth_u16"])
    kv = dict(l.split("=", 1) for l in _check(apps_built, "config", os.path.join(ds, "config.yaml")).strip().splitlines())
    assert kv["image0_id"] == "0002" and kv["image1_id"] == "0005" and kv["dataset"] == "eth_cvg"
    assert kv["use_groundtruth"] == "1" and float(kv["fy"]) < 0 and float(kv["depth_factor_inv"]) == 1.0 / 5000
    rows = [np.array([float(v) for v in l.split()]) for l in _check(apps_built, "gt", os.path.join(ds, "groundtruth.txt")).strip().splitlines()]
    assert len(rows) == 6                                   # pose of frame k on line k
    np.testing.assert_allclose(rows[2], e["T_wc0"], atol=1e-12)
    np.testing.assert_allclose(rows[5], e["T_wc1"], atol=1e-12)


def test_reference_config_file_parses(apps_built):
    """The upstream config_eth_cvg.yaml keys (restated here; the file itself stays in the reference tree)."""
    import tempfile
    txt = ("%YAML:1.0\nimage0_id: '0030'\nimage1_id: '0035'\nimage0_type: rgb\nimage1_type: rgb\n\nuse_groundtruth: '1'\n\n"
           "dataset: eth_cvg\n\nim_address: /data/ethl1_global/\n\ndepth_factor: 5000.0\n\nfx: 481.20\nfy: -480.0\ncx: 319.50\ncy: 239.50\n\n"
           "#make sure you have GPU and CUDA. 1: use GPU, 0: use CPU\nuse_gpu: 1\n")
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        f.write(txt)
    kv = dict(l.split("=", 1) for l in _check(apps_built, "config", f.name).strip().splitlines())
    os.unlink(f.name)
    assert kv["image0_id"] == "0030" and kv["im_address"] == "/data/ethl1_global/" and kv["fx"] == "481.20" and kv["use_gpu"] == "1"


@pytest.mark.gpu
def test_pose_estimation_binary_matches_the_cabi_and_the_oracle(apps_built, tmp_path):
    nid = importlib.import_module("nid-pose-estimation_b200")
    from oracle import binding as orc
    ds = _gen(tmp_path, "--rows", "240", "--cols", "320", "--cell", "4", "--bins", "16")
    r = subprocess.run([os.path.join(apps_built, "NID_pose_estimation"), os.path.join(ds, "config.yaml")], cwd=str(tmp_path),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    for line in ("optimize relative pose between 0 and 1", "use groundtruth pose", "original matrix to be optimized",
                 "the error to be minimized is (6d minimal form)", "enter optimization", "the final error is", "pose optimized"):
        assert line in out
    assert "iteration= 0\t chi2= " in r.stderr and "levenbergIter= " in r.stderr     # g2o's verbose line
    final = np.array([float(v) for v in out.split("the final error is \n")[1].splitlines()[0].split()])
    csv = open(tmp_path / "nid_error.csv").read().strip().split(",")
    assert len(csv) == 8 and csv[6:] == ["0", "1"]
    np.testing.assert_allclose([float(v) for v in csv[:6]], final, rtol=1e-5)
    # the same solve through the Python mirror of the C-ABI and through the oracle
    e = np.load(os.path.join(ds, "expected_inputs.npz"))
    depth = e["depth_u16"].astype(np.float64) * (1.0 / 5000)
    pose0 = orc.reference_perturbation(e["T_wc1"])
    ctx = nid.Context(240, 320, 4, 16)
    ctx.set_pair(0, depth, e["im0"], e["im1"], e["T_wc0"], e["intr"])
    ctx.prepare(0, orc.se3_to_mat16(pose0))
    pose, trace, stats = ctx.solve(0, pose0, 10)
    gt = orc.se3_from_mat16(importlib.import_module("nid-pose-estimation_b200.synth").mat16_inverse(e["T_wc1"]))
    np.testing.assert_allclose(final, (gt - pose)[:6], atol=2e-6)           # printed with 6 significant digits
    P = orc.Problem(e["im0"], depth, e["im1"], e["T_wc0"], e["intr"], 4, 16)
    P.prepare(pose0)
    poseo, its, _, _ = P.optimize(pose0, 10)
    assert its == stats[0]
    np.testing.assert_allclose(final, (gt - poseo)[:6], atol=1e-4)
    assert np.linalg.norm(final[:3]) < np.linalg.norm((gt - pose0)[:3])      # the solve moved towards the true pose


@pytest.mark.gpu
def test_standard_property_binary_and_cost_surface(apps_built, tmp_path):
    from oracle import binding as orc
    synth = importlib.import_module("nid-pose-estimation_b200.synth")
    ds = _gen(tmp_path, "--rows", "240", "--cols", "320", "--cell", "8", "--bins", "8")
    with open(os.path.join(ds, "config.yaml"), "a") as f:
        f.write("sweep: 5\nsweep_range: 0.02\nsweep_csv: %s\n" % (tmp_path / "surf.csv"))
    r = subprocess.run([os.path.join(apps_built, "NID_standard_property"), os.path.join(ds, "config.yaml")], cwd=str(tmp_path),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    total = float(r.stdout.split("final nid is ")[1].split()[0])
    e = np.load(os.path.join(ds, "expected_inputs.npz"))
    depth = e["depth_u16"].astype(np.float64) * (1.0 / 5000)
    gt = synth.mat16_inverse(e["T_wc1"])
    to, _ = orc.hard_nid(e["im0"], depth, e["im1"], e["T_wc0"], gt, e["intr"], 8, 8)
    assert total == pytest.approx(to, rel=1e-5)                               # printed with 6 significant digits
    rows = np.loadtxt(tmp_path / "surf.csv", delimiter=",", skiprows=1)
    assert rows.shape == (6 * 25, 6)
    pose_gt = orc.se3_from_mat16(gt)
    for row in rows[::17]:
        a = int(row[0])
        xi = np.zeros(6)
        xi[a] += row[3]
        xi[(a + 1) % 6] += row[4]
        M = orc.se3_to_mat16(orc.se3_mul(orc.se3_exp(xi), pose_gt))
        t2, _ = orc.hard_nid(e["im0"], depth, e["im1"], e["T_wc0"], M, e["intr"], 8, 8)
        assert row[5] == pytest.approx(t2, rel=1e-5)
    centre = rows[(rows[:, 1] == 2) & (rows[:, 2] == 2)][:, 5]
    assert np.allclose(centre, total, rtol=1e-5) and np.all(centre <= rows[:, 5].max())
