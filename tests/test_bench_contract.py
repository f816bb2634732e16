"""bench.py's reference arm runs without a GPU (it times the CPU restatement of the path) and prints the contract's line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                                  timeout=600).decode().strip().splitlines()
    assert len(out) == 1
    d = json.loads(out[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "NID cost+Jacobian evals/s @640x480" and d["config"]["workload"].startswith("C2")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_non_zero_ranks_of_the_reference_arm_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env, timeout=120)
    assert out.decode().strip() == ""


def test_roofline_helpers_read_the_committed_ncu_capture():
    """bench.py takes roofline.traffic and roofline.ncu_pipes from profiles/traffic.json (written by
    tools/profile_summary.py from the committed ncu --set full capture): the file must name the two pixel kernels."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for k in ("k_hist_sell", "k_jac_sell"):
        t, src = bench.measured_traffic(k, 96)
        assert t is not None and t > 96 * 3.6e6 and src.startswith("profiles/")  # at least the algorithmic bytes
        p = bench.measured_pipes(k)
        assert p is not None and 0 < p["fp64_pipe_pct"] < 100 and 0 < p["issue_active_pct"] <= 100
    peak, src = bench.peaks()
    assert 1000 < peak < 10000 and src
