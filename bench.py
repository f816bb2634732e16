#!/usr/bin/env python
"""bench.py — NID cost+Jacobian evals/s @640x480 (BASELINE.json metric), one rank per GPU.

A "step" evaluates the cost + 6-DoF Jacobian (the a9-equivalent, want_jac=1) of every frame pair this rank
owns once, at a pose that changes every step. Workload = BASELINE config[1]: 640x480, 4x4 cells, 16-bin
B-spline NID. The rank owns `--pairs` pair slots (default 96) whose device footprint (> 800 MB) exceeds the 126 MB L2,
so every step streams its inputs from HBM ("inputs larger than L2").

  value : evals/s with inputs resident in HBM (poses staged, results left on the device), CUDA events on
          the library's stream, max over ranks.
  e2e   : the same evaluations through the reference-facing C-ABI call with HOST buffers
          (nid_eval_jobs: H2D poses from pinned memory, kernels, D2H Htarget/Hjoint/der, sync).
  --impl reference : the reference's CPU implementation of the path (the fp64 restatement in oracle/,
          OpenMP over cells on all host cores) on the same workload, bounded sample per step.

Multi-GPU: whole problems are sharded across ranks (weak scaling, no data-path collective); NCCL is only
used for the barrier and the max-over-ranks of the device time.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, COLS, CELL, BINS = 480, 640, 4, 16
ALGO_BYTES_PER_EVAL = ROWS * COLS * 12 + CELL * CELL * 8 * 8  # fp32 depth + I_ref + I_tgt planes, outputs
DISTINCT_PAIRS = 6


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                               "-i", str(self.gpu)], timeout=5).decode().strip()
                self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def make_poses(orc, pose0_list, step, n_slots):
    """pose of slot s at step k: small seeded se(3) offset on the left of that pair's initial guess."""
    rng = np.random.default_rng(77 + step)
    xi = rng.uniform(-1, 1, size=(n_slots, 6)) * np.array([2e-3, 2e-3, 2e-3, 5e-3, 5e-3, 5e-3])
    return np.stack([orc_pose_to_mat(orc, orc.se3_mul(orc.se3_exp(xi[s]), pose0_list[s % len(pose0_list)])) for s in range(n_slots)])


def orc_pose_to_mat(orc, pose7):
    return orc.se3_to_mat16(pose7)


def peaks():
    """HBM copy peak in GB/s: the driver-written MEASURED_PEAKS.json when present (sustained figure preferred: the
    kernel is timed inside a long step), else the fallback B200_PROFILING.md states."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            flat = {}

            def walk(prefix, o):
                if isinstance(o, dict):
                    for k, v in o.items():
                        walk(f"{prefix}.{k}" if prefix else str(k), v)
                elif isinstance(o, (int, float)):
                    flat[prefix.lower()] = float(o)
            walk("", d)
            cands = [(k, v) for k, v in flat.items() if "hbm" in k and v > 100]
            for pref in ("sustain", "burst", ""):
                for k, v in cands:
                    if pref in k:
                        return v, f"measured (MEASURED_PEAKS.json {k})"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(kernel, n_evals):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return d["dram_bytes_per_eval"][kernel] * n_evals, d["source"]
    except Exception:
        return None, None


def measured_pipes(kernel):
    """fp64-pipe / issue utilisation of `kernel` from the committed ncu --set full capture (profiles/traffic.json), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return dict(d["pipes"][kernel], source=d["source"].split(" ")[0])
    except Exception:
        return None


def cpu_layout(cores):
    """Use every host core: `workers` independent evaluations in flight, each an OpenMP team of `team` threads over
    the 16 cells (team divides 16 so that cells split evenly)."""
    best = (1, 1)
    for team in (1, 2, 4, 8, 16):
        workers = max(1, cores // team)
        if team <= cores and workers * team >= best[0] * best[1]:
            best = (workers, team)
    return best


class CpuArm:
    """The reference's CPU implementation of the path (oracle/ restatement: the reference's own CPU code needs
    Eigen/OpenCV, absent here) on the bench workload; `workers` problems evaluated concurrently (ctypes releases
    the GIL), `team` OpenMP threads each."""

    def __init__(self, orc, synth, workers, team):
        self.orc, self.workers, self.team = orc, workers, team
        p = synth.make_pair(1000, ROWS, COLS)
        self.pose0 = orc.reference_perturbation(p.T_wc1)
        self.P = [orc.Problem(p.im0, p.depth0, p.im1, p.T_wc0, p.intr, CELL, BINS, threads=team) for _ in range(workers)]
        for P in self.P:
            P.prepare(self.pose0)
            P.eval(self.pose0, True)  # warm

    def run(self, evals_per_worker, seed):
        """evals_per_worker cost+Jacobian evaluations on every worker; returns (evals, seconds)."""
        orc = self.orc

        def work(w):
            rng = np.random.default_rng(seed * 131 + w)
            for _ in range(evals_per_worker):
                xi = rng.uniform(-1, 1, size=6) * 2e-3
                self.P[w].eval(orc.se3_mul(orc.se3_exp(xi), self.pose0), True)
        ts = [threading.Thread(target=work, args=(w,)) for w in range(self.workers)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return evals_per_worker * self.workers, time.perf_counter() - t0


def cpu_eval_rate(orc, synth, workers, team, budget_s):
    """evals/s of the CPU arm, bounded sample of about budget_s seconds."""
    arm = CpuArm(orc, synth, workers, team)
    n, el = arm.run(1, 0)
    per = max(1, min(200, int(budget_s / max(el, 1e-3))))
    n, el = arm.run(per, 1)
    return n / el, n, el


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as orc
    synth = importlib.import_module("nid-pose-estimation_b200.synth")
    cores = os.cpu_count() or 1
    workers, team = cpu_layout(cores)
    arm = CpuArm(orc, synth, workers, team)
    evals_per_worker = 2  # bounded sample of the workload per step
    for k in range(args.warmup):
        arm.run(evals_per_worker, k)
    n, el = 0, 0.0
    for k in range(args.steps):
        a, b = arm.run(evals_per_worker, 100 + k)
        n += a
        el += b
    v = n / el
    line = {
        "impl": "reference", "metric": "NID cost+Jacobian evals/s @640x480", "value": v, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: 640x480 pair, 4x4 cells, 16-bin cubic B-spline NID, cost+Jacobian", "rows": ROWS,
                   "cols": COLS, "cell": CELL, "bins": BINS},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": workers * team, "kind": "port",
                         "sample": f"{evals_per_worker * workers} evals/step of one seeded 640x480 pair: {workers} evaluations in "
                                   f"flight x {team} OpenMP threads over the 16 cells (the reference CPU path cannot be "
                                   "compiled here: Eigen/OpenCV absent)"},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=96, help="pair slots per rank (96 x 8.3 MB of inputs >> L2)")
    ap.add_argument("--solves", type=int, default=1, help="also time complete LM pose solves of every slot (0: skip)")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the NID path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    nid = importlib.import_module("nid-pose-estimation_b200")
    synth = importlib.import_module("nid-pose-estimation_b200.synth")
    from oracle import binding as orc  # pose algebra for the synthetic workload + the cpu_baseline leg only

    n_slots = args.pairs
    ctx = nid.Context(ROWS, COLS, CELL, BINS, n_pairs=n_slots, max_jobs=n_slots, device=local_rank)
    for kv in filter(None, os.environ.get("NID_OPTS", "").split(",")):  # developer knob, e.g. NID_OPTS=use_tex=0
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    pose0 = []
    distinct = [synth.make_pair(1000 + rank * DISTINCT_PAIRS + i, ROWS, COLS) for i in range(DISTINCT_PAIRS)]
    for i, p in enumerate(distinct):
        pose0.append(orc.reference_perturbation(p.T_wc1))
    t_prep = 0.0
    for s in range(n_slots):
        if s == min(DISTINCT_PAIRS, n_slots - 1):  # the first pairs pay the one-time allocations (pixel store, pinned staging): not timed
            ctx.sync()
            t_prep = time.perf_counter()
        p = distinct[s % DISTINCT_PAIRS]
        ctx.set_pair(s, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
        ctx.prepare(s, orc.se3_to_mat16(pose0[s % DISTINCT_PAIRS]))
    ctx.sync()
    t_prep = time.perf_counter() - t_prep
    n_prep = n_slots - min(DISTINCT_PAIRS, n_slots - 1)
    job_pair = np.arange(n_slots, dtype=np.int32)
    total_steps = args.warmup + args.steps
    poses = [make_poses(orc, pose0, k, n_slots) for k in range(total_steps)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    for k in range(args.warmup):
        ctx.stage_jobs(poses[k], job_pair)
        ctx.eval_staged(n_slots, True)
    ctx.sync()
    launches0 = ctx.launch_count()
    barrier()
    with ClockSampler(local_rank) as clk:
        ctx.event_record(0)
        for k in range(args.warmup, total_steps):
            ctx.stage_jobs(poses[k], job_pair)
            ctx.eval_staged(n_slots, True)
        ctx.event_record(1)
        ms = ctx.event_elapsed_ms()
        barrier()
    launches = ctx.launch_count() - launches0
    Ht, Hj, der = ctx.fetch_results(n_slots, True)
    assert np.all(np.isfinite(Hj)) and np.all(np.isfinite(der)), "non-finite results in the timed region"

    # ---------------- end to end through the host-buffer C-ABI call (`e2e`)
    for k in range(args.warmup):
        ctx.eval_jobs(poses[k], job_pair, True)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.warmup, total_steps):
        ctx.eval_jobs(poses[k], job_pair, True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # ---------------- per-kernel device time, live (roofline of the dominant kernel)
    ctx.set_option("time_kernels", 1)
    for k in range(args.warmup, total_steps):
        ctx.stage_jobs(poses[k], job_pair)
        ctx.eval_staged(n_slots, True)
    ctx.sync()
    kt = ctx.kernel_times()
    ctx.set_option("time_kernels", 0)

    # ---------------- kernel 1 on its own (warp + sample, north_star kernel (1)): HBM roofline probe
    probe = None
    try:
        for k in range(min(args.warmup, 3)):
            ctx.warp_sample_jobs(poses[k], job_pair, fetch=False)
        ctx.event_record(0)
        for k in range(args.warmup, total_steps):
            ctx.warp_sample_jobs(poses[k], job_pair, fetch=False)
        ctx.event_record(1)
        probe_ms = ctx.event_elapsed_ms() / args.steps
        probe = {"kernel": "k_warp_sample_jobs", "ms_per_launch": probe_ms,
                 "algorithmic_bytes_per_launch": ROWS * COLS * 24 * n_slots,
                 "moved_bytes_per_launch": ROWS * COLS * (8 + 4 + 16) * n_slots,
                 "note": "one launch = kernel 1 for every pair slot: fp64 depth in, packed 2x2 gather, float4 {I_c, g_x, g_y, "
                         "valid} out; algorithmic bytes = SURVEY 8(d) 24 B/px; the evaluation path fuses this front end "
                         "into both passes instead of storing its output"}
    except Exception as e:  # noqa: BLE001 - the probe must not take the bench line down
        probe = {"kernel": "k_warp_sample_jobs", "error": str(e)}

    # ---------------- complete LM pose solves (optimize(10), reference perturbation), all slots in lockstep
    solve_s, solve_stats = 0.0, None
    if args.solves:
        p7 = np.stack([pose0[s % DISTINCT_PAIRS] for s in range(n_slots)])
        ctx.solve_jobs(p7, job_pair)  # warm
        barrier()
        t0 = time.perf_counter()
        _, solve_stats = ctx.solve_jobs(p7, job_pair)
        torch.cuda.synchronize()
        solve_s = time.perf_counter() - t0

    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3, solve_s * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, solve_ms = t.tolist()
    else:
        e2e_ms, solve_ms = e2e_s * 1e3, solve_s * 1e3

    if rank == 0:
        evals = args.steps * n_slots * world
        value = evals / (ms * 1e-3)
        e2e_value = evals / (e2e_ms * 1e-3)
        peak, peak_src = peaks()
        dom = max(("k_hist_sell", "k_jac_sell"), key=lambda n: kt[n][0])
        dom_ms = kt[dom][0] / max(kt[dom][1], 1)
        share = {n: kt[n][0] for n in kt}
        tot = sum(share.values()) or 1.0
        achieved = ALGO_BYTES_PER_EVAL * n_slots / (dom_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(dom, n_slots)
        cpu_v, cpu_n, cpu_el = cpu_eval_rate(orc, synth, 1, 1, args.cpu_budget / 2)
        workers, team = cpu_layout(os.cpu_count() or 1)
        cores = workers * team
        cpu_vm, cpu_nm, cpu_elm = cpu_eval_rate(orc, synth, workers, team, args.cpu_budget / 2)
        line = {
            "metric": "NID cost+Jacobian evals/s @640x480", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: 640x480 pair, 4x4 cells, 16-bin cubic B-spline NID, cost+Jacobian",
                       "rows": ROWS, "cols": COLS, "cell": CELL, "bins": BINS, "pairs_per_gpu": n_slots,
                       "evals_per_step": n_slots * world,
                       "l2": f"inputs larger than L2: {n_slots} pair slots x 8.3 MB per GPU streamed every step"},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": n_slots * (16 * 8 + 4),
                    "d2h_bytes_per_step": n_slots * CELL * CELL * 8 * 8,
                    "call": "nid_eval_jobs (host poses in, host Htarget/Hjoint/der out, blocking)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "note": "fp64 path: the binding roofs are fp64 issue and latency, not HBM (DESIGN.md 4)",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_EVAL * n_slots,
                         "kernel_ms_per_launch": dom_ms,
                         "ncu_pipes": measured_pipes(dom),
                         "kernel_share_of_step": {n: share[n] / tot for n in share}},
            "cpu_baseline": {"value": cpu_vm, "unit": "evals/s", "cores": cores, "kind": "port",
                             "single_thread_value": cpu_v,
                             "sample": f"{cpu_nm} evals in {cpu_elm:.1f}s on {cores} threads ({workers} evaluations in flight x "
                                       f"{team} OpenMP threads over the cells) and {cpu_n} evals in {cpu_el:.1f}s on 1 thread, one "
                                       "seeded 640x480 pair (oracle/ restatement; the reference CPU path needs Eigen/OpenCV, "
                                       "absent here)"},
            "clocks": clk.summary(),
        }
        if probe and "ms_per_launch" in probe:
            probe["achieved"] = probe["algorithmic_bytes_per_launch"] / (probe["ms_per_launch"] * 1e-3) / 1e9
            probe["moved"] = probe["moved_bytes_per_launch"] / (probe["ms_per_launch"] * 1e-3) / 1e9
            probe["unit"] = "GB/s"
            probe["peak"] = peak
            probe["frac"] = probe["achieved"] / peak
            probe["frac_moved"] = probe["moved"] / peak
        line["roofline"]["warp_sample_probe"] = probe
        line["pair_setup"] = {"value": n_prep / t_prep, "unit": "pairs/s", "pairs": n_prep,
                              "call": "nid_set_pair + nid_prepare per pair (H2D of depth and both images, points, reference "
                                      "spline data, H_ref, regrouped pixel store), wall clock on the host (pageable "
                                      "source buffers: varies with host load), rank 0"}
        if args.solves and solve_ms > 0:
            line["pose_solves"] = {"value": n_slots * world / (solve_ms * 1e-3), "unit": "solves/s",
                                   "solves": n_slots * world, "ms": solve_ms,
                                   "mean_outer_iters": float(solve_stats[:, 0].mean()),
                                   "mean_jac_evals": float(solve_stats[:, 1].mean()),
                                   "mean_cost_evals": float(solve_stats[:, 2].mean()),
                                   "call": "nid_solve_jobs: optimize(10) LM schedule, 6x6 solves on the host overlapped with the other half-batch's kernels, wall clock"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
