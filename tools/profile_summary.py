#!/usr/bin/env python
"""Turns one GPU visit of tools/gpu_round.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/:
  profiles/<tag>_launches.txt      kernel launch list of `bench.py --pairs 24 --steps 3 --warmup 3` (ncu gpu__time_duration)
  profiles/<tag>_ncu_full.txt      key `ncu --set full` metrics of the evaluation kernels
  profiles/<tag>_bench.json        the bench lines of the same visit
  profiles/traffic.json            measured DRAM bytes per evaluation of every evaluation kernel (read by bench.py)
usage: python tools/profile_summary.py <tag> [pairs]"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 24
go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
EVAL = ("k_hist_sell", "k_assemble", "k_jac_sell", "k_jac_final_sorted")


def short(name):
    n = name.replace("void ", "").replace("nid::", "")
    return n.split("<")[0].split("(")[0]


rows = [r for r in csv.reader(open(os.path.join(go, f"{tag}_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    agg.setdefault((short(r[4]), r[8], r[7]), []).append(float(r[-1]))
step = {k: sum(v) / len(v) for k, v in agg.items() if k[0] in EVAL}
tot = sum(step.values())
with open(os.path.join(pr, f"{tag}_launches.txt"), "w") as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --pairs {pairs} --steps 3 --warmup 3 --cpu-budget 0.2 --solves 0 --c4-pairs 0 --c5 0 --old-gpu 0   (tools/gpu_prof.sh)\n")
    f.write("per-launch times are cold-cache and serialised under ncu: the SHARE of the step is what bench.py's live CUDA-event timing must agree with\n")
    f.write(f"{'kernel':24s} {'grid':16s} {'block':14s} {'launches':>8s} {'avg us':>10s} {'share of one evaluation step':>30s}\n")
    for k, v in agg.items():
        a = sum(v) / len(v)
        sh = f"{100 * a / tot:5.1f} %" if k in step else ("(HBM probe, bench only)" if k[0] == "k_warp_sample_jobs" else "(setup, once per pair)")
        f.write(f"{k[0]:24s} {k[1]:16s} {k[2]:14s} {len(v):8d} {a / 1e3:10.1f} {sh:>30s}\n")
    f.write(f"one evaluation step ({pairs} cost+Jacobian evaluations): {tot / 1e3:.1f} us serialised under ncu\n")

rep = os.path.join(go, f"{tag}_prof.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(out.splitlines()))
hdr, units = rr[0], rr[1]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_op_shared_atom.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle',
        'smsp__pcsamp_warps_issue_stalled_mio_throttle', 'smsp__pcsamp_warps_issue_stalled_lg_throttle',
        'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_warps_issue_stalled_selected',
        'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_branch_resolving']
traffic = {}
pipes = {}
with open(os.path.join(pr, f"{tag}_ncu_full.txt"), "w") as f:
    f.write(f"ncu --set full --clock-control none --import-source on -k regex:'k_hist_sell|k_jac_sell|k_assemble|k_jac_final' -s 8 -c 4 python bench.py --pairs {pairs} --steps 3 --warmup 3 ...   (tools/gpu_prof.sh)\n")
    f.write(f"one launch = {pairs} cost+Jacobian evaluations of 640x480 pairs (4x4 cells, 16 bins)\n")
    for r in rr[2:]:
        name = short(r[hdr.index('Kernel Name')])
        f.write(f"----- {r[hdr.index('Kernel Name')]}\n")
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k:72s} {r[i]:>22s} {units[i]}\n")

        def val(k):
            i = hdr.index(k)
            v = float(r[i].replace(",", ""))
            u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        traffic[name] = (val('dram__bytes_read.sum') + val('dram__bytes_write.sum')) / pairs
        pipes[name] = {"fp64_pipe_pct": val('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),
                       "issue_active_pct": val('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                       "warps_active_pct": val('sm__warps_active.avg.pct_of_peak_sustained_active'),
                       "registers_per_thread": val('launch__registers_per_thread')}
# fp64 warp instructions per evaluation: executed counts of the D* opcodes on the source page of the same capture
fp64 = {}
for name in EVAL:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{name}"], capture_output=True, text=True).stdout
    rows_s = list(csv.reader(out.splitlines()))
    try:
        h = next(r for r in rows_s if r and r[0] == "Address")
    except StopIteration:
        continue
    data = [r for r in rows_s if len(r) == len(h) and r[0] != "Address"]
    first = data[0][0]
    for i in range(1, len(data)):
        if data[i][0] == first:
            data = data[:i]  # the dump concatenates the captured launches: keep the first
            break
    iS, iI = h.index("Source"), h.index("Instructions Executed")
    tot = 0
    for r in data:
        txt = r[iS].strip()
        op = (txt.split()[1] if txt.startswith("@") else txt.split()[0]) if txt else ""
        if op[:1] == "D" and op.split(".")[0] in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"):
            tot += int(r[iI])
    fp64[name] = tot / pairs
json.dump({"source": f"profiles/{tag}_ncu_full.txt (dram__bytes_read.sum + dram__bytes_write.sum per launch / {pairs} evaluations; fp64 = executed DFMA/DADD/DMUL/DSETP warp instructions, ncu source page)",
           "dram_bytes_per_eval": traffic, "pipes": pipes, "fp64_warp_inst_per_eval": fp64}, open(os.path.join(pr, "traffic.json"), "w"), indent=1)
lines = []
for fn in (f"{tag}_bench_ref.json", f"{tag}_bench.json"):
    p = os.path.join(go, fn)
    if os.path.exists(p):
        lines += [l for l in open(p).read().splitlines() if l.startswith("{")]
open(os.path.join(pr, f"{tag}_bench.json"), "w").write("\n".join(lines) + "\n")
print(open(os.path.join(pr, f"{tag}_launches.txt")).read())
print(json.dumps(traffic, indent=1))
