#!/usr/bin/env python
"""Key metrics per kernel launch from an .ncu-rep (via `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_mio_throttle',
        'smsp__pcsamp_warps_issue_stalled_tex_throttle', 'smsp__pcsamp_warps_issue_stalled_lg_throttle', 'smsp__pcsamp_warps_issue_stalled_no_instructions',
        'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_dispatch_stall',
        'smsp__pcsamp_warps_issue_stalled_branch_resolving', 'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_imc_miss',
        'local_load/store: smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
sel = sys.argv[2:] 
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if sel and not any(s in name for s in sel):
        continue
    print('-----')
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f'{k:72s} {r[i]:>22s} {units[i]}')
