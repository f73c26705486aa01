#!/usr/bin/env python
"""Print the headline metrics of every kernel launch in an ncu report (raw page): usage ncu_summary.py REPORT.ncu-rep"""
import csv, io, subprocess, sys
WANT = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fmaheavy.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct', 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct',
        'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct', 'smsp__warp_issue_stalled_membar_per_warp_active.pct',
        'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'sm__sass_thread_inst_executed_op_fadd_pred_on.sum', 'sm__sass_thread_inst_executed_op_ffma_pred_on.sum', 'sm__sass_thread_inst_executed_op_fmul_pred_on.sum',
        'launch__shared_mem_per_block_dynamic', 'sm__maximum_warps_per_active_cycle_pct', 'local_load', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("-----", r[idx['Kernel Name']][:100])
    for w in WANT:
        if w in idx:
            print("  %-75s %s %s" % (w, r[idx[w]], units[idx[w]]))
