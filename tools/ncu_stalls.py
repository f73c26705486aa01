#!/usr/bin/env python
"""Key throughput / stall numbers of the first kernel in an ncu report: ncu_stalls.py REPORT"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, v = rows[0], rows[-1]
d = dict(zip(h, v))
for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]:
    if k in d:
        print("%-80s %s" % (k, d[k]))
res = []
for k, x in d.items():
    if "smsp__pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
        try:
            res.append((float(x), k))
        except ValueError:
            pass
tot = sum(r[0] for r in res) or 1
for x, k in sorted(res, reverse=True)[:8]:
    print("%5.1f%% %s" % (100 * x / tot, k.replace("smsp__pcsamp_warps_issue_stalled_", "stall ")))
