#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline metrics + per-region instruction counts.
usage: ncu_summary.py <report.ncu-rep> <units-per-launch (e.g. warp datapoint-steps)>"""
import csv, subprocess, sys, io
rep = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ',
        'launch__registers_per_thread ', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum ', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg.per_second']
for h, u, v in zip(hdr, units, vals):
    if any(w in h + ' ' for w in want):
        try:
            if float(v.replace(',', '')) < 0.005:
                continue
        except ValueError:
            pass
        print(f"{h:86s} {v:>22s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_i = sum(int(r[ix['Instructions Executed']]) for r in data)
tot_s = max(1, sum(int(r[ix['# Samples']]) for r in data))
base = int(data[0][0], 16)
runs = []
for r in data:
    a = int(r[0], 16) - base
    c = int(r[ix['Instructions Executed']]) / steps
    sm = int(r[ix['# Samples']])
    key = round(c, 2)
    if runs and abs(runs[-1][1] - key) < 0.02:
        runs[-1][2] += 1
        runs[-1][3] += sm
    else:
        runs.append([a, key, 1, sm, r[1].strip()[:50]])
print("\nregions (offset, executions per unit, #instrs, instr/unit, stall-sample share):")
for a, c, n, sm, s in runs:
    if c * n > 1.0 or sm * 100 / tot_s > 1.0:
        print(f"{a:6x} x{c:7.2f}  {n:4d} -> {c*n:7.1f}  {sm*100/tot_s:5.1f}%  {s}")
print("total instr per unit", tot_i / steps)
