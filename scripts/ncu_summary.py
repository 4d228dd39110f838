#!/usr/bin/env python
"""Prints the metrics we read from an ncu report (ncu -i X.ncu-rep --page raw --csv) side by side."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu.sum', 'smsp__inst_executed_op_shared_ld.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_registers', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg', 'launch__grid_size', 'launch__waves_per_multiprocessor',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'sm__inst_executed.sum']
cols = {}
for path in sys.argv[1:]:
    if path.endswith('.csv'):   # already exported on the GPU box (ncu -i X.ncu-rep --page raw --csv)
        out = open(path).read()
    else:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    cols[path] = d
print(f"{'metric':64s}", *[f"{p.split('/')[-1].replace('.ncu-rep','').replace('.raw.csv','')[:21]:>22s}" for p in sys.argv[1:]])
keys = [k for k in WANT if any(k in d for d in cols.values())]
stall = sorted({h for d in cols.values() for h in d if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')})
for k in keys + stall:
    short = k.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')
    print(f"{short[:64]:64s}", *[f"{(cols[p].get(k, ('-', ''))[0][:12] + ' ' + cols[p].get(k, ('', ''))[1][:7]):>22s}" for p in sys.argv[1:]])
