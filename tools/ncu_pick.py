#!/usr/bin/env python3
"""Pick metrics out of `ncu -i X.ncu-rep --page raw --csv` output. usage: ncu_pick.py raw.csv [substr ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = sys.argv[2:] or ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'sm__throughput.avg.pct', 'gpu__dram_throughput.avg.pct', 'sm__warps_active.avg.pct',
    'launch__registers_per_thread', 'launch__occupancy_limit', 'launch__grid_size', 'launch__waves',
    'sm__inst_executed.sum', 'smsp__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct',
    'sm__cycles_elapsed.max', 'smsp__warps_eligible.avg.per_cycle_active', 'pipe_fma', 'pipe_xu', 'pipe_alu', 'pipe_lsu',
    'smsp__average_warp', 'sm__cycles_active.avg', 'l1tex__data_bank_conflicts', 'smsp__pcsamp_warps_issue_stalled']
for i, h in enumerate(hdr):
    if any(w in h for w in want):
        print('%-95s %-14s %s' % (h[:95], units[i], [r[i] for r in rows[2:]]))
