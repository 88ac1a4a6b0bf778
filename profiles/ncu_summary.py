#!/usr/bin/env python
"""Prints the handful of ncu metrics the design notes quote, from a .ncu-rep (run here, no GPU needed)."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum',
        'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_global_ld.sum']
STALL = 'smsp__average_warps_issue_stalled_'
rows = list(csv.reader(subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:110])
    for w in WANT:
        if w in hdr:
            print(f'  {w} = {r[hdr.index(w)]} {units[hdr.index(w)]}')
    st = sorted(((float(r[i]), hdr[i][len(STALL):-len('_per_issue_active.ratio')]) for i in range(len(hdr))
                 if hdr[i].startswith(STALL) and hdr[i].endswith('_per_issue_active.ratio') and r[i]), reverse=True)
    print('  stalls (warps per issue):', ', '.join(f'{n}={v:.2f}' for v, n in st[:7]))
