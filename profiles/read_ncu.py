#!/usr/bin/env python3
"""Print selected metrics of every launch in an .ncu-rep (needs `ncu` on PATH, no GPU).
usage: python profiles/read_ncu.py report.ncu-rep [extra-metric-prefix ...]"""
import csv
import subprocess
import sys

KEYS = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'smsp__average_warps_issue_stalled', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = KEYS + sys.argv[2:]
for r in rows[2:]:
    print('---')
    for i, h in enumerate(hdr):
        if any(h == k or h.startswith(k) for k in keys):
            v = r[i]
            if h.startswith('smsp__average_warps_issue_stalled') and (v in ('', '0') or float(v.replace(',', '')) < 0.3):
                continue
            print('  %-92s %s %s' % (h, v, units[i]))
