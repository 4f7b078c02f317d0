#!/usr/bin/env python3
"""DRAM traffic of the tcgen05 conv launches of ONE iteration from an `ncu --set full` report.
usage: python profiles/ncu_traffic.py gpurun_out/prof.ncu-rep > profiles/<round>_tcconv_traffic.json

The capture must hold the 29 tc_conv_kernel launches of one L-BFGS iteration in order: 12 forward 3x3
convolutions (conv1_2 .. conv5_1), 5 style-gradient 1x1 contractions, 12 data-gradient convolutions.
bench.py reads the JSON to fill roofline.traffic (bytes per 3x3 launch, averaged over the 24)."""
import csv
import json
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]


def col(name):
    return hdr.index(name)


def to_bytes(v, u):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]


def to_us(v, u):
    v = float(v.replace(',', ''))
    return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}[u]


ir, iw, it = col('dram__bytes_read.sum'), col('dram__bytes_write.sum'), col('gpu__time_duration.sum')
itp = col('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') if 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active' in hdr else None
launches = []
for r in rows[2:]:
    launches.append({'kernel': r[col('Kernel Name')].split('(')[0][-24:],
                     'dram_read': to_bytes(r[ir], units[ir]), 'dram_write': to_bytes(r[iw], units[iw]),
                     'us_under_ncu': to_us(r[it], units[it]),
                     'tensor_pipe_pct': float(r[itp]) if itp is not None and r[itp] else None})
assert len(launches) == 29, 'expected 29 tc_conv launches of one iteration, got %d' % len(launches)
conv = launches[:12] + launches[17:]
style = launches[12:17]
tot = sum(l['dram_read'] + l['dram_write'] for l in conv)
print(json.dumps({
    'source': sys.argv[1], 'what': 'ncu --set full, one L-BFGS iteration at 1024x1024, tc_conv_kernel launches',
    'conv3x3_launches': 24, 'conv3x3_dram_bytes_per_iteration': tot, 'conv3x3_dram_bytes_per_launch': tot / 24,
    'conv3x3_tensor_pipe_pct_time_weighted': sum((l['tensor_pipe_pct'] or 0) * l['us_under_ncu'] for l in conv) /
    sum(l['us_under_ncu'] for l in conv),
    'style_grad_dram_bytes_per_iteration': sum(l['dram_read'] + l['dram_write'] for l in style),
    'launches': launches}, indent=1))
