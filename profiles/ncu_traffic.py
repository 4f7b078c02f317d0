#!/usr/bin/env python3
"""DRAM traffic of the tcgen05 conv launches of ONE iteration from an `ncu --set full` report.
usage: python profiles/ncu_traffic.py gpurun_out/prof.ncu-rep [offset] > profiles/<round>_tcconv_traffic.json

The capture must hold (at least) the tensor-core convolution launches of one whole L-BFGS iteration, which come in a
fixed cyclic order (see CYCLE below); it may start anywhere.
bench.py reads the JSON to fill roofline.traffic (bytes per 3x3 launch, averaged over the 24 of conv1_2..conv5_1)."""
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
    launches.append({'kernel': r[col('Kernel Name')].split('(')[0].split('::')[-1],
                     'dram_read': to_bytes(r[ir], units[ir]), 'dram_write': to_bytes(r[iw], units[iw]),
                     'us_under_ncu': to_us(r[it], units[it]),
                     'tensor_pipe_pct': float(r[itp]) if itp is not None and r[itp] else None})
# Round 2 (final kernels): conv1_1's style gradient is folded into its data-gradient weights and conv2_1's is contracted
# inside conv2_2's data-gradient kernel, so an iteration has 28 tensor-core conv launches: 12 forward 3x3
# (conv1_2 .. conv5_1), 3 style-gradient 1x1 contractions (conv3_1, conv4_1, conv5_1), 12 data-gradient 3x3
# (conv5_1 .. conv1_2) and the dual-source conv1_1 data gradient (tc_conv_first_stencil_kernel), which ends the cycle.
# The capture may start anywhere: the cycle is located by that last kernel.
CYCLE = 28
END = 'tc_conv_first_stencil_kernel'
ends = [i for i, l in enumerate(launches) if END in l['kernel']]
start = next((e + 1 for e in ends if e + 1 + CYCLE <= len(launches)), None)
assert start is not None, 'no complete iteration (28 launches after a %s) in the capture' % END
launches = launches[start:start + CYCLE]
assert END in launches[-1]['kernel'], 'cycle misaligned'
NAMES = ['conv1_2', 'conv2_1', 'conv2_2', 'conv3_1', 'conv3_2', 'conv3_3', 'conv3_4', 'conv4_1', 'conv4_2', 'conv4_3',
         'conv4_4', 'conv5_1']
conv, style, first = [], [], []
for pos, l in enumerate(launches):
    if pos < 12:
        l['what'] = NAMES[pos] + ' fwd'; conv.append(l)
    elif pos < 15:
        l['what'] = 'style grad (%s)' % ['conv3_1', 'conv4_1', 'conv5_1'][pos - 12]; style.append(l)
    elif pos < 27:
        l['what'] = NAMES[26 - pos] + ' dgrad' + (' + conv2_1 style gradient' if NAMES[26 - pos] == 'conv2_2' else ''); conv.append(l)
    else:
        l['what'] = 'conv1_1 dgrad + folded style gradient (stencil form, dual source)'; first.append(l)
assert len(conv) == 24
tot = sum(l['dram_read'] + l['dram_write'] for l in conv)
print(json.dumps({
    'source': sys.argv[1], 'what': 'ncu --set full, one L-BFGS iteration at 1024x1024, tensor-core conv launches',
    'conv3x3_launches': 24, 'conv3x3_dram_bytes_per_iteration': tot, 'conv3x3_dram_bytes_per_launch': tot / 24,
    'conv3x3_us_under_ncu': sum(l['us_under_ncu'] for l in conv),
    'conv3x3_tensor_pipe_pct_time_weighted': sum((l['tensor_pipe_pct'] or 0) * l['us_under_ncu'] for l in conv) /
    sum(l['us_under_ncu'] for l in conv),
    'style_grad_dram_bytes_per_iteration': sum(l['dram_read'] + l['dram_write'] for l in style),
    'conv1_1_dgrad_dram_bytes': sum(l['dram_read'] + l['dram_write'] for l in first),
    'launches': launches}, indent=1))
