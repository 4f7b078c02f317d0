#!/usr/bin/env python3
"""Per-layer timing of the conv stack (CUDA events, back-to-back launches of one layer) with the
algorithmic TFLOP/s each reaches; optional --flags runs the kernel with parts switched off
(st2_debug_flags) to see what bounds it.  usage: python tools/bench_layers.py [--size 1024] [--flags 0,1,2,4,8]"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument('--size', type=int, default=1024)
ap.add_argument('--height', type=int, default=0, help='canvas rows (default: --size); e.g. 512 with --size 4096 = the shape of one of 8 row strips')
ap.add_argument('--flags', default='0')
ap.add_argument('--reps', type=int, default=20)
ap.add_argument('--precision', default='fp16')
ap.add_argument('--layers', default='')
args = ap.parse_args()
from style_transfer2_b200 import vgg
from style_transfer2_b200.model import B200Model

m = B200Model(precision=args.precision)
H = args.height or args.size
plan = m.plan(H, args.size)
x = torch.randn(1, 3, H, args.size, device=m.engine.device) * 50
plan.forward(x, vgg.BLOB_INDEX['conv5_1'])
rows = []
for flags in [int(f) for f in args.flags.split(',')]:
    m.engine.call('st2_debug_flags', flags)
    for b in range(2, vgg.BLOB_INDEX['conv5_1'] + 1):
        name, kind, cout = vgg.TOPOLOGY[b]
        if kind != 'conv' or (args.layers and name not in args.layers.split(',')):
            continue
        cin = vgg.TOPOLOGY[b - 1][2]
        c, h, w = plan.blob_dims(b)
        fl = 18.0 * cin * cout * h * w
        for d in (0, 1):
            ms = C.c_float()
            plan._check(plan.lib.st2_bench_layer(plan.handle, b, d, args.reps, C.byref(ms)), 'bench_layer')
            rows.append({'flags': flags, 'layer': name, 'dir': 'fwd' if d == 0 else 'dgrad', 'us': ms.value * 1e3,
                         'tflops': fl / (ms.value * 1e-3) / 1e12, 'gflop': fl / 1e9})
            print('flags=%d %-8s %-5s %8.1f us  %7.1f TFLOP/s  (%5.1f GF, M=%d N=%d K=%d)' % (
                flags, name, rows[-1]['dir'], rows[-1]['us'], rows[-1]['tflops'], fl / 1e9, h * w,
                cout if d == 0 else cin, 9 * (cin if d == 0 else cout)))
m.engine.call('st2_debug_flags', 0)
tot = {}
for r in rows:
    tot.setdefault(r['flags'], [0.0, 0.0])
    tot[r['flags']][0] += r['us']
    tot[r['flags']][1] += r['gflop']
for f, (us, gf) in tot.items():
    print('flags=%d total %.1f us, %.1f TFLOP/s' % (f, us, gf / us / 1e3))
print(json.dumps(rows))
