"""Hardware experiment behind tc_conv_ws_kernel: conv1_2 with delta weights (one tap, identity over channels) so
that the output must equal the shifted input; run with st2_debug_flags 0 / 32 it showed that the UMMA descriptor's
base_offset must stay 0 for 128-byte-aligned starts inside a TMA-written SWIZZLE_128B patch (flag 32 has since been
removed from the kernel; kept as the record of the method)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from style_transfer2_b200 import vgg
from style_transfer2_b200.model import B200Model, Plan
np.set_printoptions(linewidth=250, precision=0, suppress=True)
params = vgg.synthetic_weights(0)
H, W = 32, 24
for flags in (0, 32):
    for tap in (4, 0, 1, 2, 3, 5, 8):
        r, s = tap // 3, tap % 3
        w = np.zeros((64, 64, 3, 3), np.float32)
        for c in range(64):
            w[c, c, r, s] = 1.0
        prm = dict(params); prm['conv1_2'] = (w, np.zeros(64, np.float32))
        m = B200Model(precision='fp16', params=prm)
        m.engine.call('st2_debug_flags', flags)
        x = torch.randn(1, 3, H, W, device=m.engine.device) * 50
        p1 = Plan(m.engine, H, W, m.precision); p1.forward(x, 2)
        a = p1.export(1).cpu().numpy()[0]; got = p1.export(2).cpu().numpy()[0]
        want = np.zeros_like(a)
        pad = np.pad(a, ((0, 0), (1, 1), (1, 1)))
        want = pad[:, r:r + H, s:s + W]
        bad = (np.abs(got - want).max(axis=0) > 1e-2)
        print('flags', flags, 'tap', tap, 'bad pixels', int(bad.sum()), 'of', H * W)
        if bad.sum() and tap in (4, 0, 5):
            print(bad.astype(int)[:18])
            # where does got come from? find the shift (dr, ds) that explains channel 0 of pixel (5,5) region
            best = None
            for dr in range(-3, 4):
                for ds in range(-9, 10):
                    sh = np.roll(np.roll(pad[:, 1:1 + H, 1:1 + W], -dr, axis=1), -ds, axis=2)
                    e = np.abs(got[:, 4:12, 8:16] - sh[:, 4:12, 8:16]).max()
                    if best is None or e < best[0]:
                        best = (e, dr, ds)
            print('best shift for interior block', best)
        p1.close()
