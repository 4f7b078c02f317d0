"""Element-wise comparison of the opt-in weight-stationary CTA-pair kernel (ST2_WSP=1) with the generic kernels on
conv2_1 / conv2_2 forward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from style_transfer2_b200 import vgg
from style_transfer2_b200.model import B200Model, Plan
np.set_printoptions(linewidth=250, precision=0, suppress=True)
m = B200Model(precision='fp16')
os.environ['ST2_FORCE_PAIR'] = '1'
for (H, W) in ((128, 128), (96, 80)):
    x = torch.randn(1, 3, H, W, device=m.engine.device) * 50
    for blob in (4, 5):
        os.environ['ST2_NO_WSP'] = '1'
        p0 = Plan(m.engine, H, W, m.precision); p0.forward(x, blob); ref = p0.export(blob).cpu().numpy()[0]
        del os.environ['ST2_NO_WSP']
        p1 = Plan(m.engine, H, W, m.precision); p1.forward(x, blob); got = p1.export(blob).cpu().numpy()[0]
        err = np.abs(got - ref).max(axis=0)
        bad = err > 1e-2 * np.abs(ref).max()
        print(H, W, vgg.BLOBS[blob], 'max ref', np.abs(ref).max(), 'max err', err.max(), 'bad px', int(bad.sum()), 'of', bad.size)
        if bad.sum():
            ys, xs = np.nonzero(bad)
            print(' bad rows', sorted(set(ys.tolist()))[:48]); print(' bad cols', sorted(set(xs.tolist()))[:48])
            ch = np.abs(got - ref).max(axis=(1, 2)); print(' bad channels', np.nonzero(ch > 1e-2 * np.abs(ref).max())[0][:64])
        p0.close(); p1.close()
