"""Element-wise comparison of the tensor-core conv1_1 kernels (sliding-window forward, N=16 data gradient) with the
CUDA-core fp32 kernels they replaced (ST2_NO_TC_FIRST=1): forward within 1 fp16 ulp, data gradient 2e-4 relative."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from style_transfer2_b200.model import B200Model, Plan
np.set_printoptions(linewidth=250, precision=3, suppress=True)
m = B200Model(precision='fp16')
for (H, W) in ((37, 45), (64, 64)):
    rs = np.random.RandomState(11)
    x = torch.from_numpy((rs.rand(1, 3, H, W) * 255 - 120).astype(np.float32)).to(m.engine.device)
    os.environ['ST2_NO_TC_FIRST'] = '1'
    p0 = Plan(m.engine, H, W, m.precision); p0.forward(x, 1); ref = p0.export(1).cpu().numpy()[0]
    del os.environ['ST2_NO_TC_FIRST']
    p1 = Plan(m.engine, H, W, m.precision); p1.forward(x, 1); got = p1.export(1).cpu().numpy()[0]
    err = np.abs(got - ref).max(axis=0)
    print(H, W, 'max ref', np.abs(ref).max(), 'max err', err.max(), 'n bad', int((err > 0.3).sum()))
    ys, xs = np.nonzero(err > 0.3)
    print('bad rows', sorted(set(ys.tolist()))[:40], 'bad cols', sorted(set(xs.tolist()))[:40])
    # backward check: gradient of sum(conv1_1 * R)
    R = torch.from_numpy(rs.randn(1, 64, H, W).astype(np.float32)).to(m.engine.device)
    g0 = torch.empty(1, 3, H, W, device=m.engine.device); g1 = torch.empty_like(g0)
    p0.backward({1: R}, g0); p1.backward({1: R}, g1)
    ge = (g1 - g0).abs().cpu().numpy()[0].max(axis=0)
    print('bwd max ref', float(g0.abs().max()), 'max err', ge.max(), 'rel', float((g1 - g0).norm() / g0.norm()))
    ys, xs = np.nonzero(ge > 0.02 * float(g0.abs().max()))
    print('bwd bad rows', sorted(set(ys.tolist()))[:40], 'bad cols', sorted(set(xs.tolist()))[:40])
