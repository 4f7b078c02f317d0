#!/usr/bin/env python3
"""Experiment: how much of an iteration is launch gaps?  Times K L-BFGS steps enqueued normally and the same steps
replayed from a CUDA graph (two steps per graph: the gradient double buffer alternates).  This was the experiment behind
StyleTransfer._graph_step (worker.py), which is what production steps use now.  usage: python tools/graph_probe.py [size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
st, _ = bench.build_job(size, 'fp16')
st.use_graphs = False          # the experiment compares plain enqueueing with its own capture (StyleTransfer now replays graphs itself)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(6):
        st.step(fetch=False)
    torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(n):
            fn()
        e1.record(s)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    plain = timed(lambda: st.step(fetch=False), 40)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        st.step(fetch=False)
        st.step(fetch=False)
    g.replay()
    torch.cuda.synchronize()
    graph = timed(g.replay, 20) / 2
    plain2 = timed(lambda: st.step(fetch=False), 40)
print('size %d: plain %.4f ms/step, graph %.4f ms/step, plain again %.4f' % (size, plain, graph, plain2))
