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

``--workload canvas`` (BASELINE config 4): ONE --size x --size canvas (default 4096) split into row
strips over the N GPUs (style_transfer2_b200/tiled.py: halo rows over peer memory, four small NCCL
all-reduces per iteration) -> strong scaling; value = iterations of the whole canvas per second.

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


def pool_extent(n):
    return (n - 2 + 1) // 2 + 1 if n > 1 else 1


def conv_flops(h, w, first=0, last=12):
    """Algorithmic conv flops of one direction, layers first..last (18 Cin Cout H_l W_l each)."""
    dims, total = [(h, w)], 0
    for _ in range(4):
        dims.append((pool_extent(dims[-1][0]), pool_extent(dims[-1][1])))
    for i in range(first, last + 1):
        cin, cout = CONV_CH[i]
        hh, ww = dims[CONV_POOL_BEFORE[i]]
        total += 18 * cin * cout * hh * ww
    return total


def load_images(size):
    from PIL import Image
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'config1.npz'))
    content = Image.fromarray(g['content']).resize((size, size), Image.LANCZOS)
    sh, sw = g['style'].shape[:2]
    style = Image.fromarray(g['style']).resize((size, max(1, int(round(size * sh / sw)))), Image.LANCZOS)
    x0 = np.uint8(np.random.RandomState(0).uniform(0, 255, (size, size, 3)))
    return np.uint8(content), np.uint8(style), x0


class ClockSampler:
    """nvidia-smi clocks line of the profiling recipe, sampled during the timed region."""

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

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        # samples inside the timed region; the sampler runs from before the warm-up, so if the region was
        # shorter than a sampling period fall back to the samples taken under the same load just around it
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


# ------------------------------------------------------------------------------------ CPU arm
def run_cpu_reference(size, steps, warmup, budget_s, full_net=True):
    """The reference's CPU path restated (oracle/): StyleTransfer + L-BFGS + Caffe-CPU layer
    semantics on torch-CPU fp32 with all host threads.  Returns (it/s, cores, sample description)."""
    import torch
    from oracle.caffe_cpu import CaffeCPUModel
    from oracle.transfer import Transfer
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    content, style, x0 = load_images(size)
    st = Transfer(CaffeCPUModel(full_net=full_net))
    st.set_input(x0)
    st.set_content(content)
    st.set_style(style)
    st.set_weights(WEIGHTS, PARAMS)
    assert st.start()
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
    return its, cores, sample, len(times)


# ------------------------------------------------------------------------------------ GPU arm
def build_job(size, precision, seed_shift=0):
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
    return st


class TiledJob:
    """The same step()/input surface as StyleTransfer for one row-tiled canvas (this rank's strip)."""

    def __init__(self, size, precision):
        from style_transfer2_b200.model import B200Model
        from style_transfer2_b200.tiled import TiledTransfer
        content, style, x0 = load_images(size)
        model = B200Model(gpu=int(os.environ.get('LOCAL_RANK', 0)), precision=precision)
        self.tt = TiledTransfer(model, size, size)
        self.tt.set_input(x0)
        self.tt.set_content(content)
        self.tt.set_style(style)
        self.tt.set_weights(WEIGHTS, PARAMS)
        self.engine = model.engine

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
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    canvas = args.workload == 'canvas'
    size = args.size or (4096 if canvas else 1024)
    flops_conv = 2 * conv_flops(size, size)
    gram_flops = 0
    dims = [(size, size)]
    for _ in range(4):
        dims.append((pool_extent(dims[-1][0]), pool_extent(dims[-1][1])))
    for i, c in enumerate((64, 128, 256, 512, 512)):
        gram_flops += 4 * c * c * dims[i][0] * dims[i][1]
    config = {'workload': ('config4: one %dx%d canvas in row strips, ' if canvas else 'config2: %dx%d canvas, ') % (size, size) +
                          'style conv1_1..conv5_1 + content conv4_2, tv/p, L-BFGS m=10',
              'canvas': [size, size], 'optimizer': 'lbfgs', 'precision': args.precision,
              'parallelism': ('row strips x%d (halo rows over peer memory, 4 all-reduces/iteration)' % world) if canvas
                             else ('independent jobs x%d' % world if world > 1 else 'single job'),
              'l2': 'working set (>=1.3 GB activations + 0.25 GB L-BFGS history per iteration) exceeds the 126 MB L2',
              'algorithmic_tflop_per_iteration': round((flops_conv + gram_flops) / 1e12, 4)}

    if args.impl == 'reference':
        if rank != 0:
            return
        its, cores, sample, n = run_cpu_reference(size, args.steps, args.warmup, max(args.cpu_budget * 6, 60.0))
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
        from style_transfer2_b200 import parallel, serving
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

    st = TiledJob(size, args.precision) if canvas else build_job(size, args.precision, seed_shift=rank)
    eng = st.engine
    jobs = 1 if canvas else world                 # whole-job units per step
    # ---- device-resident arm: `value`
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        st.step(fetch=False)
    barrier()
    l0 = eng.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        st.step(fetch=False)
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = eng.launches() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    if world > 1:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = jobs * args.steps / (ms / 1000.0)

    # ---- end-to-end arm: the public step API with HOST buffers.  Every step uploads x from pinned host
    # memory, steps, and reads back to pinned host memory the new x (next step's upload: a true data
    # dependency, stream-ordered on the compute stream), the iterate image (HxWx3 fp32) and the trace.  The
    # image travels on a side stream while the next iteration computes (StyleTransfer.step_async), as in
    # the worker loop; the host reads the loss and a pixel of every iterate.
    x_host = torch.empty(st.input.shape, dtype=torch.float32, pin_memory=True)
    x_host.copy_(st.input)
    torch.cuda.synchronize()

    def e2e_steps(n):
        pending, sink = None, 0.0
        for k in range(n):
            st.input.copy_(x_host, non_blocking=True)
            if canvas:
                _, tr = st.step()
                x_host.copy_(st.input, non_blocking=True)
                sink += float(tr['loss'])
                continue
            handle = st.step_async()
            x_host.copy_(st.input, non_blocking=True)
            if pending is not None:
                img, tr = pending.result()
                sink += float(tr['loss']) + float(img[0, 0, 0])
            pending = handle
        if pending is not None:
            img, tr = pending.result()
            sink += float(tr['loss']) + float(img[0, 0, 0])
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
               'note': 'StyleTransfer.step_async(): x uploaded from pinned host memory every step; new x + iterate image (HxWx3 fp32) + trace block read back every step, the image copy overlapping the next iteration'}

    # ---- per-category device time (CUDA events on the launch stream) for the roofline
    import ctypes as C
    from style_transfer2_b200 import _lib
    eng.call('st2_profile', 1)
    prof_steps = min(args.steps, 10)
    for _ in range(prof_steps):
        st.step(fetch=False)
    ms_cat = (C.c_double * _lib.PROF_CATS)()
    n_cat = (C.c_longlong * _lib.PROF_CATS)()
    eng.call('st2_profile_read', ms_cat, n_cat)
    eng.call('st2_profile', 0)
    cats = {name: {'ms_per_step': ms_cat[i] / prof_steps, 'launch_spans_per_step': n_cat[i] / prof_steps}
            for i, name in enumerate(_lib.PROF_NAMES) if n_cat[i]}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    roofline = None
    key = 'conv_tc' if 'conv_tc' in cats else ('conv_exact' if 'conv_exact' in cats else None)
    if key:
        fl = 2 * conv_flops(size, size, first=1) // (world if canvas else 1)   # conv1_2..conv5_1, fwd + dgrad, this GPU's share
        t_s = cats[key]['ms_per_step'] / 1000.0
        peak = peaks.get('bf16_tflops_sustained', 1400.0)
        ach = fl / t_s / 1e12
        # DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture of one
        # iteration (profiles/*_tcconv_traffic.json, made by profiles/ncu_traffic.py); 1024^2 workload only
        traffic, traffic_src = None, None
        if key == 'conv_tc' and size == 1024 and not canvas:
            import glob
            found = sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_tcconv_traffic.json')))
            if found:
                tj = json.load(open(found[-1]))
                traffic, traffic_src = tj['conv3x3_dram_bytes_per_launch'], os.path.relpath(found[-1], ROOT)
        n_launch = cats[key]['launch_spans_per_step']
        roofline = {'kernel': 'tc_conv_kernel (tcgen05 implicit GEMM, fwd + dgrad, 24 launches/iteration)' if key == 'conv_tc' else 'conv_exact_kernel',
                    'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                    'traffic': traffic, 'traffic_unit': 'DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)',
                    'traffic_source': traffic_src, 'launches_per_step': n_launch,
                    'flops_per_launch': fl / n_launch if n_launch else None,
                    'avg_launch_ms': cats[key]['ms_per_step'] / n_launch if n_launch else None,
                    'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if peaks else 'fallback 1.4 PFLOP/s (of fallback)',
                    'flops_per_step': fl, 'ms_per_step': cats[key]['ms_per_step']}

    line = {'metric': 'style-transfer iterations/sec', 'value': value, 'unit': 'it/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if canvas else 'weak', 'vs_baseline': None,
            'dtype': 'f16 operands / f32 accumulate' if args.precision == 'fp16' else 'f32',
            'data': 'synthetic', 'config': config, 'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches,
            'roofline': roofline, 'kernel_time_ms_per_step': cats}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        its, cores, sample, _ = run_cpu_reference(size, 2, 1, args.cpu_budget)
        line['cpu_baseline'] = {'value': its, 'unit': 'it/s', 'cores': cores, 'kind': 'port', 'sample': sample}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
