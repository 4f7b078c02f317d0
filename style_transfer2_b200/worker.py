#!/usr/bin/env python3
"""Drop-in worker for style_transfer2: same ZeroMQ/pickle protocol (``messages.py``), same
``config.ini`` knobs, same layer names, same optimizer step interface -- with the Caffe/NumPy
compute path replaced by libst2's sm_100a kernels.  Mirrors ``/root/reference/worker.py``:
``StyleTransfer`` (worker.py:117-315) and ``Worker`` (worker.py:318-409).

Everything from the parameters ``x`` to the next ``x`` stays on the GPU; per iterate only the
deprocessed image and one block of ~500 scalars (the trace) come back.
"""
from collections import OrderedDict
import ctypes as C
import logging
import math
import os
import pickle
import sys
import time

import numpy as np
import pandas as pd
import torch

from . import _lib, messages, optimizers, utils, vgg
from .messages import (GetImages, Iterate, PauseIteration, SetImages, SetOptimizer, SetWeights, Shutdown,
                       StartIteration, WorkerReady)
from .model import B200Model, Plan

logger = logging.getLogger('worker')
EPS_W = 1e-15


def gram_matrix(x):
    """worker.py:109-114 for a (1, C, H, W) fp32 CUDA tensor: ``X X^T / (C*H*W)``, on the device."""
    n, c, h, w = x.shape
    assert n == 1
    x = x.contiguous()
    out = torch.empty((c, c), dtype=torch.float32, device=x.device)
    utils.default_engine().call('st2_gram_nchw', C.c_void_p(x.data_ptr()), c, h * w, C.c_void_p(out.data_ptr()))
    return out


class LazyTrace:
    """One evaluation's trace (utils.Trace, utils.py:257-282).  The scalar block is copied to pinned
    host memory asynchronously when the evaluation is enqueued; the dict is built (and the stream
    event waited on) the first time ``data`` is read, so evaluations never stall the pipeline.
    ``time`` is the host clock when the evaluation was ENQUEUED (the reference stamps it after its
    synchronous evaluation, worker.py:298); differences between consecutive entries are still one
    iteration each in steady state."""

    def __init__(self, host_block, event, spec, want_grad, when):
        self._host, self._event, self._spec = host_block, event, spec
        self._detached = False
        self._want_grad, self._when = want_grad, when
        self._data = None
        self._extra = []

    def __call__(self, name, value):
        if self._data is None:
            self._extra.append((name, value))
        else:
            self._put(name, value)
        return value

    def detach(self):
        """Move the scalar block out of the shared pinned ring (its row is about to be reused)."""
        if not self._detached and self._host is not None:
            self._event.synchronize()
            self._host = self._host.clone()
            self._detached = True

    def _put(self, name, value):
        while name in self._data:
            name += '_'
        if isinstance(value, np.floating):
            value = float(value)
        elif isinstance(value, np.integer):
            value = int(value)
        self._data[name] = value

    @property
    def data(self):
        if self._data is None:
            self._event.synchronize()
            s = self._host.numpy()
            self._data = OrderedDict()
            for b, c_on, s_on, d_on in self._spec:
                base = b * _lib.SCAL_PER_BLOB
                name = vgg.BLOBS[b]
                if c_on:
                    self._put('%s_c_loss' % name, float(s[base + _lib.SB_C_LOSS]))
                    self._put('%s_c_grad' % name, float(s[base + _lib.SB_C_GRAD]))
                if s_on:
                    self._put('%s_s_loss' % name, float(s[base + _lib.SB_S_LOSS]))
                    self._put('%s_s_grad' % name, float(s[base + _lib.SB_S_GRAD]))
                if d_on:
                    self._put('%s_d_loss' % name, float(s[base + _lib.SB_D_LOSS]))
                    self._put('%s_d_grad' % name, float(s[base + _lib.SB_D_GRAD]))
            g = s[_lib.SCAL_GLOBAL_BASE:]
            self._put('scd_loss', float(g[_lib.G_SCD_LOSS]))
            self._put('t_loss', float(g[_lib.G_T_LOSS]))
            self._put('p_loss', float(g[_lib.G_P_LOSS]))
            if self._want_grad:
                self._put('scd_grad', float(g[_lib.G_SCD_GRAD]))
                self._put('t_grad', float(g[_lib.G_T_GRAD]))
                self._put('p_grad', float(g[_lib.G_P_GRAD]))
                self._put('time', self._when)
            self._put('loss', float(g[_lib.G_LOSS]))
            if self._want_grad:
                self._put('grad', float(g[_lib.G_GRAD]))
            for name, value in self._extra:
                self._put(name, value)
            self._extra = []
        return self._data

    @property
    def loss(self):
        return self.data['loss']

    def norms(self):
        """{'c'|'s'|'d': {layer: frozen normaliser}} as of this evaluation."""
        self.data
        s = self._host.numpy()
        out = {k: {} for k in 'cds'}
        for b in range(_lib.NUM_BLOBS):
            base = b * _lib.SCAL_PER_BLOB
            for k, (nf, vf) in zip('csd', ((_lib.SB_C_NORM, _lib.SB_C_VALID), (_lib.SB_S_NORM, _lib.SB_S_VALID),
                                          (_lib.SB_D_NORM, _lib.SB_D_VALID))):
                if s[base + vf] != 0.0:
                    out[k][vgg.BLOBS[b]] = float(s[base + nf])
        return out

    def __str__(self):
        return ', '.join('%s: %g' % item for item in self.data.items())


class LazyLoss:
    """Loss of an evaluation; converts to float on demand (``float(loss)``, comparisons, format)."""

    def __init__(self, trace):
        self.trace = trace

    def __float__(self):
        return float(self.trace.loss)

    def __repr__(self):
        return repr(float(self))

    def __format__(self, spec):
        return format(float(self), spec)


class IterateHandle:
    """An iterate on its way to the host (``StyleTransfer.step_async``): the deprocessed image is copied
    device -> pinned host on a side stream while the next iteration computes.  ``result()`` waits for the
    copy and returns ``(image, trace)`` exactly as ``step()`` does.  The image array is a view of a
    double-buffered pinned block: consume (or copy) it before the second-next ``step_async``."""

    def __init__(self, host, done, trace, t, start=None):
        self._host, self._done, self.trace, self.t = host, done, trace, t
        self._start = start              # deferred download (step_async(download=False)): enqueues the copy, returns its event

    def download(self, after=None):
        """Enqueue the device -> host copy of a handle made with ``step_async(download=False)``; ``after``: an event
        the copy must wait for (e.g. another read-back that should get the copy engine first)."""
        if self._start is not None:
            self._done = self._start(after)
            self._start = None

    def result(self):
        self.download()
        self._done.synchronize()
        return self._host.numpy(), self.trace.data


class StyleTransfer:
    """worker.py:117-315 with device-resident state."""

    TRACE_KEEP = 256

    def __init__(self, model, private_plans=False):
        self.model = model
        self.engine = model.engine
        # Plans carry per-job state (content/style targets, frozen normalisers).  The reference runs one job
        # per worker process; several jobs resident on one GPU (serving.JobScheduler) each need their own.
        self._private_plans = {} if private_plans else None
        utils.set_default_engine(self.engine)
        self.is_running = False
        self.is_starting = False
        self.t = 0
        self.input = None              # fp32 CUDA tensor (1, 3, H, W), preprocessed
        self.content = None            # fp32 CUDA tensor (1, 3, H, W), preprocessed
        self.style = None              # fp32 CUDA tensor (1, 3, Hs, Ws), preprocessed
        shape = (len(model.layers()), len(SetWeights.loss_names))
        self.weights = pd.DataFrame(np.ones(shape), model.layers(), SetWeights.loss_names, np.float32)
        self.params = {w: 1 for w in SetWeights.scalar_loss_names}
        self.optimizer = None
        self.optimizer_cls = optimizers.LBFGSOptimizer
        self.step_size = SetOptimizer.step_sizes['lbfgs']
        self.traces = []
        self._plan = None
        self._weights_dirty = True
        self._content_done = set()
        self._style_done = set()
        self._pending_norms = {}       # (kind, layer) -> value to install on the next plan sync
        self._grad_bufs = [None, None]
        self._grad_turn = 0
        self._spec = []
        self._pinned = []
        self._norm_source = None       # the latest evaluation since the last reset()
        self._norms_reset = False
        # CUDA graphs of the steady-state L-BFGS iteration (see _graph_step): dropped by every state change
        self.use_graphs = os.environ.get('ST2_NO_GRAPH') is None
        self._graphs = None
        self._stable_steps = 0

    # ------------------------------------------------------------------ reference-compatible views
    @property
    def grams(self):
        """Truthy once a style image is set (worker.py:140-144 only tests truthiness)."""
        return {'style': self.style} if self.style is not None else None

    @property
    def features(self):
        return {'content': self.content} if self.content is not None else None

    @property
    def norms(self):
        if self._norm_source is not None:
            out = self._norm_source.norms()
        else:
            out = {k: {} for k in 'cds'}
        for (kind, layer), v in self._pending_norms.items():
            out[kind][layer] = v
        return out

    def set_norms(self, norms):
        """Install frozen normalisers ({'c'|'s'|'d': {layer: value}}), e.g. from a checkpoint."""
        self._drop_graphs()
        for kind, table in norms.items():
            for layer, v in table.items():
                self._pending_norms[(kind, layer)] = float(v)

    # ------------------------------------------------------------------ host <-> device plumbing
    def _upload_image(self, image):
        """HxWx3 RGB array -> preprocessed (1, 3, H, W) CUDA tensor (CaffeModel.preprocess, worker.py:63-66)."""
        arr = np.asarray(image)
        h, w = arr.shape[:2]
        out = self.engine.empty(1, 3, h, w)
        if arr.dtype == np.uint8:
            dev = torch.from_numpy(np.array(arr, copy=True)).to(self.engine.device)     # (messages may hold read-only arrays)
            self.engine.call('st2_preprocess_u8', C.c_void_p(dev.data_ptr()), C.c_void_p(out.data_ptr()), h, w)
        else:
            dev = torch.from_numpy(np.ascontiguousarray(arr, np.float32)).to(self.engine.device)
            self.engine.call('st2_preprocess_f32', C.c_void_p(dev.data_ptr()), C.c_void_p(out.data_ptr()), h, w)
        return out

    def image(self, x=None):
        """Deprocessed iterate as an HxWx3 fp32 host array (CaffeModel.deprocess, worker.py:68-71)."""
        x = self.input if x is None else x
        h, w = x.shape[2:]
        hwc = self.engine.empty(h, w, 3)
        self.engine.call('st2_deprocess', C.c_void_p(x.data_ptr()), C.c_void_p(hwc.data_ptr()), h, w)
        host = self._pinned_buffer((h, w, 3))
        host.copy_(hwc, non_blocking=True)
        torch.cuda.current_stream(self.engine.device).synchronize()
        # a fresh array per iterate, as the reference returns (worker.py:68-71): callers may keep iterates across
        # steps.  The pipelined path (step_async / IterateHandle) hands out the pinned double buffer itself.
        return np.array(host.numpy())

    def image_async(self, x, trace, download=True):
        """Enqueue deprocess + device->host copy of ``x`` on the download stream; no host wait."""
        dev = self.engine.device
        main = torch.cuda.current_stream(dev)
        h, w = x.shape[2:]
        if getattr(self, '_dl', None) is None or self._dl['shape'] != (h, w):
            self._dl = {'shape': (h, w), 'stream': torch.cuda.Stream(dev), 'turn': 0,
                        'dev': [self.engine.empty(h, w, 3) for _ in range(2)],
                        'host': [torch.empty((h, w, 3), dtype=torch.float32, pin_memory=True) for _ in range(2)],
                        'done': [None, None]}
        d = self._dl
        d['turn'] ^= 1
        k = d['turn']
        if d['done'][k] is not None:
            main.wait_event(d['done'][k])              # the slot's previous copy must have left the device buffer
        self.engine.call('st2_deprocess', C.c_void_p(x.data_ptr()), C.c_void_p(d['dev'][k].data_ptr()), h, w)
        ready = torch.cuda.Event()
        ready.record(main)

        def start(after=None):
            with torch.cuda.stream(d['stream']):
                d['stream'].wait_event(ready)
                if after is not None:
                    d['stream'].wait_event(after)
                d['host'][k].copy_(d['dev'][k], non_blocking=True)
                done = torch.cuda.Event()
                done.record(d['stream'])
            d['done'][k] = done
            return done

        if download:
            return IterateHandle(d['host'][k], start(), trace, self.t)
        return IterateHandle(d['host'][k], None, trace, self.t, start)

    def _pinned_buffer(self, shape):
        for buf in self._pinned:
            if tuple(buf.shape) == tuple(shape):
                return buf
        buf = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        self._pinned = [buf] + self._pinned[:1]
        return buf

    # ------------------------------------------------------------------ state machine
    def check_consistency(self):
        return bool(self.input is not None and self.content is not None and self.grams
                    and self.input.shape == self.content.shape)

    def objective_changed(self):
        self._drop_graphs()
        if self.optimizer is not None:
            self.optimizer.objective_changed()

    def pause(self):
        self.is_running = False
        self.is_starting = False

    def reset(self):
        """worker.py:172-175."""
        self._drop_graphs()
        self._pending_norms = {}
        if self._plan is not None:
            self._plan.reset_norms()
        self._norms_reset = True
        self._norm_source = None
        self.t = 0
        self.optimizer = self.optimizer_cls(self.input, self.opfunc, step_size=self.step_size)

    def start(self):
        self.is_starting = True
        self._start()
        return self.is_running

    def _start(self):
        """worker.py:182-189."""
        if self.is_starting and self.check_consistency():
            if self.optimizer is None:
                self.reset()
            self.is_starting = False
            self.is_running = True

    def set_input(self, image):
        """worker.py:191-202."""
        self._drop_graphs()
        image = self._upload_image(image)
        if self.input is not None and self.input.shape == image.shape:
            self.input.copy_(image)
            self.objective_changed()
        elif self.optimizer is not None:
            self.input = self.optimizer.resample(None, new_x=image)
            self._start()
        else:
            self.input = image
            self.reset()
            self._start()

    def set_content(self, image):
        """worker.py:204-209.  Content features are (re)captured lazily for the layers that carry
        a content weight, from the stored content image."""
        self.content = self._upload_image(image)
        self._content_done = set()
        self._start()
        self.objective_changed()

    def set_style(self, image):
        """worker.py:211-218.  Gram targets are computed lazily for the style-weighted layers."""
        self.style = self._upload_image(image)
        self._style_done = set()
        self._start()
        self.objective_changed()

    def resample_input(self, size):
        """worker.py:154-160."""
        if self.input is not None and self.optimizer is not None:
            self.input = self.optimizer.resample(tuple(size))
        else:
            self.input = self.engine.zeros(1, 3, *size)
        self._start()
        self.objective_changed()

    def resample_content(self, size):
        """worker.py:162-170."""
        if self.content is not None:
            self.content = utils.resample_nchw(self.content, tuple(size))
        else:
            self.content = self.engine.zeros(1, 3, *size)
        self._content_done = set()
        self._start()
        self.objective_changed()

    def set_step_size(self, step_size):
        self._drop_graphs()
        self.step_size = step_size
        if self.optimizer is not None:
            self.optimizer.step_size = step_size

    def set_optimizer_class(self, cls, step_size):
        """worker.py:387-391 (SetOptimizer): switch class / step size; a class change resets the job state."""
        self.optimizer_cls = cls
        self.set_step_size(step_size)
        if not isinstance(self.optimizer, self.optimizer_cls):
            self.reset()

    def set_weights(self, weights, params):
        """worker.py:226-229."""
        self._drop_graphs()
        self.weights = pd.DataFrame.from_dict(weights, dtype=np.float32)
        self.params = params
        self._weights_dirty = True
        self.objective_changed()

    # ------------------------------------------------------------------ objective
    def active_layers(self):
        """worker.py:234-235: rows with any |w| > 1e-15 (NaN counts as off), in table order."""
        nonzeros = abs(self.weights) > EPS_W
        return list(self.weights.index[abs(nonzeros.sum(axis=1)) > EPS_W])

    def _sync_plan(self):
        h, w = self.input.shape[2:]
        if self._private_plans is None:
            plan = self.model.plan(h, w)
        else:
            plan = self._private_plans.get((h, w))
            if plan is None:
                for old in self._private_plans.values():
                    self.model.release_plan(old)
                self._private_plans.clear()
                plan = self._private_plans[(h, w)] = self.model.acquire_plan(h, w)
        if plan is not self._plan:
            if self._plan is not None and self._plan.handle and not getattr(self, '_norms_reset', False):
                for kind, table in self.norms.items():          # norms survive a scale change
                    for layer, v in table.items():
                        self._pending_norms.setdefault((kind, layer), v)
            self._plan = plan
            plan.reset_norms()
            self._weights_dirty = True
            self._content_done, self._style_done = set(), set()
        self._norms_reset = False
        if self._weights_dirty:
            spec, order = [], []
            table = self.weights
            for b, name in enumerate(vgg.BLOBS):
                plan.set_blob_weights(b, 0.0, 0.0, 0.0)
            for name in self.active_layers():
                if name not in vgg.BLOB_INDEX:
                    raise KeyError('unknown layer %r' % name)
                b = vgg.BLOB_INDEX[name]
                vals = [float(table[col][name]) if col in table.columns else 0.0
                        for col in SetWeights.loss_names]
                vals = [0.0 if (math.isnan(v) or abs(v) <= EPS_W) else v for v in vals]
                plan.set_blob_weights(b, *vals)
                order.append(b)
                spec.append((b, vals[0] != 0.0, vals[1] != 0.0, vals[2] != 0.0))
            plan.set_eval_order(order)
            plan.set_params(float(self.params['tv']), float(self.params['tv_power']), float(self.params['p']),
                            float(self.params['p_power']))
            self._spec = spec
            self._weights_dirty = False
        for (kind, layer), v in list(self._pending_norms.items()):
            plan.set_norm(kind, vgg.BLOB_INDEX[layer], v)
        self._pending_norms = {}
        need_c = [b for b, c_on, _, _ in self._spec if c_on and b not in self._content_done]
        if need_c:
            plan.forward(self.content, max(need_c))
            for b in need_c:
                plan.capture_content(b)
                self._content_done.add(b)
        need_s = [b for b, _, s_on, _ in self._spec if s_on and b not in self._style_done]
        if need_s:
            hs, ws = self.style.shape[2:]
            sp = plan if (hs, ws) == (h, w) else self.model.acquire_plan(hs, ws)
            try:
                sp.forward(self.style, max(need_s))
                for b in need_s:
                    plan.set_style_gram(b, sp.gram(b))
                    self._style_done.add(b)
            finally:
                if sp is not plan:
                    self.model.release_plan(sp, keep=4 if self._private_plans is None else 12)
        return plan

    def _next_grad(self):
        self._grad_turn ^= 1
        buf = self._grad_bufs[self._grad_turn]
        if buf is None or buf.shape != self.input.shape:
            buf = self._grad_bufs[self._grad_turn] = torch.empty_like(self.input)
        return buf

    def opfunc(self, x, return_grad=True):
        """worker.py:231-301: objective and gradient at ``x`` (fp32 CUDA tensor).  Returns
        ``(loss, grad)`` -- ``loss`` converts to float on demand, ``grad`` is a device tensor owned
        by the caller until the call after next -- or just ``loss`` when ``return_grad`` is false."""
        self.engine.sync_stream()
        plan = self._sync_plan()
        grad = self._next_grad() if return_grad else None
        plan.eval(x, grad, return_grad)
        tr = LazyTrace(None, None, list(self._spec), return_grad, time.perf_counter())
        tr._host = host = self.engine.trace_slot(tr)
        plan.copy_scalars_async(host)
        tr._event = ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.engine.device))
        self.traces.append(tr)
        self._norm_source = tr
        if len(self.traces) > self.TRACE_KEEP:
            del self.traces[:-self.TRACE_KEEP]
        if not return_grad:
            return LazyLoss(tr)
        return LazyLoss(tr), grad

    # ------------------------------------------------------------------ CUDA graphs of the steady-state iteration
    GRAPH_ROWS = 4                     # trace rows (and evaluation graphs) in rotation

    def _drop_graphs(self):
        self._graphs = None
        self._stable_steps = 0

    def _graph_capture(self):
        """Capture one L-BFGS iteration as two graphs -- [advance] and [evaluate + commit] -- so that a step is two
        graph launches instead of ~56 kernel launches (measured: 2.10 -> 2.00 ms at 1024^2, 0.97 -> 0.86 ms at 512^2:
        the rest was launch gaps).  Everything a step touches has a fixed address: x, the plan's buffers, the L-BFGS
        state, and two gradient buffers -- the evaluation always writes G1, the commit reads (G1, G0) and G0 <- G1
        follows inside the graph (the eager path alternates two buffers instead).  The scalar block of an evaluation
        lands in one of GRAPH_ROWS pinned rows, one evaluation graph per row, used round-robin."""
        opt, dev = self.optimizer, self.engine.device
        plan = self._sync_plan()
        x = self.input
        g0, g1 = torch.empty_like(x), torch.empty_like(x)
        g0.copy_(opt.grad)
        opt.grad = g0
        rows = torch.empty((self.GRAPH_ROWS, _lib.SCAL_TOTAL), dtype=torch.float64, pin_memory=True)
        torch.cuda.current_stream(dev).synchronize()
        l0 = self.engine.launches()
        adv, evals = None, []
        for k in range(self.GRAPH_ROWS):
            ga = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga, capture_error_mode='thread_local'):
                self.engine.sync_stream()
                opt._call('st2_lbfgs_advance', optimizers._p(x), optimizers._p(g0), float(opt.step_size))
            if adv is None:
                adv = ga                       # the others only keep libst2's host-side step bookkeeping in sequence
            ge = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ge, capture_error_mode='thread_local'):
                self.engine.sync_stream()
                plan.eval(x, g1, True)
                plan.copy_scalars_async(rows[k])
                opt._call('st2_lbfgs_commit', optimizers._p(g1), optimizers._p(g0))
                g0.copy_(g1)
            evals.append(ge)
        self.engine.sync_stream()
        captured = self.engine.launches() - l0
        per_step = captured // self.GRAPH_ROWS
        self.engine.graph_launches = getattr(self.engine, 'graph_launches', 0) - captured    # recorded, not run
        self._graphs = {'adv': adv, 'eval': evals, 'rows': rows, 'g0': g0, 'g1': g1, 'x': x, 'turn': 0, 'plan': plan,
                        'traces': [None] * self.GRAPH_ROWS, 'launches': per_step, 'step_size': opt.step_size}

    def _graph_step(self):
        """One L-BFGS step from the captured graphs; returns False when the eager path must run instead."""
        opt = self.optimizer
        if (not self.use_graphs or getattr(self.engine, 'profiling', False) or self._private_plans is not None
                or not isinstance(opt, optimizers.LBFGSOptimizer) or opt.loss is None):
            return False
        if self._graphs is None:
            if self._stable_steps < 3 or self._weights_dirty or self._pending_norms:
                return False
            try:
                self._graph_capture()
            except Exception:                  # never lose a job over the fast path: go on kernel by kernel
                logging.getLogger(__name__).exception('CUDA graph capture failed; continuing without graphs')
                self.use_graphs = False
                self._graphs = None
                self.engine.sync_stream()
                opt.objective_changed()        # libst2's step bookkeeping may be mid-sequence: start a clean history
                return False
        g = self._graphs
        if (g['plan'] is not self._plan or not g['plan'].handle or g['step_size'] != opt.step_size
                or opt.grad is not g['g0'] or g['x'] is not self.input or opt.x is not self.input):
            self._drop_graphs()
            return False
        k = g['turn']
        g['turn'] = (k + 1) % self.GRAPH_ROWS
        old = g['traces'][k]
        if old is not None:
            old.detach()                       # its row is about to be rewritten (GRAPH_ROWS steps later: long done)
        self.engine.sync_stream()
        g['adv'].replay()
        if opt.after_advance is not None:
            opt.after_advance(self.input)
        g['eval'][k].replay()
        self.engine.graph_launches = getattr(self.engine, 'graph_launches', 0) + g['launches']
        tr = LazyTrace(g['rows'][k], None, list(self._spec), True, time.perf_counter())
        tr._event = ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.engine.device))
        g['traces'][k] = tr
        self.traces.append(tr)
        self._norm_source = tr
        if len(self.traces) > self.TRACE_KEEP:
            del self.traces[:-self.TRACE_KEEP]
        opt.loss = LazyLoss(tr)
        return True

    def _advance(self):
        """One optimizer step (worker.py:303-310), from the captured graphs once the job is in steady state."""
        self.t += 1
        if not self._graph_step():
            self.optimizer.step()
            self._stable_steps += 1
        tr = self.traces[-1]
        tr('fevals', self.t)
        return self.input, tr

    def step(self, fetch=True):
        """worker.py:303-310.  ``fetch=False`` skips the device->host copy of the iterate and the
        trace (device-resident benchmarking)."""
        x, tr = self._advance()
        if not fetch:
            return None, None
        return self.image(x), tr.data

    def close(self):
        """Release the plans this job owns (private_plans=True)."""
        self._drop_graphs()
        if self._private_plans:
            for plan in self._private_plans.values():
                self.model.release_plan(plan)
            self._private_plans.clear()
        self._plan = None

    def step_async(self, download=True):
        """``step()`` without the host wait: returns an ``IterateHandle``.  Lets the caller overlap the
        iterate's trip to the host (12.6 MB at 1024^2) and its own pickling / sending with the next iteration
        (the reference pickles every iterate synchronously, worker.py:351-353).
        ``download=False``: the image is deprocessed but its copy is only enqueued by ``handle.download(after)``."""
        x, tr = self._advance()
        return self.image_async(x, tr, download)

    def write_trace(self, filename):
        df = pd.DataFrame(t.data for t in self.traces)
        df.index.name = 'step'
        df.to_csv(filename)


class Worker:
    """worker.py:318-409: bind PULL on ``worker_socket``, connect PUSH to ``app_socket``, announce
    ``WorkerReady``, then drain-all-messages-then-one-step until ``Shutdown``.

    Two optional ``config.ini`` keys put the tiling scheduler (``tiled.TiledTransfer``, BASELINE config 4) behind
    the same protocol; absent, the worker holds the canvas on one GPU exactly as before:
      * ``gpus = 0,1,2,3``  one canvas in row strips over these GPUs, one process per GPU (``main`` re-launches
        itself under ``torch.distributed.run``; NCCL for the sums, CUDA-IPC peer memory for the halo rows).  Rank 0
        owns the two sockets and tells the other ranks, once per loop turn, which messages arrived; every rank
        replays them on its strip, rank 0 assembles and sends the iterates.
      * ``tiles = P``  P strips inside this one process on one GPU (same kernels and flag protocol; what the
        single-GPU tests exercise).
    """

    def __init__(self, config, model=None):
        import torch.distributed as dist
        self.rank, self.world, self._ctl = 0, 1, None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
            self._ctl = dist.new_group(backend='gloo')       # control messages travel host-to-host, off the GPU streams
        self._zmq = None
        self.sock_in = self.sock_out = None
        if self.rank == 0:
            import zmq
            self._zmq = zmq
            self.ctx = zmq.Context.instance()
            self.sock_in = self.ctx.socket(zmq.PULL)
            self.sock_out = self.ctx.socket(zmq.PUSH)
            self.sock_in.bind(config['worker_socket'])
            self.sock_out.connect(config['app_socket'])
        else:
            self.ctx = None
            self.sock_out = _NullSocket()
        self.run_should_stop = False
        if model is None:
            base = utils.REPO_DIR
            gpu = config.getint('gpu', fallback=-1) if hasattr(config, 'getint') else int(config.get('gpu', -1))
            if self.world > 1:
                gpu = int(os.environ.get('LOCAL_RANK', self.rank))
            model = B200Model(base / config.get('prototxt', 'models/vgg19.prototxt'),
                              base / config.get('caffemodel', 'models/vgg19.caffemodel'), gpu,
                              precision=config.get('precision', None))
        tiles = int(config.get('tiles', 1) or 1)
        if self.world > 1 or tiles > 1:
            from .tiled import TiledTransfer
            self.transfer = TiledTransfer(model, local_world=None if self.world > 1 else tiles)
        else:
            self.transfer = StyleTransfer(model)
        # wire format of the iterates: the reference's send_pyobj pickles with the default protocol, so an app on
        # an older Python can still read them; `pickle_protocol` in config.ini overrides (5 saves one 12.6 MB copy)
        proto = config.get('pickle_protocol', None) if hasattr(config, 'get') else None
        self.pickle_protocol = int(proto) if proto not in (None, '') else pickle.DEFAULT_PROTOCOL
        self.sock_out.send_pyobj(WorkerReady(layers=self.transfer.model.layers()))

    # ---- message intake.  One process: straight from the socket.  One process per GPU: rank 0 reads the socket
    # and broadcasts what it got (possibly nothing) once per loop turn; every rank then does the same thing.
    def _poll(self, block):
        """Messages to handle now: everything queued (``block`` = False) or the next one (``block`` = True)."""
        msgs = []
        if self.rank == 0:
            zmq = self._zmq
            try:
                if block:
                    msgs.append(self.sock_in.recv_pyobj())
                while True:
                    msgs.append(self.sock_in.recv_pyobj(zmq.NOBLOCK))
            except zmq.ZMQError:
                pass
            except pickle.UnpicklingError:
                logger.error('Invalid message received over ZeroMQ.')
        if self.world > 1:
            import torch.distributed as dist
            box = [msgs]
            dist.broadcast_object_list(box, src=0, group=self._ctl)
            msgs = box[0]
        return msgs

    def run(self):
        try:
            while not self.run_should_stop:
                running = self.transfer.is_running
                for msg in self._poll(block=not running):
                    if self.process_message(msg):
                        self.run_should_stop = True
                        break
                if self.run_should_stop:
                    break
                if running and self.transfer.is_running:
                    if self.transfer.check_consistency():
                        # one iterate stays in flight: iteration t+1 is enqueued before iterate t is
                        # waited for, pickled and sent, so the transport overlaps the compute
                        handle = self.transfer.step_async()
                        self._flush()
                        self._pending = handle
                    else:
                        self._flush()
                        self.sock_out.send_pyobj(GetImages())
                if not self.transfer.is_running:
                    self._flush()
        except KeyboardInterrupt:
            pass
        finally:
            try:
                self._flush()
            finally:
                self.sock_out.send_pyobj(Shutdown())

    _pending = None

    def _flush(self):
        """Send the iterate that is still in flight, if any (worker.py:351-353)."""
        handle, self._pending = self._pending, None
        if handle is not None:
            image, trace = handle.result()
            if image is None:                # a rank that does not assemble iterates (row strips, rank > 0)
                return
            # `image` is a view of a pinned double buffer: pickling copies it out right here, before the buffer can
            # be reused, so no intermediate np.array() copy; one frame, as recv_pyobj on the app side expects
            # (send_pyobj would pickle with the default protocol and then copy the 12.6 MB frame once more).
            frame = pickle.dumps(Iterate(image, handle.t, dict(trace)), protocol=self.pickle_protocol)
            self.sock_out.send(frame, copy=False)

    def process_message(self, msg):
        """worker.py:366-409.  Returns True when the loop should end."""
        def is_image(obj):
            return obj is not None and not isinstance(obj, int)

        tr = self.transfer
        if isinstance(msg, SetImages):
            if is_image(msg.input_image):
                tr.set_input(msg.input_image)
            elif msg.input_image == SetImages.RESAMPLE:
                tr.resample_input(msg.size)
            if is_image(msg.content_image):
                tr.set_content(msg.content_image)
            elif msg.content_image == SetImages.RESAMPLE:
                tr.resample_content(msg.size)
            if is_image(msg.style_image):
                tr.set_style(msg.style_image)
            if msg.reset_state:
                tr.reset()
        elif isinstance(msg, SetOptimizer):
            tr.set_optimizer_class(SetOptimizer.classes[msg.optimizer], msg.step_size)
        elif isinstance(msg, SetWeights):
            tr.set_weights(msg.weights, msg.params)
        elif isinstance(msg, Shutdown):
            return True
        elif isinstance(msg, StartIteration):
            if not tr.start():
                self.sock_out.send_pyobj(GetImages())
        elif isinstance(msg, PauseIteration):
            tr.pause()
        else:
            logger.error('Invalid message received over ZeroMQ.')
        return False


class _NullSocket:
    """Outbound socket of the ranks that do not talk to the app."""

    def send_pyobj(self, obj):
        pass

    def send(self, *args, **kwargs):
        pass


def _gpu_list(config):
    raw = (config.get('gpus', '') or '').replace(' ', '')
    return [int(v) for v in raw.split(',') if v != '']


def main():
    """worker.py:412-428."""
    messages.install_as_toplevel()
    args = utils.parse_args(__doc__)
    config = utils.read_config(args)
    gpus = _gpu_list(config)
    if len(gpus) > 1 and 'WORLD_SIZE' not in os.environ:
        # `gpus = 0,1,...`: the same command line, one process per GPU (the app starts `worker.py` as one process,
        # app.py:336-344); rendezvous on the loopback interface
        import socket
        s = socket.socket()
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
        s.close()
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=','.join(str(g) for g in gpus))
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(len(gpus)),
               '--master-addr', '127.0.0.1', '--master-port', str(port), '-m', 'style_transfer2_b200.worker'] + sys.argv[1:]
        os.execvpe(sys.executable, cmd, env)
    if int(os.environ.get('WORLD_SIZE', 1)) > 1:
        import torch.distributed as dist
        local = int(os.environ.get('LOCAL_RANK', 0))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    debug = args.debug + config.getint('debug', 0)
    utils.setup_logging(debug)
    utils.setup_signals()
    worker = None
    try:
        worker = Worker(config)
        worker.run()
    finally:
        logger.info('Shutting down worker process.')
        if worker is not None and worker.ctx is not None:
            worker.ctx.destroy(0)


if __name__ == '__main__':
    main()
