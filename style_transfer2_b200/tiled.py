"""One large canvas split into row strips over several B200s (BASELINE config 4).

The reference holds the whole image in one ``caffe.Net`` (worker.py:84-86) and caps its size
(``max_size``, app.py:183-185); this module is the tiling scheduler that replaces that cap.  ``TiledTransfer``
is ``worker.StyleTransfer`` (worker.py:117-315) -- the whole state machine the worker drives:
``set_input / resample_input / set_content / resample_content / set_style / set_weights /
set_optimizer_class / reset / start / pause / check_consistency / opfunc / step / step_async`` -- for a
canvas whose rows are partitioned over ``world`` strips (``Worker`` builds it when ``config.ini`` has
``gpus`` / ``tiles``):

* every strip owns rows ``[row0, row1)`` (boundaries at multiples of 32 rows, ``parallel.strip_bounds``)
  of x, of the gradient, of the L-BFGS history / Adam moments and of every activation;
* before each 3x3 convolution (forward and data-gradient) the strips swap ONE boundary row.  libst2
  does that itself through peer memory (CUDA IPC mapping -> NVLink): with one process per GPU the push
  and the wait live INSIDE the convolution kernel that consumes the tensor; strips sharing a process
  use a small exchange kernel.  No NCCL, no host synchronisation on the halo path;
* what remains are small sum all-reduces, issued here between the phases of the evaluation: the
  strips' Gram sums (<= 2.4 MB), 3 sums per weighted blob, 6 pixel-space sums, and the L-BFGS
  dot-product block (one block per step thanks to the compact form, st2_lbfgs.cu) -- four collectives
  in general, TWO per iteration once the normalisers are frozen (``opfunc``).

Two ways to place the strips:
* ``torch.distributed`` (one process per GPU, NCCL): each process holds the strip of its rank;
* ``local_world=P``: P strips inside one process on one GPU, one CUDA stream each -- the same kernels
  and the same flag protocol, used by the single-GPU parity tests.

PyTorch supplies device memory, streams and the NCCL process group; nothing else.
"""
from collections import OrderedDict
import ctypes as C
import time

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

from . import _lib, optimizers, parallel, utils, vgg
from .messages import SetOptimizer, SetWeights
from .model import Plan, _ptr
from .worker import EPS_W, LazyLoss, LazyTrace


class _DevView:
    """Zero-copy torch view of library-owned device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {'shape': (int(count),), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 2, 'strides': None}


def dev_tensor(ptr, count, dtype, device):
    typestr = {torch.float32: '<f4', torch.float64: '<f8'}[dtype]
    if count == 0:
        return torch.empty(0, dtype=dtype, device=device)
    return torch.as_tensor(_DevView(ptr, count, typestr), device=device)


class StripPlan(Plan):
    """``model.Plan`` for one row strip (st2_strip_plan_create)."""

    def __init__(self, engine, height_total, width, row0, row1, rank, world, precision):
        self.engine, self.W = engine, int(width)
        self.H = int(row1 - row0)
        self.H_total, self.row0, self.row1, self.rank, self.world = int(height_total), int(row0), int(row1), rank, world
        self.lib = engine.lib
        handle = C.c_void_p()
        _lib.check(engine.ctx, self.lib.st2_strip_plan_create(engine.ctx, self.H_total, self.W, self.row0, self.row1,
                                                              rank, world, precision, C.byref(handle)),
                   'st2_strip_plan_create')
        self.handle = handle
        self._scal_host = (C.c_double * _lib.SCAL_TOTAL)()

    def ipc_handle(self):
        buf = (C.c_ubyte * _lib.IPC_HANDLE_BYTES)()
        self._check(self.lib.st2_strip_ipc_handle(self.handle, buf), 'st2_strip_ipc_handle')
        return bytes(buf)

    def attach(self, side, peer_rows, ipc_handle=None, local_peer=None):
        hb = (C.c_ubyte * _lib.IPC_HANDLE_BYTES).from_buffer_copy(ipc_handle) if ipc_handle is not None else None
        self._check(self.lib.st2_strip_attach(self.handle, side, hb, local_peer.handle if local_peer else None,
                                              int(peer_rows)), 'st2_strip_attach')

    def reduce_block(self, which):
        ptr, n = C.c_void_p(), C.c_longlong()
        self._check(self.lib.st2_strip_reduce_block(self.handle, which, C.byref(ptr), C.byref(n)),
                    'st2_strip_reduce_block')
        return dev_tensor(ptr.value or 0, n.value, torch.float32 if which == 0 else torch.float64, self.engine.device)

    def eval_begin(self, x, want_grad):
        self._check(self.lib.st2_eval_begin(self.handle, _ptr(x), 1 if want_grad else 0), 'st2_eval_begin')

    def eval_mid(self):
        self._check(self.lib.st2_eval_mid(self.handle), 'st2_eval_mid')

    def eval_end(self, grad):
        self._check(self.lib.st2_eval_end(self.handle, _ptr(grad)), 'st2_eval_end')

    def eval_final(self):
        self._check(self.lib.st2_eval_final(self.handle), 'st2_eval_final')

    def set_deferred(self, enable):
        self._check(self.lib.st2_strip_set_deferred(self.handle, 1 if enable else 0), 'st2_strip_set_deferred')

    def set_fold(self, enable):
        self._check(self.lib.st2_strip_set_fold(self.handle, 1 if enable else 0), 'st2_strip_set_fold')

    def halo_error(self):
        err = C.c_int()
        self._check(self.lib.st2_strip_halo_error(self.handle, C.byref(err)), 'st2_strip_halo_error')
        return err.value


class _Strip:
    """Per-strip state held by this process."""

    def __init__(self, rank, row0, row1):
        self.rank, self.row0, self.row1 = rank, row0, row1
        self.plan = None
        self.stream = None
        self.x = None
        self.content = None
        self.grads = [None, None]
        self.turn = 0
        self.opt = None                # st2_lbfgs handle
        self.m1 = self.m2 = None       # Adam moments
        self.grad = None               # gradient at x owned by the optimizer

    @property
    def rows(self):
        return self.row1 - self.row0


class TiledIterate:
    """An iterate of a tiled canvas on its way to the host (``TiledTransfer.step_async``): same surface as
    ``worker.IterateHandle``.  Only the rank that assembles iterates (rank 0) gets an image; the others get None."""

    def __init__(self, host, done, trace, t, check):
        self._host, self._done, self.trace, self.t, self._check = host, done, trace, t, check

    def result(self):
        data = self.trace.data                  # waits for this evaluation's scalar block
        self._check(self.trace)
        if self._host is None:
            return None, data
        self._done.synchronize()
        return self._host.numpy(), data


class TiledTransfer:
    """Row-tiled ``StyleTransfer`` (worker.py:117-315): the same state machine and message-facing methods, for a
    canvas whose rows are partitioned over ``world`` strips.  All ranks call every method in the same order (SPMD);
    in a one-process-per-GPU worker rank 0 owns the sockets and every rank replays the same messages."""

    def __init__(self, model, height=None, width=None, local_world=None, optimizer='lbfgs', step_size=None):
        self.model, self.engine = model, model.engine
        self.dist = local_world is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if local_world is not None:
            # every strip spins on its neighbours' flags from its own stream: more streams than hardware queues
            # (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default) could serialise a wait ahead of the push it waits for
            if int(local_world) > 8:
                raise ValueError('local_world is limited to 8 strips per process')
            self.world, self._ranks = int(local_world), list(range(int(local_world)))
        elif self.dist:
            self.world, self._ranks = dist.get_world_size(), [dist.get_rank()]
        else:
            self.world, self._ranks = 1, [0]
        self.rank0 = self._ranks[0] == 0
        self.H = self.W = None
        self.bounds, self.strips = [], []
        self.weights = pd.DataFrame(np.ones((len(vgg.BLOBS), len(SetWeights.loss_names))), list(vgg.BLOBS),
                                    SetWeights.loss_names, np.float32)
        self.params = {w: 1 for w in SetWeights.scalar_loss_names}
        self.optimizer_name = optimizer
        self.step_size = SetOptimizer.step_sizes[optimizer] if step_size is None else step_size
        self.b1, self.b2 = 0.9, 0.999
        self.t = 0
        self.traces = []
        self.style = None
        self.content = None            # the whole preprocessed content image (1, 3, Hc, Wc): strips slice it lazily
        self.is_running = self.is_starting = False
        self._have_opt = False         # an optimizer state exists (the reference's `self.optimizer is not None`)
        self._spec = []
        self._weights_dirty = True
        self._content_done, self._style_done = set(), set()
        self._have_eval = False
        self._cold = True
        self._adam_items = self._adam_items2 = 0
        self._pending_norms = {}
        self._norm_source = None
        self._dl = None
        self.loss = None
        # every active normaliser is frozen (an evaluation with the current weights has completed): the per-blob sums
        # are then needed for trace values only and their all-reduce merges with the later ones (st2_strip_set_deferred)
        self._norms_frozen = False
        self._deferred_on = None
        if height is not None:
            self._layout(int(height), int(width))

    # ------------------------------------------------------------------ plumbing
    def _layout(self, height, width):
        """(Re)build the strips for a canvas size (the reference reshapes its net on demand, worker.py:84)."""
        self._release_strips()
        self.H, self.W = int(height), int(width)
        bounds = parallel.strip_bounds(self.H, self.world)
        if any(e <= s for s, e in bounds):
            raise ValueError('canvas of %d rows is too small for %d strips of >= 32 rows' % (self.H, self.world))
        self.bounds = bounds
        dev = self.engine.device
        main = torch.cuda.current_stream(dev)
        for r in self._ranks:
            st = _Strip(r, *bounds[r])
            st.stream = main if len(self._ranks) == 1 else torch.cuda.Stream(dev)
            with torch.cuda.stream(st.stream):
                self.engine.sync_stream()
                st.plan = StripPlan(self.engine, self.H, self.W, st.row0, st.row1, r, self.world, self.model.precision)
            self.strips.append(st)
        self.engine.sync_stream()
        torch.cuda.synchronize(dev)
        # conv1_1's style gradient is folded into its data-gradient weights only when EVERY strip can (>= 16 rows):
        # the strips read a row of each other's grad(conv1_1), which means something else with the fold
        fold = self.model.precision_name == 'fp16' and self.W >= 16 and min(e - s for s, e in bounds) >= 16
        for st in self.strips:
            st.plan.set_fold(fold)
        self._attach()
        self._weights_dirty = True
        self._content_done, self._style_done = set(), set()
        self._have_eval = False
        self._dl = None
        self._norms_frozen = False
        self._deferred_on = None

    def _release_strips(self):
        if self.strips:
            torch.cuda.synchronize(self.engine.device)
            if self.dist:
                dist.barrier()                   # nobody may still be pushing halo rows into a slab that is about to go
        for st in self.strips:
            if st.opt is not None:
                self.engine.lib.st2_lbfgs_destroy(st.opt)
                st.opt = None
            if st.plan is not None:
                st.plan.close()
                st.plan = None
        self.strips = []

    def _each(self):
        """Iterate the local strips with their stream current and libst2 pointed at it."""
        for st in self.strips:
            with torch.cuda.stream(st.stream):
                self.engine.sync_stream()
                yield st
        self.engine.sync_stream()

    def _attach(self):
        """Circular neighbours: side 0 = the strip above (rank-1), side 1 = the strip below."""
        rows = [e - s for s, e in self.bounds]
        if self.dist:
            me = self.strips[0]
            handles = [None] * self.world
            dist.all_gather_object(handles, me.plan.ipc_handle())
            for side, peer in ((0, (me.rank - 1) % self.world), (1, (me.rank + 1) % self.world)):
                if peer == me.rank:
                    me.plan.attach(side, rows[peer], local_peer=me.plan)
                else:
                    me.plan.attach(side, rows[peer], ipc_handle=handles[peer])
            dist.barrier()
        else:
            by_rank = {st.rank: st for st in self.strips}
            for st in self.strips:
                for side, peer in ((0, (st.rank - 1) % self.world), (1, (st.rank + 1) % self.world)):
                    st.plan.attach(side, rows[peer], local_peer=by_rank[peer].plan)

    def all_reduce(self, tensors):
        """Sum ``tensors[i]`` (one per local strip) over all strips, in place, stream-ordered."""
        if not tensors or tensors[0].numel() == 0:
            return
        prof = getattr(self, 'allreduce_spans', None)      # bench.py: list of (start, end) events on strip 0's stream
        if prof is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record(self.strips[0].stream)
        self._all_reduce(tensors)
        if prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(self.strips[0].stream)
            prof.append((e0, e1))

    def _all_reduce(self, tensors):
        if len(self.strips) > 1:
            s0 = self.strips[0].stream
            for st in self.strips[1:]:
                s0.wait_event(st.stream.record_event())
            with torch.cuda.stream(s0):
                for t in tensors[1:]:
                    tensors[0].add_(t)
            done = s0.record_event()
            for st, t in zip(self.strips[1:], tensors[1:]):
                st.stream.wait_event(done)
                with torch.cuda.stream(st.stream):
                    t.copy_(tensors[0])
                # strip 0 goes on to overwrite its block (the next phase zeroes / refills it): not before every
                # other strip has taken its copy
                s0.wait_event(st.stream.record_event())
        if self.dist:
            with torch.cuda.stream(self.strips[0].stream):
                dist.all_reduce(tensors[0])

    def _rows_of(self, full, st):
        """Rows of a (1, 3, H, W) device tensor that belong to strip ``st`` (contiguous copy)."""
        return full[:, :, st.row0:st.row1, :].contiguous()

    def _scatter(self, full, attr):
        """Give every local strip its rows of a whole-canvas tensor (stored as ``strip.<attr>``)."""
        torch.cuda.current_stream(self.engine.device).synchronize()
        for st in self._each():
            setattr(st, attr, self._rows_of(full, st))

    def _upload(self, image):
        arr = np.asarray(image)
        h, w = arr.shape[:2]
        out = self.engine.empty(1, 3, h, w)
        dev = torch.from_numpy(np.array(arr, copy=True)).to(self.engine.device)
        fn = 'st2_preprocess_u8' if arr.dtype == np.uint8 else 'st2_preprocess_f32'
        if arr.dtype != np.uint8:
            dev = dev.float().contiguous()
        self.engine.sync_stream()
        self.engine.call(fn, C.c_void_p(dev.data_ptr()), C.c_void_p(out.data_ptr()), h, w)
        return out

    # ------------------------------------------------------------------ state machine (worker.py:140-189)
    @property
    def input_shape(self):
        return None if self.H is None or not self.strips or self.strips[0].x is None else (1, 3, self.H, self.W)

    @property
    def grams(self):
        return {'style': self.style} if self.style is not None else None

    @property
    def features(self):
        return {'content': self.content} if self.content is not None else None

    def check_consistency(self):
        return bool(self.input_shape is not None and self.content is not None and self.grams
                    and tuple(self.input_shape) == tuple(self.content.shape))

    def pause(self):
        self.is_running = False
        self.is_starting = False

    def start(self):
        self.is_starting = True
        self._start()
        return self.is_running

    def _start(self):
        if self.is_starting and self.check_consistency():
            if not self._have_opt:
                self.reset()
            self.is_starting = False
            self.is_running = True

    # ------------------------------------------------------------------ job set-up (worker.py:154-229)
    def set_input(self, image):
        """worker.py:191-202."""
        full = self._upload(image)
        shape = tuple(full.shape[2:])
        if self.input_shape is not None and shape == (self.H, self.W):
            torch.cuda.current_stream(self.engine.device).synchronize()
            for st in self._each():
                st.x.copy_(full[:, :, st.row0:st.row1, :])
            self.objective_changed()
        elif self._have_opt:
            self._adopt(full)                    # optimizer.resample(None, new_x=image)
            self._start()
        else:
            if (self.H, self.W) != shape or not self.strips:
                self._relayout(*shape)
            self._scatter(full, 'x')
            self.reset()
            self._start()

    def set_content(self, image):
        """worker.py:204-209.  The whole content image is kept; strips take their rows when the canvas matches."""
        self.content = self._upload(image)
        self._content_done = set()
        self._start()
        self.objective_changed()

    def set_style(self, image):
        self.style = self._upload(image)
        self._style_done = set()
        self._start()
        self.objective_changed()

    def resample_input(self, size):
        """worker.py:154-160."""
        size = tuple(int(v) for v in size)
        if self.input_shape is not None and self._have_opt:
            self._adopt(None, size)
        else:
            self._relayout(*size)
            self._scatter(self.engine.zeros(1, 3, *size), 'x')
        self._start()
        self.objective_changed()

    def resample_content(self, size):
        """worker.py:162-170."""
        size = tuple(int(v) for v in size)
        utils.set_default_engine(self.engine)
        self.engine.sync_stream()
        if self.content is not None:
            self.content = utils.resample_nchw(self.content, size)
        else:
            self.content = self.engine.zeros(1, 3, *size)
        self._content_done = set()
        self._start()
        self.objective_changed()

    def set_step_size(self, step_size):
        self.step_size = step_size

    def set_optimizer_class(self, cls, step_size):
        """worker.py:387-391 (SetOptimizer): switch class / step size; a class change resets the job state."""
        name = 'adam' if cls is optimizers.AdamOptimizer or getattr(cls, '__name__', '') == 'AdamOptimizer' else 'lbfgs'
        self.set_step_size(step_size)
        if name != self.optimizer_name:
            self.optimizer_name = name
            if self.input_shape is not None:
                self.reset()
            else:
                self._have_opt = False

    @property
    def optimizer_cls(self):
        return optimizers.AdamOptimizer if self.optimizer_name == 'adam' else optimizers.LBFGSOptimizer

    def set_weights(self, weights, params):
        self.weights = pd.DataFrame.from_dict(weights, dtype=np.float32)
        self.params = params
        self._weights_dirty = True
        self._norms_frozen = False               # a newly weighted layer freezes its normaliser at the next evaluation
        self.objective_changed()

    def objective_changed(self):
        """optimizers.py:42-46 / 121-125."""
        self._have_eval = False
        self.loss = None
        if self.optimizer_name == 'lbfgs':
            for st in self._each():
                if st.opt is not None:
                    self.engine.lib.st2_lbfgs_reset(st.opt)
            self._cold = True
        else:
            self._adam_items = 0
            for st in self.strips:
                if st.m1 is not None:
                    st.m1.zero_()

    def _new_optimizer_state(self):
        for st in self._each():
            if st.opt is not None:
                self.engine.lib.st2_lbfgs_destroy(st.opt)
                st.opt = None
            st.m1 = st.m2 = None
            if self.optimizer_name == 'lbfgs':
                h = C.c_void_p()
                self.engine.call('st2_lbfgs_create', st.x.numel(), 10, C.byref(h))
                _lib.check(self.engine.ctx, self.engine.lib.st2_lbfgs_set_global_length(h, float(3 * self.H * self.W)),
                           'st2_lbfgs_set_global_length')
                st.opt = h
            else:
                st.m1 = torch.zeros_like(st.x)
                st.m2 = torch.zeros_like(st.x)
        self._cold = True
        self._have_eval = False
        self._have_opt = True

    def reset(self):
        """worker.py:172-175: new norms, t = 0, fresh optimizer state."""
        for st in self._each():
            st.plan.reset_norms()
        self._pending_norms = {}
        self._norm_source = None
        self._norms_frozen = False
        self._new_optimizer_state()
        self._adam_items = self._adam_items2 = 0
        self.t = 0

    # ------------------------------------------------------------------ scale changes (cold path)
    def _whole(self, attr):
        """The whole-canvas tensor assembled from ``strip.<attr>`` on EVERY rank (cold path: scale changes)."""
        main = torch.cuda.current_stream(self.engine.device)
        full = self.engine.zeros(1, 3, self.H, self.W)
        for st in self.strips:
            main.wait_event(st.stream.record_event())
            full[:, :, st.row0:st.row1, :] = getattr(st, attr)
        if self.dist:
            dist.all_reduce(full)            # zero-filled sum = all-gather of ragged strips; cold path only
        return full

    def _relayout(self, height, width):
        """New strips for a new canvas size; frozen normalisers survive the change (they are job state, worker.py:172)."""
        if self._norm_source is not None:
            for kind, table in self.norms.items():
                for layer, v in table.items():
                    self._pending_norms.setdefault((kind, layer), v)
        self._norm_source = None
        self._layout(height, width)

    def _adopt(self, new_x, size=None):
        """``optimizer.resample(size, new_x)`` (optimizers.py:29-40, 110-119) on sharded state: x is replaced by
        ``new_x`` or Lanczos-resampled; Adam's first moment Lanczos, second moment bilinear then max(0, .);
        L-BFGS drops its history."""
        utils.set_default_engine(self.engine)
        self.engine.sync_stream()
        adam = self.optimizer_name == 'adam'
        m1 = m2 = None
        if new_x is None:
            new_x = utils.resample_nchw(self._whole('x'), size)
        else:
            size = tuple(new_x.shape[2:])
        if adam:
            m2 = utils.resample_nchw(self._whole('m2'), size, method=utils.BILINEAR, clamp_min_zero=True)
            m1 = utils.resample_nchw(self._whole('m1'), size) if self._adam_items else None
        items = (self._adam_items, self._adam_items2)
        self._relayout(*size)
        self._scatter(new_x, 'x')
        self._new_optimizer_state()
        if adam:
            self._scatter(m2, 'm2')
            if m1 is not None:
                self._scatter(m1, 'm1')
            self._adam_items, self._adam_items2 = items
        else:
            self.objective_changed()

    # ------------------------------------------------------------------ norms (checkpointing, scale changes)
    @property
    def norms(self):
        out = self._norm_source.norms() if self._norm_source is not None else {k: {} for k in 'cds'}
        for (kind, layer), v in self._pending_norms.items():
            out[kind][layer] = v
        return out

    def set_norms(self, norms):
        for kind, table in norms.items():
            for layer, v in table.items():
                self._pending_norms[(kind, layer)] = float(v)

    def active_layers(self):
        nonzeros = abs(self.weights) > EPS_W
        return list(self.weights.index[abs(nonzeros.sum(axis=1)) > EPS_W])

    def _sync_plans(self):
        if self._weights_dirty:
            spec, order = [], []
            table = self.weights
            rows = []
            for name in self.active_layers():
                if name not in vgg.BLOB_INDEX:
                    raise KeyError('unknown layer %r' % name)
                b = vgg.BLOB_INDEX[name]
                vals = [float(table[col][name]) if col in table.columns else 0.0 for col in SetWeights.loss_names]
                vals = [0.0 if (np.isnan(v) or abs(v) <= EPS_W) else v for v in vals]
                rows.append((b, vals))
                order.append(b)
                spec.append((b, vals[0] != 0.0, vals[1] != 0.0, vals[2] != 0.0))
            for st in self._each():
                for b in range(_lib.NUM_BLOBS):
                    st.plan.set_blob_weights(b, 0.0, 0.0, 0.0)
                for b, vals in rows:
                    st.plan.set_blob_weights(b, *vals)
                st.plan.set_eval_order(order)
                st.plan.set_params(float(self.params['tv']), float(self.params['tv_power']), float(self.params['p']),
                                   float(self.params['p_power']))
            self._spec = spec
            self._weights_dirty = False
        if self._pending_norms:
            for st in self._each():
                for (kind, layer), v in self._pending_norms.items():
                    st.plan.set_norm(kind, vgg.BLOB_INDEX[layer], v)
            self._pending_norms = {}
        need_c = [b for b, c_on, _, _ in self._spec if c_on and b not in self._content_done]
        if need_c:
            if tuple(self.content.shape[2:]) != (self.H, self.W):
                raise RuntimeError('content is %s, canvas is %s' % (tuple(self.content.shape[2:]), (self.H, self.W)))
            torch.cuda.current_stream(self.engine.device).synchronize()
            for st in self._each():
                st.content = self._rows_of(self.content, st)
                st.plan.forward(st.content, max(need_c))
                for b in need_c:
                    st.plan.capture_content(b)
            self._content_done.update(need_c)
        need_s = [b for b, _, s_on, _ in self._spec if s_on and b not in self._style_done]
        if need_s:
            # the style image has its own size: every process computes its Gram targets on a whole-canvas plan
            hs, ws = self.style.shape[2:]
            self.engine.sync_stream()
            sp = Plan(self.engine, hs, ws, self.model.precision)
            try:
                sp.forward(self.style, max(need_s))
                grams = {b: sp.gram(b) for b in need_s}
                torch.cuda.current_stream(self.engine.device).synchronize()
            finally:
                sp.close()
            for st in self._each():
                for b in need_s:
                    st.plan.set_style_gram(b, grams[b])
            self._style_done.update(need_s)

    # ------------------------------------------------------------------ objective (worker.py:231-301)
    def _all_reduce_merged(self, blocks):
        """One all-reduce for several small blocks: ``blocks`` = list of per-strip tensor lists (same dtype)."""
        cats = []
        for i, st in enumerate(self._each()):
            cats.append(torch.cat([b[i] for b in blocks]))
        self.all_reduce(cats)
        for i, st in enumerate(self._each()):
            off = 0
            for b in blocks:
                n = b[i].numel()
                b[i].copy_(cats[i][off:off + n])
                off += n

    def opfunc(self, return_grad=True, extra_sums=None):
        """Collective evaluation at the strips' current x.  Returns (loss, [grad per local strip]).

        Four sum all-reduces in general (Gram sums; per-blob sums; pixel sums; the optimizer's own afterwards).  Once
        every active normaliser is frozen the per-blob sums only feed trace values, so they travel with the pixel sums
        after the backward pass -- and with ``extra_sums(grads)`` (the L-BFGS dot-product block, computed from the new
        gradient): two all-reduces per iteration.  ``self.merged_extra`` tells the caller whether that happened."""
        self._sync_plans()
        deferred = bool(return_grad and self._norms_frozen)
        if deferred != self._deferred_on:
            for st in self.strips:
                st.plan.set_deferred(deferred)
            self._deferred_on = deferred
        grads = []
        for st in self._each():
            if return_grad:
                st.turn ^= 1
                if st.grads[st.turn] is None or st.grads[st.turn].shape != st.x.shape:
                    st.grads[st.turn] = torch.empty_like(st.x)
                grads.append(st.grads[st.turn])
            st.plan.eval_begin(st.x, return_grad)
        self.all_reduce([st.plan.reduce_block(0) for st in self.strips])
        for st in self._each():
            st.plan.eval_mid()
        if not deferred:
            self.all_reduce([st.plan.reduce_block(1) for st in self.strips])
        for st, g in zip(self._each(), grads if return_grad else [None] * len(self.strips)):
            st.plan.eval_end(g)
        self.merged_extra = False
        if deferred:
            blocks = [[st.plan.reduce_block(1) for st in self.strips], [st.plan.reduce_block(2) for st in self.strips]]
            if extra_sums is not None:
                blocks.append(extra_sums(grads))
                self.merged_extra = True
            self._all_reduce_merged(blocks)
        else:
            self.all_reduce([st.plan.reduce_block(2) for st in self.strips])
        for st in self._each():
            st.plan.eval_final()
        self._norms_frozen = True
        st0 = self.strips[0]
        tr = LazyTrace(None, None, list(self._spec), return_grad, time.perf_counter())
        tr._host = host = self.engine.trace_slot(tr)
        with torch.cuda.stream(st0.stream):
            self.engine.sync_stream()
            st0.plan.copy_scalars_async(host)
            tr._event = ev = torch.cuda.Event()
            ev.record(st0.stream)
        self.engine.sync_stream()
        tr.halo_timeout = lambda tr=tr: bool(tr._host[_lib.SCAL_GLOBAL_BASE + _lib.G_HALO_TIMEOUT] != 0.0 or
                                             tr._host[_lib.SCAL_GLOBAL_BASE + _lib.G_PROTOCOL_ERROR] != 0.0)
        self.traces.append(tr)
        self._norm_source = tr
        del self.traces[:-256]
        return (LazyLoss(tr), grads) if return_grad else LazyLoss(tr)

    # ------------------------------------------------------------------ optimizers (optimizers.py)
    def _lbfgs_sums(self):
        n = self.engine.lib.st2_lbfgs_sums_count()
        return [dev_tensor(self.engine.lib.st2_lbfgs_sums_dev(st.opt), n, torch.float64, self.engine.device)
                for st in self.strips]

    def _lcall(self, st, name, *args):
        _lib.check(self.engine.ctx, getattr(self.engine.lib, name)(st.opt, *args), name)

    def _step_lbfgs(self):
        """optimizers.py:62-77 on sharded vectors: one all-reduce of the dot-product block per step."""
        if not self._have_eval:
            self.loss, grads = self.opfunc()
            for st, g in zip(self.strips, grads):
                st.grad = g
            self._have_eval = True
        for st in self._each():
            self._lcall(st, 'st2_lbfgs_advance_begin', _ptr(st.grad))
        if self._cold:
            self.all_reduce(self._lbfgs_sums())
        for st in self._each():
            self._lcall(st, 'st2_lbfgs_advance_end', _ptr(st.x), _ptr(st.grad), float(self.step_size))
        def commit_begin(grads):
            for st, g in zip(self._each(), grads):
                self._lcall(st, 'st2_lbfgs_commit_begin', _ptr(g), _ptr(st.grad))
            return self._lbfgs_sums()

        loss, grads = self.opfunc(extra_sums=commit_begin)     # steady state: the dot products ride with the loss sums
        if not self.merged_extra:
            self.all_reduce(commit_begin(grads))
        for st, g in zip(self._each(), grads):
            self._lcall(st, 'st2_lbfgs_commit_end')
            st.grad = g
        self._cold = False
        self.loss = loss
        return loss

    def _step_adam(self):
        """optimizers.py:20-27: purely element-wise, no reduction."""
        loss, grads = self.opfunc()
        self._adam_items += 1
        self._adam_items2 += 1
        for st, g in zip(self._each(), grads):
            self.engine.call('st2_adam_step', _ptr(st.x), _ptr(g), _ptr(st.m1), _ptr(st.m2), st.x.numel(),
                             float(self.step_size), self.b1, self.b2, self._adam_items, self._adam_items2)
        self.loss = loss
        return loss

    def _advance(self):
        if not self._have_opt:
            self.reset()
        self.t += 1
        loss = self._step_lbfgs() if self.optimizer_name == 'lbfgs' else self._step_adam()
        tr = self.traces[-1]
        tr('fevals', self.t)
        return tr

    def _raise_on_halo_timeout(self, tr):
        if tr.halo_timeout():
            raise RuntimeError('a halo exchange timed out (a neighbouring strip stopped) or the deferred-sums protocol was '
                               'used before the normalisers were frozen (rank %d)' % self.strips[0].rank)

    def step(self, fetch=True):
        """worker.py:303-310."""
        tr = self._advance()
        if not fetch:
            return None, None
        data = tr.data                       # waits for this evaluation's scalar block
        self._raise_on_halo_timeout(tr)
        return self.image(), data

    def step_async(self):
        """``step()`` without the host wait (worker.StyleTransfer.step_async): the iterate is assembled on rank 0,
        deprocessed and copied to a pinned double buffer on a side stream while the next iteration computes."""
        tr = self._advance()
        dev = self.engine.device
        main = torch.cuda.current_stream(dev)
        x = self.gather([st.x for st in self.strips])
        if x is None:
            return TiledIterate(None, None, tr, self.t, self._raise_on_halo_timeout)
        if self._dl is None or self._dl['shape'] != (self.H, self.W):
            self._dl = {'shape': (self.H, self.W), 'stream': torch.cuda.Stream(dev), 'turn': 0,
                        'dev': [self.engine.empty(self.H, self.W, 3) for _ in range(2)],
                        'host': [torch.empty((self.H, self.W, 3), dtype=torch.float32, pin_memory=True) for _ in range(2)],
                        'done': [None, None]}
        d = self._dl
        d['turn'] ^= 1
        k = d['turn']
        if d['done'][k] is not None:
            main.wait_event(d['done'][k])
        self.engine.sync_stream()
        self.engine.call('st2_deprocess', C.c_void_p(x.data_ptr()), C.c_void_p(d['dev'][k].data_ptr()), self.H, self.W)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(d['stream']):
            d['stream'].wait_event(ready)
            d['host'][k].copy_(d['dev'][k], non_blocking=True)
            done = torch.cuda.Event()
            done.record(d['stream'])
        d['done'][k] = done
        return TiledIterate(d['host'][k], done, tr, self.t, self._raise_on_halo_timeout)

    # ------------------------------------------------------------------ results
    def gather(self, per_strip, dst=0):
        """Full (1, 3, H, W) device tensor from per-strip row tensors.  One process: assembled locally.  One
        process per GPU: a gather to rank ``dst`` (every other rank sends its rows once and returns None)."""
        main = torch.cuda.current_stream(self.engine.device)
        for st in self.strips:
            main.wait_event(st.stream.record_event())
        if not self.dist:
            full = self.engine.empty(1, 3, self.H, self.W)
            for st, t in zip(self.strips, per_strip):
                full[:, :, st.row0:st.row1, :] = t
            return full
        me = self.strips[0]
        mine = per_strip[0].contiguous()
        if me.rank != dst:
            dist.send(mine, dst)
            return None
        full = self.engine.empty(1, 3, self.H, self.W)
        full[:, :, me.row0:me.row1, :] = mine
        for r, (r0, r1) in enumerate(self.bounds):
            if r == dst:
                continue
            part = self.engine.empty(1, 3, r1 - r0, self.W)
            dist.recv(part, r)
            full[:, :, r0:r1, :] = part
        return full

    def image(self):
        """Deprocessed iterate, HxWx3 fp32 host array (CaffeModel.deprocess, worker.py:68-71)."""
        x = self.gather([st.x for st in self.strips])
        if x is None:                        # one process per GPU: rank 0 holds the iterate
            return None
        hwc = self.engine.empty(self.H, self.W, 3)
        self.engine.sync_stream()
        self.engine.call('st2_deprocess', C.c_void_p(x.data_ptr()), C.c_void_p(hwc.data_ptr()), self.H, self.W)
        return hwc.cpu().numpy()

    def check(self):
        """Raise if a halo wait ever timed out (a neighbouring strip died)."""
        for st in self._each():
            if st.plan.halo_error():
                raise RuntimeError('strip %d: halo exchange timed out' % st.rank)

    def close(self):
        self._release_strips()
