"""``B200Model``: the model seam of the reference's ``worker.CaffeModel`` (worker.py:32-106) backed by
libst2's sm_100a kernels.  Same seven members -- ``mean``, ``preprocess``, ``deprocess``, ``layers``,
``forward``, ``backward``, ``reload_net`` -- with NumPy arrays in and out, plus a device-resident
path (``plan`` / ``Plan``) used by this package's own ``StyleTransfer`` so that nothing but the
iterate and ~30 trace scalars ever leaves the GPU.

PyTorch is used for device memory, streams and host<->device copies only.
"""
from collections import OrderedDict
import ctypes as C
import logging
import os

import numpy as np
import torch

from . import _lib, vgg

logger = logging.getLogger('worker')

PRECISIONS = {'fp32': _lib.PREC_FP32, 'fp16': _lib.PREC_FP16}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    """One libst2 context (device + packed weights)."""

    def __init__(self, device=0, params=None):
        if not torch.cuda.is_available():
            raise RuntimeError('style_transfer2_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        self.lib = _lib.lib()
        self.device_index = max(int(device), 0)           # config.ini gpu = -1 meant "CPU" upstream
        self.device = torch.device('cuda', self.device_index)
        torch.cuda.set_device(self.device)
        handle = C.c_void_p()
        _lib.check(None, self.lib.st2_ctx_create(self.device_index, C.byref(handle)), 'st2_ctx_create')
        self.ctx = handle
        self.sync_stream()
        self.load_weights(params if params is not None else vgg.synthetic_weights(0))

    def sync_stream(self):
        """Point libst2 at torch's current stream so CUDA events / graphs see its launches."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.ctx, self.lib.st2_set_stream(self.ctx, C.c_void_p(stream)), 'st2_set_stream')

    def load_weights(self, params):
        for idx, name in enumerate(vgg.CONVS):
            if name not in params:               # a network cut below this layer (vgg.net_from_prototxt): never evaluated
                continue
            w, b = params[name]
            w = np.ascontiguousarray(w, np.float32)
            b = np.ascontiguousarray(b, np.float32)
            rc = self.lib.st2_set_conv_weights(self.ctx, idx, w.ctypes.data_as(C.c_void_p),
                                               b.ctypes.data_as(C.c_void_p), w.shape[0], w.shape[1])
            _lib.check(self.ctx, rc, 'st2_set_conv_weights(%s)' % name)

    def trace_slot(self, owner):
        """A row of pinned host memory for one evaluation's scalar block.  Rows come from a ring
        allocated once (cudaHostAlloc costs ~1 ms, far too much per iteration); when the ring wraps, the
        previous owner of a row is told to move its data out first (``owner.detach()``)."""
        if getattr(self, '_ring', None) is None:
            self._ring = torch.empty((128, _lib.SCAL_TOTAL), dtype=torch.float64, pin_memory=True)
            self._ring_owner = [None] * self._ring.shape[0]
            self._ring_next = 0
        i = self._ring_next
        self._ring_next = (i + 1) % self._ring.shape[0]
        prev = self._ring_owner[i]
        if prev is not None:
            prev = prev()
            if prev is not None:
                prev.detach()
        import weakref
        self._ring_owner[i] = weakref.ref(owner)
        return self._ring[i]

    def launches(self):
        """Kernels launched through this context, incl. those replayed from CUDA graphs (worker._graph_step)."""
        return int(self.lib.st2_launch_count(self.ctx)) + int(getattr(self, 'graph_launches', 0))

    def profile(self, enable):
        """Per-category CUDA-event timing (st2_profile); the graph path is bypassed while it is on."""
        self.profiling = bool(enable)
        self.call('st2_profile', 1 if enable else 0)

    def call(self, name, *args):
        _lib.check(self.ctx, getattr(self.lib, name)(self.ctx, *args), name)

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=torch.float32):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def __del__(self):
        try:
            self.lib.st2_ctx_destroy(self.ctx)
        except Exception:
            pass


class Plan:
    """Activations, gradients, loss targets and the scalar block for one canvas size."""

    def __init__(self, engine, height, width, precision):
        self.engine, self.H, self.W = engine, int(height), int(width)
        self.lib = engine.lib
        handle = C.c_void_p()
        _lib.check(engine.ctx, self.lib.st2_plan_create(engine.ctx, self.H, self.W, precision, C.byref(handle)),
                   'st2_plan_create')
        self.handle = handle
        self._scal_host = (C.c_double * _lib.SCAL_TOTAL)()

    def _check(self, rc, what):
        _lib.check(self.engine.ctx, rc, what)

    def blob_dims(self, blob):
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.st2_plan_blob_dims(self.handle, blob, C.byref(c), C.byref(h), C.byref(w)), 'blob_dims')
        return c.value, h.value, w.value

    def forward(self, x, top):
        self._check(self.lib.st2_forward(self.handle, _ptr(x), top), 'st2_forward')

    def export(self, blob):
        c, h, w = self.blob_dims(blob)
        out = self.engine.empty(1, c, h, w)
        self._check(self.lib.st2_blob_export(self.handle, blob, _ptr(out)), 'st2_blob_export')
        return out

    def backward(self, diffs, grad_out):
        """diffs: {blob index: device fp32 NCHW tensor}."""
        n = len(diffs)
        blobs = (C.c_int * n)(*diffs.keys())
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in diffs.values()])
        self._check(self.lib.st2_backward(self.handle, n, blobs, ptrs, _ptr(grad_out)), 'st2_backward')

    def capture_content(self, blob):
        self._check(self.lib.st2_capture_content(self.handle, blob), 'st2_capture_content')

    def gram(self, blob):
        c = vgg.TOPOLOGY[blob][2]
        out = self.engine.empty(c, c)
        self._check(self.lib.st2_gram(self.handle, blob, _ptr(out)), 'st2_gram')
        return out

    def set_style_gram(self, blob, gram):
        self._check(self.lib.st2_set_style_gram(self.handle, blob, _ptr(gram)), 'st2_set_style_gram')

    def set_blob_weights(self, blob, c, s, d):
        self._check(self.lib.st2_set_blob_weights(self.handle, blob, c, s, d), 'st2_set_blob_weights')

    def set_eval_order(self, blobs):
        arr = (C.c_int * max(len(blobs), 1))(*blobs)
        self._check(self.lib.st2_set_eval_order(self.handle, len(blobs), arr), 'st2_set_eval_order')

    def set_params(self, tv, tv_power, p, p_power):
        self._check(self.lib.st2_set_params(self.handle, tv, tv_power, p, p_power), 'st2_set_params')

    def reset_norms(self):
        self._check(self.lib.st2_reset_norms(self.handle), 'st2_reset_norms')

    def set_norm(self, kind, blob, value):
        self._check(self.lib.st2_set_norm(self.handle, 'csd'.index(kind), blob, float(value)), 'st2_set_norm')

    def eval(self, x, grad, want_grad=True):
        self._check(self.lib.st2_eval(self.handle, _ptr(x), _ptr(grad), 1 if want_grad else 0), 'st2_eval')

    def read_scalars(self):
        """One device->host copy of the scalar block; synchronises the stream."""
        self._check(self.lib.st2_read_scalars(self.handle, self._scal_host), 'st2_read_scalars')
        return np.frombuffer(self._scal_host, dtype=np.float64).copy()

    def copy_scalars_async(self, pinned):
        """Enqueue a copy of the scalar block into a pinned float64 host tensor (no sync)."""
        self._check(self.lib.st2_copy_scalars_async(self.handle, C.c_void_p(pinned.data_ptr())),
                    'st2_copy_scalars_async')

    def close(self):
        if self.handle:
            self.lib.st2_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class B200Model:
    """Drop-in for ``worker.CaffeModel``."""

    mean = np.float32((123.68, 116.779, 103.939)).reshape((3, 1, 1))      # worker.py:34, RGB

    def __init__(self, prototxt=None, caffemodel=None, gpu=-1, precision=None, params=None):
        self.prototxt = None if prototxt is None else str(prototxt)
        self.caffemodel = None if caffemodel is None else str(caffemodel)
        precision = precision or os.environ.get('ST2_PRECISION', 'fp16')
        if precision not in PRECISIONS:
            raise ValueError('precision must be one of %s' % sorted(PRECISIONS))
        self.precision_name = precision
        self.precision = PRECISIONS[precision]
        self.gpu = gpu
        self._params = params
        self._plans = OrderedDict()
        self._free_plans = {}            # (H, W) -> idle plans handed back by finished jobs (serving)
        self._last_plan = None
        logger.info('Initializing the B200 engine (%s).', precision)
        self.reload_net()

    # -- worker.py:58-61
    def reload_net(self):
        self._n_blobs = len(vgg.BLOBS)
        if self.prototxt and os.path.exists(self.prototxt):
            # the blobs the prototxt describes: the reference's truncated VGG-19, a shorter cut of it, or a deploy
            # file with the classifier tail still attached (ignored: the path never evaluates it)
            with open(self.prototxt) as f:
                blobs, ignored = vgg.net_from_prototxt(f.read(), strict=False)
            self._n_blobs = len(blobs)
            if ignored:
                logger.warning('Ignoring %d layer(s) above %s: %s', len(ignored), blobs[-1][0], ', '.join(ignored))
        params = self._params
        if params is None and self.caffemodel and os.path.exists(self.caffemodel):
            params = vgg.read_caffemodel(self.caffemodel)
        if params is None:
            logger.warning('No caffemodel at %s: using seeded synthetic weights.', self.caffemodel)
            params = vgg.synthetic_weights(0)
        for plan in self._plans.values():
            plan.close()
        self._plans.clear()
        for plans in self._free_plans.values():
            for plan in plans:
                plan.close()
        self._free_plans.clear()
        self.engine = Engine(self.gpu, params)

    # -- worker.py:63-71 (host versions, identical arithmetic)
    def preprocess(self, image):
        arr = np.float32(image).transpose((2, 0, 1)) - self.mean
        return np.ascontiguousarray(arr[None])

    def deprocess(self, image):
        return (image.squeeze() + self.mean).transpose((1, 2, 0))

    # -- worker.py:73-75
    def layers(self):
        return list(vgg.BLOBS[:getattr(self, '_n_blobs', len(vgg.BLOBS))])

    def plan(self, height, width, keep=3):
        """Plan for a canvas size (reshape-on-demand, worker.py:84); a few sizes stay cached."""
        key = (int(height), int(width))
        plan = self._plans.pop(key, None)
        if plan is None:
            plan = Plan(self.engine, key[0], key[1], self.precision)
            while len(self._plans) >= keep:
                _, old = self._plans.popitem(last=False)
                old.close()
        self._plans[key] = plan
        return plan

    def acquire_plan(self, height, width):
        """A plan owned by the caller until ``release_plan``.  Creating / destroying a plan costs ~50
        synchronising cudaMalloc / cudaFree calls, so a serving process recycles them: a recycled plan keeps
        its buffers, the next owner resets norms, weights and targets."""
        free = self._free_plans.get((int(height), int(width)))
        if free:
            return free.pop()
        return Plan(self.engine, int(height), int(width), self.precision)

    def release_plan(self, plan, keep=12):
        if plan is None or not plan.handle:
            return
        bucket = self._free_plans.setdefault((plan.H, plan.W), [])
        if sum(len(v) for v in self._free_plans.values()) >= keep:
            torch.cuda.current_stream(self.engine.device).synchronize()
            plan.close()
        else:
            bucket.append(plan)

    # -- worker.py:77-86
    def forward(self, image, layers=None):
        """NumPy seam: run the net on a preprocessed (1, 3, H, W) array and return blob name ->
        fp32 NCHW array for the requested blobs.  (Caffe returns views that the next forward
        overwrites; these are fresh copies.)  Runs to the highest requested blob only -- layers
        above it cannot influence the result."""
        names = self.layers() if layers is None else list(layers)
        image = np.ascontiguousarray(image, np.float32)
        if image.ndim != 4 or image.shape[0] != 1 or image.shape[1] != 3:
            raise ValueError('expected a (1, 3, H, W) array, got %s' % (image.shape,))
        self.engine.sync_stream()
        plan = self.plan(image.shape[2], image.shape[3])
        x = torch.from_numpy(image).to(self.engine.device)
        top = max((vgg.BLOB_INDEX[n] for n in names), default=0)
        if top >= self._n_blobs:
            raise KeyError('%s is not a blob of this network (%s is its last)' % (vgg.BLOBS[top], self.layers()[-1]))
        plan.forward(x, top)
        self._last_plan, self._last_x = plan, x
        out = OrderedDict()
        for n in names:
            out[n] = plan.export(vgg.BLOB_INDEX[n]).cpu().numpy()
        return out

    # -- worker.py:88-106
    def backward(self, diffs):
        if self._last_plan is None:
            raise RuntimeError('backward() before forward()')
        plan = self._last_plan
        dev = {}
        for name, arr in diffs.items():
            b = vgg.BLOB_INDEX[name]
            c, h, w = plan.blob_dims(b)
            a = np.ascontiguousarray(arr, np.float32).reshape(1, c, h, w)
            dev[b] = torch.from_numpy(a).to(self.engine.device)
        grad = self.engine.empty(1, 3, plan.H, plan.W)
        plan.backward(dev, grad)
        return grad.cpu().numpy()
