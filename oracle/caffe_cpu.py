"""Oracle (test infrastructure): CPU restatement of ``worker.CaffeModel``
(``/root/reference/worker.py:32-106``) on top of ``models/vgg19.prototxt:1-337``.
Not a product path -- see ``oracle/__init__.py``.

The arithmetic of the reference's model lives in BVLC Caffe, which is not in
``/root/reference`` and is not pinned to any version (``config.ini:7``).  The
layer semantics restated here are Caffe's published ones [ext]:

* ``Convolution`` (prototxt ``convolution_param {num_output, pad: 1, kernel_size: 3}``):
  cross-correlation, stride 1, zero pad 1, bias added.
* ``ReLU`` in place on the conv's top blob: blob ``convX_Y`` holds post-ReLU data;
  backward multiplies the diff by ``data > 0``.
* ``Pooling`` MAX 2x2 stride 2, pad 0: output extent ``ceil((n - 2) / 2) + 1``, windows clipped
  at the border, arg-max = first maximum in (h, w) scan order (strict ``>`` update).
* ``net.forward`` runs every layer; ``net.backward(start=L, end=E)`` runs layers L..E inclusive,
  *layer* names -- so a diff added to blob ``convX_Y`` enters below ``reluX_Y`` and is not masked,
  while gradient from above passes through ``reluX_Y`` and is (``worker.py:88-106``).
* weight gradients (Caffe computes them, nobody reads them) are not computed.

Arithmetic: torch-CPU ``conv2d`` in fp32 (``dtype=torch.float64`` for error budgeting).
"""
from collections import OrderedDict
import math

import numpy as np
import torch
import torch.nn.functional as F

# (name, kind, out_channels): blob order of models/vgg19.prototxt:3-337
TOPOLOGY = [('data', 'input', 3)]
for _blk, (_n, _c) in enumerate([(2, 64), (2, 128), (4, 256), (4, 512), (4, 512)], start=1):
    for _i in range(1, _n + 1):
        TOPOLOGY.append(('conv%d_%d' % (_blk, _i), 'conv', _c))
    TOPOLOGY.append(('pool%d' % _blk, 'pool', _c))
BLOB_NAMES = [t[0] for t in TOPOLOGY]
CONV_NAMES = [t[0] for t in TOPOLOGY if t[1] == 'conv']


def synthetic_weights(seed=0):
    """Seeded He-normal stand-in for the unavailable ``vgg19.caffemodel``:
    one RandomState, convs visited in prototxt order, ``W ~ N(0, 2/(9 Cin))`` OIHW fp32,
    ``b ~ 0.1 N(0, 1)``."""
    rs = np.random.RandomState(seed)
    params = OrderedDict()
    cin = 3
    for name, kind, cout in TOPOLOGY:
        if kind != 'conv':
            continue
        w = (rs.randn(cout, cin, 3, 3) * math.sqrt(2.0 / (9 * cin))).astype(np.float32)
        b = (rs.randn(cout) * 0.1).astype(np.float32)
        params[name] = (w, b)
        cin = cout
    return params


def pool_out(n):
    """Caffe pooled extent for kernel 2, stride 2, pad 0 (ceil mode)."""
    return int(math.ceil((n - 2) / 2.0)) + 1 if n > 1 else 1


class CaffeCPUModel:
    """The seven-member model seam of ``worker.CaffeModel`` (worker.py:32-106)."""

    mean = np.float32((123.68, 116.779, 103.939)).reshape((3, 1, 1))      # worker.py:34 (RGB)

    def __init__(self, params=None, dtype=torch.float32, full_net=True, threads=None, emulate_fp16=False):
        if threads:
            torch.set_num_threads(threads)
        self.dtype = dtype
        # emulate_fp16: round the operands the way the tensor-core path stores them (weights of
        # conv1_2.. and every activation / gradient tensor to fp16, round-to-nearest-even; fp32
        # accumulation) -- separates "precision" from "semantics" when checking that path.
        self.emulate_fp16 = emulate_fp16
        self.full_net = full_net          # Caffe always runs to pool5 (worker.py:86)
        params = params if params is not None else synthetic_weights(0)
        self.params = OrderedDict(
            (k, (torch.from_numpy(np.ascontiguousarray(w)).to(dtype),
                 torch.from_numpy(np.ascontiguousarray(b)).to(dtype)))
            for k, (w, b) in params.items())
        if emulate_fp16:
            for k in list(self.params)[1:]:
                w, b = self.params[k]
                self.params[k] = (w.half().to(dtype), b)
        self._act = {}
        self._argmax = {}

    # -- worker.py:63-66
    def preprocess(self, image):
        chw = np.float32(image).transpose((2, 0, 1)) - self.mean
        return np.ascontiguousarray(chw[None])

    # -- worker.py:68-71 (no clipping here)
    def deprocess(self, image):
        return (image.squeeze() + self.mean).transpose((1, 2, 0))

    # -- worker.py:73-75
    def layers(self):
        return list(BLOB_NAMES)

    # -- worker.py:77-86
    def forward(self, image, layers=None, top=None):
        """Run the net on a preprocessed (1, 3, H, W) array; return blob name -> fp32 array for
        the requested blobs (all 22 when ``layers`` is None).  ``top`` (oracle-only knob) stops
        early; Caffe itself always runs all layers."""
        want = list(layers) if layers is not None else self.layers()
        cur = torch.from_numpy(np.ascontiguousarray(image)).to(self.dtype)
        self._act = {'data': cur}
        self._argmax = {}
        last = BLOB_NAMES[-1] if (self.full_net and top is None) else (
            top if top is not None else max(want, key=BLOB_NAMES.index))
        with torch.no_grad():
            for name, kind, _ in TOPOLOGY[1:]:
                if kind == 'conv':
                    w, b = self.params[name]
                    cur = self._q(F.relu(F.conv2d(cur, w, b, padding=1)))
                else:
                    cur, idx = F.max_pool2d(cur, 2, 2, ceil_mode=True, return_indices=True)
                    self._argmax[name] = idx
                self._act[name] = cur
                if name == last:
                    break
        out = OrderedDict()
        for name in want:
            out[name] = self._act[name].to(torch.float32).numpy()
        return out

    def _q(self, t):
        return t.half().to(self.dtype) if self.emulate_fp16 else t

    # -- worker.py:88-106
    def backward(self, diffs):
        """Gradient w.r.t. ``data`` of ``sum_l <diffs[l], blob_l>`` with Caffe's segment-wise
        semantics (module docstring).  Needs the activations of the preceding ``forward``."""
        present = [n for n in BLOB_NAMES if n in diffs]
        if not present:
            raise ValueError('no diffs given')
        top = max(present, key=BLOB_NAMES.index)
        g = None                                            # d/d(blob) arriving from above
        with torch.no_grad():
            for idx in range(BLOB_NAMES.index(top), 0, -1):
                name, kind, _ = TOPOLOGY[idx]
                below = BLOB_NAMES[idx - 1]
                inj = diffs.get(name)
                inj = None if inj is None else torch.from_numpy(np.ascontiguousarray(inj)).to(self.dtype)
                if kind == 'conv':
                    if g is not None:
                        g = g * (self._act[name] > 0).to(self.dtype)      # reluX_Y backward
                    if inj is not None:
                        g = inj if g is None else g + inj                 # blob.diff += diffs[l]
                    w, _ = self.params[name]
                    g = F.conv_transpose2d(self._q(g), w, padding=1)      # data gradient only
                    if idx > 1:
                        g = self._q(g)
                else:
                    if inj is not None:
                        g = inj if g is None else g + inj
                    shape = self._act[below].shape
                    g = F.max_unpool2d(self._q(g), self._argmax[name], 2, 2, output_size=shape[-2:])
            if 'data' in diffs:
                d0 = torch.from_numpy(np.ascontiguousarray(diffs['data'])).to(self.dtype)
                g = d0 if g is None else g + d0
        return g.to(torch.float32).numpy()
