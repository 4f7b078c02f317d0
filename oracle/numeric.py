"""Oracle (test infrastructure): NumPy restatement of the numeric helpers in
``/root/reference/utils.py``.  Not a product path -- see ``oracle/__init__.py``.

Every function names the reference lines it follows.  All arithmetic is fp32
unless the reference itself widens (Pillow's resampler accumulates in fp64).
"""
from collections import OrderedDict
import math

import numpy as np


# --------------------------------------------------------------------------- level-1 ops
def sdot(a, b):
    """``utils.dot`` (utils.py:29-35): single-precision dot product of two equal-shape arrays,
    returned as a Python float.  BLAS ``sdot`` accumulates in fp32; we accumulate pairwise in fp32
    (``np.dot`` on fp32 vectors does the same call)."""
    if a.shape != b.shape:
        raise ValueError('Sizes do not match: x=%s y=%s' % (a.shape, b.shape))
    return float(np.dot(a.reshape(-1), b.reshape(-1)))


def saxpy(alpha, x, y):
    """``utils.axpy`` (utils.py:38-46): ``y <- alpha*x + y`` in place, fp32; returns ``y``."""
    if x.shape != y.shape:
        raise ValueError('Sizes do not match: x=%s y=%s' % (x.shape, y.shape))
    y += np.float32(alpha) * x
    return y


class EMA:
    """``utils.DecayingMean`` (utils.py:49-69): bias-corrected exponentially decaying mean.
    ``update(v)`` folds a sample in; ``value()`` is ``mean / (1 - decay**items)`` (or the raw mean,
    integer 0, while empty)."""

    def __init__(self, decay=0.9):
        self.decay = decay
        self.clear()

    def clear(self):
        self.mean = 0
        self.items = 0

    def update(self, sample):
        self.mean = self.decay * self.mean + (1 - self.decay) * sample
        self.items += 1
        return self.value()

    def value(self):
        if not self.items:
            return self.mean
        return self.mean / (1 - self.decay ** self.items)


# --------------------------------------------------------------------------- pixel-space terms
def total_variation(x, beta=2):
    """``utils.tv_norm`` (utils.py:285-297) with ``roll_by_one`` (utils.py:232-254) folded in.

    ``x`` is (1, C, H, W).  Forward differences wrap around (circular) on both axes; the gradient
    is the adjoint of that circular difference.  Returns ``(norm, grad)``.
    """
    nxt_w = np.roll(x, -1, axis=3)                      # x[..., j+1], wrap
    nxt_h = np.roll(x, -1, axis=2)
    dw = x - nxt_w
    dh = x - nxt_h
    mag2 = dw ** 2 + dh ** 2 + 1e-8
    norm = np.sum(mag2 ** (beta / 2))
    k = (beta / 2) * mag2 ** (beta / 2 - 1)
    gw = 2 * dw * k
    gh = 2 * dh * k
    grad = gw + gh
    grad -= np.roll(gw, 1, axis=3)
    grad -= np.roll(gh, 1, axis=2)
    return norm, grad


def p_norm(x, p=2):
    """``utils.p_norm`` (utils.py:300-304): ``sum |x|^p / p`` and ``sign(x)|x|^(p-1)``."""
    mag = abs(x)
    return np.sum(mag ** p) / p, np.sign(x) * mag ** (p - 1)


def rms(a):
    """``Trace.rms`` body (utils.py:280-282): ``sqrt(mean(a**2))``."""
    return np.sqrt(np.mean(a ** 2))


class TraceLog:
    """``utils.Trace`` (utils.py:257-282): ordered name -> scalar record of one objective
    evaluation.  Duplicate names get ``_`` appended; NumPy scalars become Python numbers."""

    def __init__(self):
        self.data = OrderedDict()

    def put(self, key, value):
        while key in self.data:
            key += '_'
        if isinstance(value, np.floating):
            self.data[key] = float(value)
        elif isinstance(value, np.integer):
            self.data[key] = int(value)
        else:
            self.data[key] = value
        return value

    def put_rms(self, key, arr):
        self.put(key, rms(arr))
        return arr


# --------------------------------------------------------------------------- resampling
def _lanczos3(t):
    t = np.abs(t)
    out = np.zeros_like(t)
    inside = t < 3.0
    tz = t[inside]
    with np.errstate(divide='ignore', invalid='ignore'):
        v = np.where(tz == 0.0, 1.0,
                     (np.sin(np.pi * tz) / (np.pi * tz)) * (np.sin(np.pi * tz / 3.0) / (np.pi * tz / 3.0)))
    out[inside] = v
    return out


def _triangle(t):
    return np.maximum(0.0, 1.0 - np.abs(t))


_FILTERS = {'lanczos': (_lanczos3, 3.0), 'bilinear': (_triangle, 1.0)}


def resample_coeffs(n_in, n_out, method='lanczos'):
    """Per-output-index window ``(xmin, weights)`` of Pillow's ``precompute_coeffs``
    (libImaging/Resample.c [ext]; what ``Image.resize`` on a mode-'F' image uses, reached from
    utils.py:130-131).  fp64 weights, window truncated at the borders and renormalised."""
    filt, support0 = _FILTERS[method]
    scale = n_in / n_out
    fscale = max(scale, 1.0)
    support = support0 * fscale
    out = []
    for xx in range(n_out):
        centre = (xx + 0.5) * scale
        xmin = max(int(centre - support + 0.5), 0)
        xmax = min(int(centre + support + 0.5), n_in)
        xs = np.arange(xmin, xmax, dtype=np.float64)
        w = filt((xs - centre + 0.5) / fscale)
        tot = w.sum()
        if tot != 0.0:
            w = w / tot
        out.append((xmin, w))
    return out


def resample_plane(plane, hw, method='lanczos'):
    """Separable resize of one fp32 (H, W) plane to ``hw``: horizontal pass into an fp32
    temporary, then vertical, each output an fp64-accumulated dot product (Resample.c [ext]
    ``ImagingResampleHorizontal_32bpc`` / ``Vertical_32bpc``).  Pillow skips a pass whose size is
    unchanged."""
    src = np.asarray(plane, np.float32)
    h_in, w_in = src.shape
    h_out, w_out = hw
    if w_out != w_in:
        tmp = np.empty((h_in, w_out), np.float32)
        for xx, (x0, w) in enumerate(resample_coeffs(w_in, w_out, method)):
            tmp[:, xx] = (src[:, x0:x0 + len(w)].astype(np.float64) @ w).astype(np.float32)
        src = tmp
    if h_out != h_in:
        tmp = np.empty((h_out, src.shape[1]), np.float32)
        for yy, (y0, w) in enumerate(resample_coeffs(h_in, h_out, method)):
            tmp[yy, :] = (w @ src[y0:y0 + len(w), :].astype(np.float64)).astype(np.float32)
        src = tmp
    return src


def resample_nchw(a, hw, method='lanczos'):
    """``utils.resample_nchw`` (utils.py:148-160): resize every (n, c) plane of an NCHW array
    independently in floating point.  The reference fans planes out to a thread pool and lets
    Pillow do the arithmetic; this restatement runs the same per-plane algorithm serially."""
    a = np.float32(a)
    n, c = a.shape[:2]
    out = np.zeros((n, c, hw[0], hw[1]), np.float32)
    for i in range(n):
        for j in range(c):
            out[i, j] = resample_plane(a[i, j], hw, method)
    return out


def fit_into_square(current_wh, size, scale_up=False):
    """``utils.fit_into_square`` (utils.py:210-223): aspect-preserving (w, h) inside size x size."""
    size = int(round(size))
    w, h = current_wh
    if not scale_up and max(w, h) <= size:
        return current_wh
    if w > h:
        return (size, int(round(size * h / w)))
    return (int(round(size * w / h)), size)


def scale_pyramid(size, min_size=1, factor=math.sqrt(2)):
    """``utils.scales`` (utils.py:193-207): list of sizes growing by ``factor`` up to ``size``."""
    cur = np.float64(size)
    min_size = int(min_size)
    assert min_size >= 1
    sizes = [tuple(int(round(v)) for v in cur)]
    while True:
        cur = cur / factor
        as_int = tuple(int(round(v)) for v in cur)
        if max(as_int) < min_size or min(as_int) < 1:
            break
        sizes.append(as_int)
    return sizes[::-1]
