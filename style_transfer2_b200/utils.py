"""Host-side helpers of the drop-in worker: the numeric entry points of the reference's ``utils.py``
re-targeted at device tensors (``dot``, ``axpy``, ``DecayingMean``, ``tv_norm``, ``p_norm``,
``resample_nchw``, ``Trace``) plus the configuration / logging plumbing the worker needs
(``parse_args``, ``read_config``, ``setup_logging``, ``setup_signals``).  Reference: utils.py:29-309.
"""
import argparse
import configparser
import ctypes as C
from collections import OrderedDict
import logging
from pathlib import Path
import signal
import warnings

import numpy as np
import torch

from . import _lib

MODULE_DIR = Path(__file__).parent.resolve()
REPO_DIR = MODULE_DIR.parent
LANCZOS, BILINEAR = _lib.RESAMPLE_LANCZOS, _lib.RESAMPLE_BILINEAR

_ENGINE = None


def set_default_engine(engine):
    global _ENGINE
    _ENGINE = engine


def default_engine():
    if _ENGINE is None:
        raise RuntimeError('no B200 engine yet: create a B200Model first')
    return _ENGINE


def _p(t):
    return C.c_void_p(t.data_ptr())


def _same_shape(x, y):
    if x.shape != y.shape:
        raise ValueError('Sizes do not match: x=%s y=%s' % (tuple(x.shape), tuple(y.shape)))


def dot(x, y):
    """utils.py:29-35 -- dot product of two equal-shape fp32 device tensors as a Python float."""
    _same_shape(x, y)
    out = C.c_double()
    default_engine().call('st2_dot', _p(x), _p(y), x.numel(), C.byref(out))
    return float(np.float32(out.value))


def axpy(a, x, y):
    """utils.py:38-46 -- y <- a*x + y in place; returns y."""
    _same_shape(x, y)
    default_engine().call('st2_axpy', float(a), _p(x), _p(y), x.numel())
    return y


def rms(x):
    out = C.c_double()
    default_engine().call('st2_sumsq', _p(x), x.numel(), C.byref(out))
    return float(np.sqrt(out.value / x.numel()))


class DecayingMean:
    """utils.py:49-69 -- bias-corrected exponentially decaying mean over device tensors.  The update
    itself is fused into ``st2_adam_step``; this object carries the state (``mean``, ``items``)."""

    def __init__(self, decay=0.9):
        self.decay = decay
        self.mean = 0
        self.items = 0

    def clear(self):
        if torch.is_tensor(self.mean):
            self.mean.zero_()
        self.items = 0

    def correction(self):
        return 1 - self.decay ** self.items if self.items else 1.0

    def __call__(self):
        if self.items == 0:
            return self.mean
        return self.mean / (1 - self.decay ** self.items)


def resample_nchw(a, hw, method=LANCZOS, clamp_min_zero=False):
    """utils.py:148-160 -- per-plane floating point resize of an NCHW device tensor with Pillow's
    antialiased Lanczos-3 / bilinear arithmetic (fp64 weights, horizontal then vertical pass)."""
    eng = default_engine()
    a = a.contiguous()
    n, c, h, w = a.shape
    out = torch.empty((n, c, int(hw[0]), int(hw[1])), dtype=torch.float32, device=a.device)
    eng.call('st2_resample', _p(a), n * c, h, w, _p(out), int(hw[0]), int(hw[1]), method,
             1 if clamp_min_zero else 0)
    return out


def _pixel_terms(x, tv, tv_power, p, p_power):
    eng = default_engine()
    x = x.contiguous()
    scal = torch.zeros(32, dtype=torch.float64, device=x.device)
    grad = torch.empty_like(x)
    zero = torch.zeros_like(x)
    eng.call('st2_pixel_terms', _p(x), _p(zero), _p(grad), x.shape[1], x.shape[2], x.shape[3], tv, tv_power, p,
             p_power, 1.0, _p(scal))
    return scal.cpu().numpy(), grad


def tv_norm(x, beta=2):
    """utils.py:285-297 -- total-variation norm (circular differences) and its gradient."""
    s, g = _pixel_terms(x, 1.0, float(beta), 0.0, 2.0)
    return float(s[_lib.G_TV_NORM]), g


def p_norm(x, p=2):
    """utils.py:300-304 -- (sum|x|^p / p, sign(x)|x|^(p-1))."""
    s, g = _pixel_terms(x, 0.0, 2.0, 1.0, float(p))
    return float(s[_lib.G_P_NORM]) / p, g


class Trace:
    """utils.py:257-282 -- ordered snapshot of named scalars of one objective evaluation."""

    def __init__(self):
        self.data = OrderedDict()

    def __call__(self, name, expr):
        while name in self.data:
            name += '_'
        if isinstance(expr, np.floating):
            expr = float(expr)
        elif isinstance(expr, np.integer):
            expr = int(expr)
        elif isinstance(expr, np.generic):
            warnings.warn('Did not convert NumPy scalar to Python scalar, may not be pickleable', RuntimeWarning)
        self.data[name] = expr
        return expr

    def __str__(self):
        return ', '.join('%s: %g' % item for item in self.data.items())

    def rms(self, name, expr):
        self(name, rms(expr) if torch.is_tensor(expr) else float(np.sqrt(np.mean(np.square(expr)))))
        return expr


# ------------------------------------------------------------------------------- plumbing
def parse_args(description=None):
    """utils.py:110-117."""
    parser = argparse.ArgumentParser(description=description)
    parser.add_argument('config', nargs='?', help='a configuration file to load after config.ini')
    parser.add_argument('--debug', '-d', action='count', default=0, help='enable debug logging')
    return parser.parse_args()


def read_config(args=None, search_dirs=None):
    """utils.py:120-127 -- ``[DEFAULT]`` of config.ini, then config_non_git.ini, then the CLI file.
    Looks next to this package's parent (the repo root, where a checkout of the reference's files
    would sit) unless ``search_dirs`` says otherwise.  Missing files are skipped, as upstream."""
    cp = configparser.ConfigParser()
    files = []
    for d in (search_dirs or [REPO_DIR]):
        files += [str(Path(d) / 'config.ini'), str(Path(d) / 'config_non_git.ini')]
    if args is not None and getattr(args, 'config', None):
        files.append(args.config)
    cp.read(files)
    return cp['DEFAULT']


def setup_logging(debug=0):
    """utils.py:172-184."""
    fmt = '%(asctime)s.%(msecs)03d %(process)d %(name)s %(levelname)s: %(message)s'
    logging.basicConfig(level=logging.DEBUG if debug else logging.INFO, format=fmt, datefmt='%H:%M:%S')
    logging.captureWarnings(True)


def setup_signals():
    """utils.py:187-190 -- SIGHUP ends the worker like Ctrl-C."""
    def handler(*_):
        raise KeyboardInterrupt()
    signal.signal(signal.SIGHUP, handler)


def scales(size, min_size=1, factor=np.sqrt(2)):
    """utils.py:193-207."""
    cur = np.float64(size)
    min_size = int(min_size)
    assert min_size >= 1
    sizes = [tuple(int(round(v)) for v in cur)]
    while True:
        cur = cur / factor
        nxt = tuple(int(round(v)) for v in cur)
        if max(nxt) < min_size or min(nxt) < 1:
            break
        sizes.append(nxt)
    return sizes[::-1]


def fit_into_square(current_size, size, scale_up=False):
    """utils.py:210-223."""
    size = int(round(size))
    w, h = current_size
    if not scale_up and max(w, h) <= size:
        return current_size
    if w > h:
        return (size, int(round(size * h / w)))
    return (int(round(size * w / h)), size)
