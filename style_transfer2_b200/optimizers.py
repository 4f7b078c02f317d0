"""Device-resident optimizers with the step interface of the reference's ``optimizers.py``:
``Optimizer(x, opfunc, step_size=...)``, ``.step() -> (x, loss)``, ``.resample(size, new_x=None)``,
``.objective_changed()``, writable ``.step_size``.  ``x`` is an fp32 CUDA tensor updated in place
(it aliases the caller's parameters, worker.py:175); ``opfunc(x) -> (loss, grad)`` returns a device
gradient that the optimizer owns until the next call.  Arithmetic runs in libst2's fused kernels.
"""
import ctypes as C

import numpy as np
import torch

from . import utils


def _p(t):
    return C.c_void_p(t.data_ptr())


class AdamOptimizer:
    """optimizers.py:7-46.  One fused kernel per step updates both decaying means and x."""

    def __init__(self, x, opfunc, step_size=1, b1=0.9, b2=0.999):
        self.x = x
        self.opfunc = opfunc
        self.step_size = step_size
        self.t = 0
        self.g1 = utils.DecayingMean(b1)
        self.g2 = utils.DecayingMean(b2)

    def _state(self):
        for g in (self.g1, self.g2):
            if not torch.is_tensor(g.mean) or g.mean.shape != self.x.shape:
                g.mean = torch.zeros_like(self.x)
        return self.g1.mean, self.g2.mean

    def step(self):
        """optimizers.py:20-27."""
        self.t += 1
        loss, grad = self.opfunc(self.x)
        m1, m2 = self._state()
        self.g1.items += 1
        self.g2.items += 1
        utils.default_engine().call('st2_adam_step', _p(self.x), _p(grad), _p(m1), _p(m2), self.x.numel(),
                                    float(self.step_size), float(self.g1.decay), float(self.g2.decay),
                                    self.g1.items, self.g2.items)
        return self.x, loss

    def resample(self, size, new_x=None):
        """optimizers.py:29-40: x and the first moment Lanczos, the second moment bilinear then
        max(0, .).  (Upstream raises if the first moment was just cleared; here it restarts at 0.)"""
        if new_x is not None:
            self.x = new_x
            size = tuple(self.x.shape[2:])
        else:
            self.x = utils.resample_nchw(self.x, size)
        if torch.is_tensor(self.g1.mean) and self.g1.items:
            self.g1.mean = utils.resample_nchw(self.g1.mean, size)
        else:
            self.g1.mean = torch.zeros_like(self.x)
        if torch.is_tensor(self.g2.mean):
            # max(0, .) (optimizers.py:39) is applied by the resampling kernel itself (clamp_min_zero)
            self.g2.mean = utils.resample_nchw(self.g2.mean, size, method=utils.BILINEAR, clamp_min_zero=True)
        return self.x

    def objective_changed(self):
        """optimizers.py:42-46: forget the first moment only."""
        self.t = 0
        self.g1.clear()


class LBFGSOptimizer:
    """optimizers.py:49-125.  History, s.y, y.y, alphas and every dot product of the two-loop
    recursion live on the device; a step issues a fixed launch sequence and never synchronises."""

    def __init__(self, x, opfunc, step_size=1, n_corr=10):
        self.x = x
        self.opfunc = opfunc
        self.step_size = step_size
        self.n_corr = n_corr
        self.loss = None
        self.grad = None
        # optional hook, called with x right after it has been advanced (and before the objective is evaluated there):
        # lets a caller start moving the new iterate -- e.g. to the host -- while the evaluation runs
        self.after_advance = None
        self._eng = utils.default_engine()
        self._h = None
        self._n = 0
        self._ensure()

    def _ensure(self):
        if self.x is None:                      # the reference builds optimizers before any image is set
            return
        if self._h is not None and self._n == self.x.numel():
            return
        self._free()
        h = C.c_void_p()
        self._eng.call('st2_lbfgs_create', self.x.numel(), self.n_corr, C.byref(h))
        self._h, self._n = h, self.x.numel()

    def _free(self):
        if self._h is not None:
            self._eng.lib.st2_lbfgs_destroy(self._h)
            self._h = None

    def _call(self, name, *args):
        from . import _lib
        _lib.check(self._eng.ctx, getattr(self._eng.lib, name)(self._h, *args), name)

    def step(self):
        """optimizers.py:62-77."""
        self._ensure()
        if self.loss is None:
            self.loss, self.grad = self.opfunc(self.x)
        self._call('st2_lbfgs_advance', _p(self.x), _p(self.grad), float(self.step_size))
        if self.after_advance is not None:
            self.after_advance(self.x)          # the new iterate is final here: the evaluation below only reads it
        loss, grad = self.opfunc(self.x)
        if grad.data_ptr() == self.grad.data_ptr():
            raise RuntimeError('opfunc must hand out a fresh gradient buffer (the optimizer owns the previous one)')
        self._call('st2_lbfgs_commit', _p(grad), _p(self.grad))
        self.loss, self.grad = loss, grad
        return self.x, loss

    def resample(self, size, new_x=None):
        """optimizers.py:110-119."""
        self.x = new_x if new_x is not None else utils.resample_nchw(self.x, size)
        self._ensure()
        self.objective_changed()
        return self.x

    def objective_changed(self):
        """optimizers.py:121-125."""
        if self._h is not None:
            self._call('st2_lbfgs_reset')
        self.loss, self.grad = None, None

    # -- state access (teacher-forced parity tests, checkpointing)
    def export_state(self):
        if self._h is None:                     # no x yet (the reference builds optimizers before any image is set)
            return (torch.empty((0,)), torch.empty((0,)), [])
        cnt = C.c_int()
        S = torch.empty((self.n_corr, self._n), dtype=torch.float32, device=self.x.device)
        Y = torch.empty_like(S)
        sy = (C.c_double * self.n_corr)()
        self._call('st2_lbfgs_export', C.byref(cnt), _p(S), _p(Y), sy)
        m = cnt.value
        shape = (m,) + tuple(self.x.shape)
        return S[:m].reshape(shape), Y[:m].reshape(shape), [sy[i] for i in range(m)]

    def load_state(self, S, Y, sy, grad, loss):
        """S, Y: (m, ...) arrays, oldest pair first; grad: gradient at the current x."""
        self._ensure()
        dev = self.x.device
        m = len(sy)
        arr = (C.c_double * max(m, 1))(*[float(v) for v in sy])
        if m == 0:                              # e.g. a checkpoint taken right after objective_changed()
            self._call('st2_lbfgs_load', 0, C.c_void_p(0), C.c_void_p(0), arr)
        else:
            S = torch.as_tensor(np.ascontiguousarray(S, np.float32)).reshape(m, -1).to(dev).contiguous()
            Y = torch.as_tensor(np.ascontiguousarray(Y, np.float32)).reshape(m, -1).to(dev).contiguous()
            self._call('st2_lbfgs_load', m, _p(S), _p(Y), arr)
        torch.cuda.synchronize(dev)
        self.grad = torch.as_tensor(np.ascontiguousarray(grad, np.float32)).to(dev).reshape(self.x.shape).contiguous()
        self.loss = loss

    @property
    def sk(self):
        return list(self.export_state()[0])

    @property
    def yk(self):
        return list(self.export_state()[1])

    @property
    def syk(self):
        return self.export_state()[2]

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass
