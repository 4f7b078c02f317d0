"""Oracle (test infrastructure): NumPy restatement of the reference's objective and optimizers
(``/root/reference/worker.py:109-315`` and ``/root/reference/optimizers.py:7-125``).
Not a product path -- see ``oracle/__init__.py``.  Pinned against the reference's own Python by
``oracle/make_golden.py`` -> ``tests/golden/`` -> ``tests/test_oracle_golden.py``.
"""
import time

import numpy as np
import pandas as pd

from . import numeric as nm

LOSS_NAMES = ('content', 'style', 'deepdream')              # messages.py:143
SCALAR_LOSS_NAMES = ('tv', 'tv_power', 'p', 'p_power')      # messages.py:144
EPS_W = 1e-15


def gram(feat):
    """``worker.gram_matrix`` (worker.py:109-114): ``X X^T / (C*H*W)`` for a (1, C, H, W) blob."""
    n, c, h, w = feat.shape
    assert n == 1
    flat = feat.reshape((c, h * w))
    return np.dot(flat, flat.T) / np.float32(flat.size)


# =========================================================================== optimizers
class LBFGS:
    """``optimizers.LBFGSOptimizer`` (optimizers.py:49-125): memory-10 L-BFGS, fixed step length,
    no line search.  ``x`` is updated in place (it aliases the caller's array)."""

    def __init__(self, x, opfunc, step_size=1, n_corr=10):
        self.x, self.opfunc, self.step_size, self.n_corr = x, opfunc, step_size, n_corr
        self.loss = self.grad = None
        self.S, self.Y, self.SY = [], [], []

    def step(self):                                             # optimizers.py:62-77
        if self.loss is None:
            self.loss, self.grad = self.opfunc(self.x)
        s = -self.step_size * self.direction(self.grad)
        self.x += s
        loss, grad = self.opfunc(self.x)
        self.remember(s, grad - self.grad)
        self.loss, self.grad = loss, grad
        return self.x, loss

    def remember(self, s, y):                                   # optimizers.py:79-87
        sy = nm.sdot(s, y)
        if sy > 1e-10:
            self.S.append(s), self.Y.append(y), self.SY.append(sy)
        if len(self.S) > self.n_corr:
            del self.S[0], self.Y[0], self.SY[0]

    def direction(self, g):                                     # optimizers.py:89-108 (inv_hv)
        q = g.copy()
        m = len(self.S)
        alpha = [0.0] * m
        for i in range(m - 1, -1, -1):
            alpha[i] = nm.sdot(self.S[i], q) / self.SY[i]
            nm.saxpy(-alpha[i], self.Y[i], q)
        if m:
            q *= self.SY[-1] / nm.sdot(self.Y[-1], self.Y[-1])
        else:
            q /= np.sqrt(nm.sdot(q, q) / q.size)
        for i in range(m):
            beta = nm.sdot(self.Y[i], q) / self.SY[i]
            nm.saxpy(alpha[i] - beta, self.S[i], q)
        return q

    def resample(self, size, new_x=None):                       # optimizers.py:110-119
        self.x = new_x if new_x is not None else nm.resample_nchw(self.x, size)
        self.objective_changed()
        return self.x

    def objective_changed(self):                                # optimizers.py:121-125
        self.S, self.Y, self.SY = [], [], []
        self.loss = self.grad = None


class Adam:
    """``optimizers.AdamOptimizer`` (optimizers.py:7-46) over ``utils.DecayingMean``."""

    def __init__(self, x, opfunc, step_size=1, b1=0.9, b2=0.999):
        self.x, self.opfunc, self.step_size = x, opfunc, step_size
        self.t = 0
        self.m1, self.m2 = nm.EMA(b1), nm.EMA(b2)

    def step(self):                                             # optimizers.py:20-27
        self.t += 1
        loss, g = self.opfunc(self.x)
        self.m1.update(g)
        self.m2.update(g ** 2)
        self.x -= self.step_size * self.m1.value() / (np.sqrt(self.m2.value()) + 1e-8)
        return self.x, loss

    def resample(self, size, new_x=None):                       # optimizers.py:29-40
        if new_x is not None:
            self.x = new_x
            size = self.x.shape[2:]
        else:
            self.x = nm.resample_nchw(self.x, size)
        self.m1.mean = nm.resample_nchw(self.m1.mean, size)
        self.m2.mean = np.maximum(0, nm.resample_nchw(self.m2.mean, size, method='bilinear'))
        return self.x

    def objective_changed(self):                                # optimizers.py:42-46
        self.t = 0
        self.m1.clear()


OPTIMIZERS = {'adam': Adam, 'lbfgs': LBFGS}                     # messages.py:117-118
DEFAULT_STEP = {'adam': 10, 'lbfgs': 1}                         # messages.py:119


# =========================================================================== objective
class Transfer:
    """``worker.StyleTransfer`` (worker.py:117-315) restated."""

    def __init__(self, model):
        self.model = model
        self.is_running = self.is_starting = False
        self.t = 0
        self.input = self.content = self.features = self.grams = None
        names = model.layers()
        self.weights = pd.DataFrame(np.ones((len(names), len(LOSS_NAMES))), names, LOSS_NAMES,
                                    np.float32)                 # worker.py:130-132
        self.params = {k: 1 for k in SCALAR_LOSS_NAMES}
        self.optimizer = None
        self.optimizer_cls = LBFGS
        self.step_size = DEFAULT_STEP['lbfgs']
        self.norms = {k: {} for k in 'cds'}
        self.traces = []

    # ---- state machine (worker.py:140-229)
    def check_consistency(self):
        return bool(self.input is not None and self.content is not None and self.grams
                    and self.input.shape == self.content.shape)

    def objective_changed(self):
        if self.optimizer is not None:
            self.optimizer.objective_changed()

    def pause(self):
        self.is_running = self.is_starting = False

    def reset(self):                                            # worker.py:172-175
        self.norms = {k: {} for k in self.norms}
        self.t = 0
        self.optimizer = self.optimizer_cls(self.input, self.opfunc, step_size=self.step_size)

    def start(self):
        self.is_starting = True
        self._maybe_run()
        return self.is_running

    def _maybe_run(self):                                       # worker.py:182-189
        if self.is_starting and self.check_consistency():
            if self.optimizer is None:
                self.reset()
            self.is_starting, self.is_running = False, True

    def set_input(self, image):                                 # worker.py:191-202
        image = self.model.preprocess(image)
        if self.input is not None and self.input.shape == image.shape:
            self.input[:] = image
            self.objective_changed()
        elif self.optimizer is not None:
            self.input = self.optimizer.resample(None, new_x=image)
            self._maybe_run()
        else:
            self.input = image
            self.reset()
            self._maybe_run()

    def _content_features(self):
        self.features = {k: v.copy() for k, v in self.model.forward(self.content).items()}
        self._maybe_run()
        self.objective_changed()

    def set_content(self, image):                               # worker.py:204-209
        self.content = self.model.preprocess(image)
        self._content_features()

    def set_style(self, image):                                 # worker.py:211-218
        feats = self.model.forward(self.model.preprocess(image))
        self.grams = {k: gram(v) for k, v in feats.items()}
        self._maybe_run()
        self.objective_changed()

    def resample_input(self, size):                             # worker.py:154-160
        if self.input is not None and self.optimizer is not None:
            self.input = self.optimizer.resample(size)
        else:
            self.input = np.zeros((1, 3) + tuple(size), np.float32)
        self._maybe_run()
        self.objective_changed()

    def resample_content(self, size):                           # worker.py:162-170
        if self.content is not None:
            self.content = nm.resample_nchw(self.content, size)
        else:
            self.content = np.zeros((1, 3) + tuple(size), np.float32)
        self._content_features()

    def set_step_size(self, step_size):                         # worker.py:220-224
        self.step_size = step_size
        if self.optimizer is not None:
            self.optimizer.step_size = step_size

    def set_weights(self, weights, params):                     # worker.py:226-229
        self.weights = pd.DataFrame.from_dict(weights, dtype=np.float32)
        self.params = params
        self.objective_changed()

    # ---- the objective (worker.py:231-301)
    def active_layers(self):
        """Rows of the weight table with any |w| > 1e-15, in table order (worker.py:234-235)."""
        hot = abs(self.weights) > EPS_W
        return list(self.weights.index[abs(hot.sum(axis=1)) > EPS_W])

    def opfunc(self, x, return_grad=True):
        layers = self.active_layers()
        tr = nm.TraceLog()
        feats = self.model.forward(x, layers)
        loss = 0
        diffs = {}
        for layer in layers:
            cw = self.weights['content'][layer]
            sw = self.weights['style'][layer]
            dw = self.weights['deepdream'][layer]
            cur = feats[layer]
            acc = np.zeros_like(cur)

            if abs(cw) > EPS_W:                                 # worker.py:249-256
                d = cur - self.features[layer]
                g = (2 / d.size) * d
                if layer not in self.norms['c']:
                    self.norms['c'][layer] = nm.rms(g)          # frozen until reset()
                cn = self.norms['c'][layer]
                loss += tr.put('%s_c_loss' % layer, cw * np.mean(d ** 2) / cn)
                acc += tr.put_rms('%s_c_grad' % layer, cw * g / cn)

            if abs(sw) > EPS_W:                                 # worker.py:258-269
                _, c, fh, fw = cur.shape
                gd = gram(cur) - self.grams[layer]
                flat = cur.reshape((c, fh * fw))
                g = np.dot(gd, flat).reshape((1, c, fh, fw))
                g *= 2 / (gd.size * flat.size)
                if layer not in self.norms['s']:
                    self.norms['s'][layer] = nm.rms(g)
                sn = self.norms['s'][layer]
                loss += tr.put('%s_s_loss' % layer, sw * np.mean(gd ** 2) / sn)
                tr.put_rms('%s_s_grad' % layer, sw / sn * g)
                nm.saxpy(sw / sn, g, acc)

            if abs(dw) > EPS_W:                                 # worker.py:271-277
                g = (-2 / cur.size) * cur
                if layer not in self.norms['d']:
                    self.norms['d'][layer] = nm.rms(g)
                dn = self.norms['d'][layer]
                loss += tr.put('%s_d_loss' % layer, -dw * np.mean(cur ** 2) / dn)
                acc += tr.put_rms('%s_d_grad' % layer, dw * g / dn)
            diffs[layer] = acc

        tr.put('scd_loss', loss)
        tv_loss, tv_grad = nm.total_variation(x / 255, self.params['tv_power'])    # worker.py:283
        loss += tr.put('t_loss', self.params['tv'] * tv_loss)
        p_loss, p_grad = nm.p_norm(x / 255, self.params['p_power'])                # worker.py:287
        loss += tr.put('p_loss', self.params['p'] * p_loss)

        if not return_grad:                                     # worker.py:290-292
            self.traces.append(tr)
            return tr.put('loss', loss)

        grad = tr.put_rms('scd_grad', self.model.backward(diffs).copy())           # worker.py:295
        grad += tr.put_rms('t_grad', self.params['tv'] * tv_grad)
        grad += tr.put_rms('p_grad', self.params['p'] * p_grad)
        tr.put('time', time.perf_counter())
        self.traces.append(tr)
        return tr.put('loss', loss), tr.put_rms('grad', grad)

    def step(self):                                             # worker.py:303-310
        self.t += 1
        x, _ = self.optimizer.step()
        tr = self.traces[-1]
        tr.put('fevals', self.t)
        return self.model.deprocess(x), tr.data
