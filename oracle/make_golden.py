#!/usr/bin/env python3
"""Generate ``tests/golden/*.npz`` by EXECUTING THE REFERENCE'S OWN PYTHON, unmodified.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python -m oracle.make_golden

``worker.StyleTransfer``, ``optimizers.*`` and ``utils.*`` are imported from
``/root/reference`` as they lie (nothing is copied).  The single thing the reference cannot do
here is ``worker.CaffeModel`` -- Caffe is not installable offline and ``vgg19.caffemodel`` is not
on disk -- so its model seam is served by ``oracle.caffe_cpu.CaffeCPUModel`` with seeded synthetic
weights.  Everything stored below is therefore: reference objective + reference optimizers +
reference numeric utils, on top of the restated Caffe layers.

The fixtures are the pin for ``oracle/transfer.py`` and ``oracle/numeric.py``
(``tests/test_oracle_golden.py``) and the known-answer vectors for the CUDA path
(``tests/test_gpu_*.py``).
"""
import os
import sys

import numpy as np

REF = os.environ.get('ST2_REFERENCE_DIR', '/root/reference')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden')


def import_reference():
    sys.dont_write_bytecode = True            # the reference tree is read-only
    sys.path.insert(0, REF)
    argv, sys.argv = sys.argv, ['worker']
    import worker as ref_worker                # noqa: E402  (first: the reference's import cycle
    import optimizers as ref_opt               # noqa: E402   messages<->optimizers<->utils only
    import utils as ref_utils                  # noqa: E402   resolves when entered via worker)
    sys.argv = argv
    sys.excepthook = sys.__excepthook__
    return ref_worker, ref_opt, ref_utils


def load_images(ref_utils, size, square=False):
    """The config's two example images, sized as app.py does (``resize_to_fit``)."""
    from PIL import Image
    content = Image.open(os.path.join(REF, 'examples', 'golden_gate.jpg')).convert('RGB')
    style = Image.open(os.path.join(REF, 'examples', 'starry_night.jpg')).convert('RGB')
    if square:
        content = content.resize((size, size), Image.LANCZOS)
    else:
        content = ref_utils.resize_to_fit(content, size)
    style = ref_utils.resize_to_fit(style, size)
    return np.uint8(content), np.uint8(style)


def stock_weights():
    import yaml
    with open(os.path.join(REF, 'initial_weights.yaml')) as f:
        w, p = yaml.safe_load(f)
    return w, p


def flat_trace(traces, keys=None):
    keys = keys or [k for k in traces[0] if k != 'time']
    return keys, np.array([[t[k] for k in keys] for t in traces], np.float64)


def numeric_fixture(ref_utils, ref_opt):
    rs = np.random.RandomState(1234)
    out = {}
    x = (rs.randn(1, 3, 13, 17) * 0.4).astype(np.float32)
    for beta in (2, 1.5, 3):
        n, g = ref_utils.tv_norm(x.copy(), beta)
        out['tv_b%s_norm' % beta] = np.float64(n)
        out['tv_b%s_grad' % beta] = g
    for p in (2, 6, 3):
        n, g = ref_utils.p_norm(x.copy(), p)
        out['pn_p%s_norm' % p] = np.float64(n)
        out['pn_p%s_grad' % p] = g
    out['x'] = x
    a = rs.randn(1, 3, 9, 11).astype(np.float32)
    b = rs.randn(1, 3, 9, 11).astype(np.float32)
    out['a'], out['b'] = a, b
    out['dot_ab'] = np.float64(ref_utils.dot(a, b))
    out['axpy_ab'] = ref_utils.axpy(0.37, a, b.copy())
    dm = ref_utils.DecayingMean(0.9)
    seq = [dm(a * (i + 1)) for i in range(4)]
    out['ema_seq'] = np.stack(seq)
    # Pillow resampling through the reference's own entry point
    from PIL import Image
    plane = (rs.rand(1, 2, 37, 53) * 255).astype(np.float32)
    out['rs_in'] = plane
    for tag, hw in (('up2', (74, 106)), ('upsqrt2', (52, 75)), ('down2', (18, 26)),
                    ('downsqrt2', (26, 37)), ('same', (37, 53))):
        out['rs_lanczos_' + tag] = ref_utils.resample_nchw(plane, hw)
        out['rs_bilinear_' + tag] = ref_utils.resample_nchw(plane, hw, method=Image.BILINEAR)
    out['scales_300_200'] = np.array(ref_utils.scales((300, 200), 32), np.int64)
    out['fit'] = np.array([ref_utils.fit_into_square((979, 734), 256, True),
                           ref_utils.fit_into_square((1024, 640), 256, True),
                           ref_utils.fit_into_square((100, 80), 256, False)], np.int64)
    np.savez_compressed(os.path.join(OUT, 'numeric.npz'), **out)
    print('numeric.npz', len(out), 'arrays')


def small_fixture(ref_worker, ref_opt, ref_utils):
    """40x56 canvas (odd pooled extents exercise ceil-mode), rich weights incl. deepdream, a pool
    layer and data; opfunc at x0; 13 L-BFGS steps with the complete optimizer state before the
    last one (teacher forcing); 13 Adam steps."""
    from oracle.caffe_cpu import CaffeCPUModel
    from PIL import Image
    H, W = 40, 56
    content = np.uint8(Image.open(os.path.join(REF, 'examples', 'golden_gate.jpg')).convert('RGB')
                       .resize((W, H), Image.LANCZOS))
    style = np.uint8(Image.open(os.path.join(REF, 'examples', 'starry_night.jpg')).convert('RGB')
                     .resize((W + 8, H - 6), Image.LANCZOS))
    x0 = np.uint8(np.random.RandomState(0).uniform(0, 255, (H, W, 3)))
    weights = {'content': {'conv4_2': 0.08, 'pool2': 0.02, 'data': 0.01},
               'style': {'conv1_1': 1, 'conv2_1': 1, 'conv3_1': 1, 'conv4_1': 1, 'conv5_1': 0.5},
               'deepdream': {'conv3_2': 0.03}}
    params = {'p': 50, 'p_power': 6, 'tv': 5, 'tv_power': 2}
    out = {'content': content, 'style': style, 'x0': x0,
           'weights_repr': np.array(repr(weights)), 'params_repr': np.array(repr(params))}

    def fresh(opt_name):
        st = ref_worker.StyleTransfer(CaffeCPUModel())
        if opt_name == 'adam':
            st.optimizer_cls, st.step_size = ref_opt.AdamOptimizer, 10
        st.set_input(x0)
        st.set_content(content)
        st.set_style(style)
        st.set_weights(weights, params)
        assert st.start()
        return st

    # --- single evaluation at x0 with everything it produced
    st = fresh('lbfgs')
    loss, grad = st.opfunc(st.input)
    out['eval_layers'] = np.array(list(st.weights.index[(abs(st.weights) > 1e-15).any(axis=1)]))
    out['eval_loss'] = np.float64(loss)
    out['eval_grad'] = grad
    keys, vals = flat_trace([st.traces[-1].data])
    out['eval_trace_keys'], out['eval_trace'] = np.array(keys), vals[0]
    for kind in 'cds':
        for layer, v in st.norms[kind].items():
            out['eval_norm_%s_%s' % (kind, layer)] = np.float64(v)
    feats = st.model.forward(st.input)
    for layer in ('conv1_1', 'conv3_1', 'pool2', 'conv4_2', 'conv5_1', 'pool5'):
        out['feat_' + layer] = feats[layer].copy()
    for layer in ('conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1'):
        out['gram_style_' + layer] = st.grams[layer]
    out['eval_loss_only'] = np.float64(st.opfunc(st.input, return_grad=False))

    # --- L-BFGS trajectory with a full state checkpoint before the final step
    st = fresh('lbfgs')
    xs, traces = [], []
    n_steps = 13
    for k in range(n_steps):
        if k == n_steps - 1:
            o = st.optimizer
            out['ck_x'] = st.input.copy()
            out['ck_S'] = np.stack(o.sk)
            out['ck_Y'] = np.stack(o.yk)
            out['ck_SY'] = np.array(o.syk, np.float64)
            out['ck_grad'] = o.grad.copy()
            out['ck_loss'] = np.float64(o.loss)
            for kind in 'cds':
                for layer, v in st.norms[kind].items():
                    out['ck_norm_%s_%s' % (kind, layer)] = np.float64(v)
        img, tr = st.step()
        xs.append(st.input.copy())
        traces.append(dict(tr))
    keys, vals = flat_trace(traces)
    out['lbfgs_trace_keys'], out['lbfgs_trace'] = np.array(keys), vals
    out['lbfgs_x'] = np.stack(xs)
    out['lbfgs_final_S'] = np.stack(st.optimizer.sk)
    out['lbfgs_final_SY'] = np.array(st.optimizer.syk, np.float64)
    out['lbfgs_image_last'] = np.float32(img)

    # --- Adam trajectory
    st = fresh('adam')
    xs, traces = [], []
    for k in range(n_steps):
        img, tr = st.step()
        xs.append(st.input.copy())
        traces.append(dict(tr))
    keys, vals = flat_trace(traces)
    out['adam_trace_keys'], out['adam_trace'] = np.array(keys), vals
    out['adam_x'] = np.stack(xs)
    out['adam_m1'] = np.float32(st.optimizer.g1.mean)
    out['adam_m2'] = np.float32(st.optimizer.g2.mean)
    # scale change through the optimizer seam (optimizers.py:29-40)
    xr = st.optimizer.resample((60, 84))
    out['adam_rs_x'] = xr.copy()
    out['adam_rs_m1'] = np.float32(st.optimizer.g1.mean)
    out['adam_rs_m2'] = np.float32(st.optimizer.g2.mean)
    np.savez_compressed(os.path.join(OUT, 'small.npz'), **out)
    print('small.npz', len(out), 'arrays')


def config1_fixture(ref_worker, ref_opt, ref_utils):
    """BASELINE config 1: golden_gate + starry_night at 256 px, stock YAML, 100 L-BFGS steps."""
    from oracle.caffe_cpu import CaffeCPUModel
    content, style = load_images(ref_utils, 256)
    H, W = content.shape[:2]
    x0 = np.uint8(np.random.RandomState(0).uniform(0, 255, (H, W, 3)))      # mirrors app.py:251
    weights, params = stock_weights()
    st = ref_worker.StyleTransfer(CaffeCPUModel())
    st.set_input(x0)
    st.set_content(content)
    st.set_style(style)
    st.set_weights(weights, params)
    assert st.start()
    out = {'content': content, 'style': style, 'x0': x0,
           'weights_repr': np.array(repr(weights)), 'params_repr': np.array(repr(params))}
    traces = []
    for k in range(1, 101):
        img, tr = st.step()
        traces.append(dict(tr))
        if k in (1, 2, 5):
            out['image_%03d' % k] = np.float32(img)
        if k in (10, 100):
            out['image_u8_%03d' % k] = np.uint8(np.clip(img, 0, 255))
        if k % 20 == 0:
            print('  config1 step', k, 'loss %.6g' % tr['loss'])
    keys, vals = flat_trace(traces)
    out['trace_keys'], out['trace'] = np.array(keys), vals
    for kind in 'cds':
        for layer, v in st.norms[kind].items():
            out['norm_%s_%s' % (kind, layer)] = np.float64(v)
    np.savez_compressed(os.path.join(OUT, 'config1.npz'), **out)
    print('config1.npz', len(out), 'arrays')


def main():
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, ROOT)
    ref_worker, ref_opt, ref_utils = import_reference()
    which = sys.argv[1:] or ['numeric', 'small', 'config1']
    if 'numeric' in which:
        numeric_fixture(ref_utils, ref_opt)
    if 'small' in which:
        small_fixture(ref_worker, ref_opt, ref_utils)
    if 'config1' in which:
        config1_fixture(ref_worker, ref_opt, ref_utils)


if __name__ == '__main__':
    main()
