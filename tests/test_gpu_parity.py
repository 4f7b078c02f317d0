"""GPU parity tests: the CUDA path (through the C ABI in libst2.so) against the oracle and the
golden vectors produced by the reference's own Python.  Run with ``-m gpu`` on a B200.

Tolerances (stated per BASELINE.json's north_star):
  * fp32 conv path: features / losses / traces within 1e-4 relative, gradients within 1e-3
    (the fp32-vs-fp64 floor of the gradient is 5.5e-4);
  * fp16 tensor-core path: per-layer features and losses within 1e-3 relative (the north-star
    bound); gradients within 5e-2 relative (operand rounding, SURVEY appendix B measured 1.9e-2).
"""
import ast

import numpy as np
import pytest
import torch

from conftest import rel_err, psnr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def models():
    from style_transfer2_b200.model import B200Model
    cache = {}

    def get(precision):
        if precision not in cache:
            cache[precision] = B200Model(precision=precision)
        from style_transfer2_b200 import utils
        utils.set_default_engine(cache[precision].engine)
        return cache[precision]
    return get


def dev(a, model):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(model.engine.device)


# ------------------------------------------------------------------------------- small kernels
def test_pixel_terms_match_reference_utils(golden, models):
    from style_transfer2_b200 import utils
    m = models('fp32')
    g = golden('numeric')
    x = dev(g['x'], m)
    for beta in ('2', '1.5', '3'):
        n, grad = utils.tv_norm(x, float(beta))
        assert np.isclose(n, g['tv_b%s_norm' % beta], rtol=2e-5)
        np.testing.assert_allclose(grad.cpu().numpy(), g['tv_b%s_grad' % beta], rtol=2e-4, atol=2e-6)
    for p in (2, 6, 3):
        n, grad = utils.p_norm(x, p)
        assert np.isclose(n, g['pn_p%s_norm' % p], rtol=2e-5)
        np.testing.assert_allclose(grad.cpu().numpy(), g['pn_p%s_grad' % p], rtol=2e-5, atol=1e-8)


def test_level1_and_shape_errors(golden, models):
    from style_transfer2_b200 import utils
    m = models('fp32')
    g = golden('numeric')
    a, b = dev(g['a'], m), dev(g['b'], m)
    assert np.isclose(utils.dot(a, b), g['dot_ab'], rtol=1e-5)
    y = b.clone()
    assert utils.axpy(0.37, a, y) is y
    np.testing.assert_allclose(y.cpu().numpy(), g['axpy_ab'], rtol=1e-6, atol=1e-7)
    with pytest.raises(ValueError):
        utils.dot(a, b[:, :2])
    with pytest.raises(ValueError):
        utils.axpy(1.0, a, b[:, :2])


@pytest.mark.parametrize('tag,hw', [('up2', (74, 106)), ('upsqrt2', (52, 75)), ('down2', (18, 26)),
                                    ('downsqrt2', (26, 37)), ('same', (37, 53))])
def test_resample_matches_pillow(golden, models, tag, hw):
    from style_transfer2_b200 import utils
    m = models('fp32')
    g = golden('numeric')
    src = dev(g['rs_in'], m)
    for name, method in (('lanczos', utils.LANCZOS), ('bilinear', utils.BILINEAR)):
        got = utils.resample_nchw(src, hw, method).cpu().numpy()
        np.testing.assert_allclose(got, g['rs_%s_%s' % (name, tag)], rtol=0, atol=6e-5)


def test_pre_deprocess_round_trip(models):
    from style_transfer2_b200.worker import StyleTransfer
    m = models('fp32')
    st = StyleTransfer(m)
    img = np.uint8(np.random.RandomState(3).uniform(0, 255, (9, 13, 3)))
    x = st._upload_image(img)
    np.testing.assert_allclose(x.cpu().numpy(), m.preprocess(img), atol=1e-5)
    np.testing.assert_allclose(st.image(x), np.float32(img), atol=1e-4)
    xf = st._upload_image(np.float32(img) + 0.25)
    np.testing.assert_allclose(xf.cpu().numpy(), m.preprocess(np.float32(img) + 0.25), atol=1e-5)


# ------------------------------------------------------------------------------- model seam
FEATS = ('conv1_1', 'conv3_1', 'pool2', 'conv4_2', 'conv5_1', 'pool5')


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('fp16', 1e-3)])
def test_forward_features_match_golden(golden, models, precision, tol):
    g = golden('small')
    m = models(precision)
    x = m.preprocess(g['x0'])
    feats = m.forward(x, FEATS)
    for name in FEATS:
        err = rel_err(feats[name], g['feat_' + name])
        assert err < tol, (name, err)


def test_forward_odd_sizes_ceil_mode(models):
    from oracle.caffe_cpu import CaffeCPUModel
    rs = np.random.RandomState(5)
    x = (rs.rand(1, 3, 75, 101) * 255 - 120).astype(np.float32)
    want = CaffeCPUModel().forward(x)
    for precision, tol in (('fp32', 1e-4), ('fp16', 1e-3)):
        got = models(precision).forward(x)
        assert list(got) == list(want)
        for name in want:
            assert got[name].shape == want[name].shape
            assert rel_err(got[name], want[name]) < tol, (precision, name)


@pytest.mark.parametrize('layers', [['conv2_1'], ['conv4_2', 'conv1_1', 'conv3_1'], ['pool2', 'conv3_2', 'data'],
                                    ['pool5', 'conv5_4'], ['data'], ['conv1_1']])
@pytest.mark.parametrize('precision,tol', [('fp32', 2e-4), ('fp16', 1e-2)])
def test_backward_segment_semantics(models, layers, precision, tol):
    """Injected diffs enter below the layer's own ReLU; gradient from above is masked.
    The data gradient is DISCONTINUOUS in the features (ReLU masks, pool arg-max): a fraction f of
    flipped decisions costs ~sqrt(f) relative error, so the fp16 path is checked against the oracle
    with the same operand rounding (semantics), and against the fp32 oracle only loosely."""
    from oracle.caffe_cpu import CaffeCPUModel
    rs = np.random.RandomState(11)
    x = (rs.rand(1, 3, 37, 45) * 255 - 120).astype(np.float32)
    ref = CaffeCPUModel(emulate_fp16=(precision == 'fp16'))
    feats = ref.forward(x)
    diffs = {l: rs.randn(*feats[l].shape).astype(np.float32) for l in layers}
    want = ref.backward(diffs)
    m = models(precision)
    m.forward(x, layers)
    got = m.backward(diffs)
    if precision == 'fp16':
        # tensor-core accumulation order differs from the CPU's even with identical operands; every
        # layer the diff crosses adds flipped mask / arg-max decisions (measured: 2e-2 through
        # conv4_2, 9e-2 from pool5 with white-noise diffs) -- bound grows with depth.  Since conv1_1 runs on
        # the tensor cores too (hi/lo split operands, fp32-grade but a different summation order), its fp16
        # outputs differ from the oracle's in the last bit here and there; at fp16 resolution 2x2 pool windows
        # hold ties, so a few first-maximum decisions move (measured 5.1e-2 for pool2 + conv3_2 + data, and
        # 2e-4 between our own two conv1_1 implementations, tools/debug_first.py)
        from style_transfer2_b200 import vgg
        depth = max(vgg.BLOB_INDEX[l] for l in layers)
        tol = 1e-2 if depth <= 4 else (7e-2 if depth <= 13 else 0.15)
    assert rel_err(got, want) < tol, rel_err(got, want)


@pytest.mark.parametrize('hw', [(40, 56), (75, 101)])
def test_tcgen05_gram_matches_numpy_on_the_same_features(models, hw):
    """Gram kernel (tcgen05, MN-major operands, split-K) vs X X^T / size computed in NumPy from the
    exported fp16-valued features: same operands, so only the accumulation order differs."""
    from oracle.transfer import gram
    from style_transfer2_b200 import vgg
    m = models('fp16')
    rs = np.random.RandomState(4)
    x = (rs.rand(1, 3, *hw) * 255 - 120).astype(np.float32)
    names = ['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1', 'pool1', 'pool4']
    feats = m.forward(x, names)
    plan = m._last_plan
    for name in names:
        got = plan.gram(vgg.BLOB_INDEX[name]).cpu().numpy()
        want = gram(feats[name])
        assert rel_err(got, want) < 2e-5, (name, rel_err(got, want))


def test_reference_objective_runs_on_our_model_seam(golden, models):
    """Drop-in at the model seam: the oracle's restatement of worker.StyleTransfer (NumPy, host)
    drives B200Model.forward/backward exactly as it would drive CaffeModel."""
    from oracle.transfer import Transfer
    g = golden('small')
    st = Transfer(models('fp32'))
    st.set_input(g['x0'])
    st.set_content(g['content'])
    st.set_style(g['style'])
    st.set_weights(ast.literal_eval(str(g['weights_repr'])), ast.literal_eval(str(g['params_repr'])))
    assert st.start()
    loss, grad = st.opfunc(st.input)
    assert np.isclose(loss, g['eval_loss'], rtol=1e-4)
    assert rel_err(grad, g['eval_grad']) < 1e-3


# ------------------------------------------------------------------------------- objective
def _transfer(g, model, opt='lbfgs'):
    from style_transfer2_b200.worker import StyleTransfer
    from style_transfer2_b200 import optimizers
    st = StyleTransfer(model)
    if opt == 'adam':
        st.optimizer_cls, st.step_size = optimizers.AdamOptimizer, 10
    st.set_input(g['x0'])
    st.set_content(g['content'])
    st.set_style(g['style'])
    st.set_weights(ast.literal_eval(str(g['weights_repr'])), ast.literal_eval(str(g['params_repr'])))
    assert st.start()
    return st


@pytest.mark.parametrize('precision,tol_loss,tol_grad', [('fp32', 1e-4, 1e-3), ('fp16', 1e-3, 5e-2)])
def test_objective_single_eval(golden, models, precision, tol_loss, tol_grad):
    g = golden('small')
    st = _transfer(g, models(precision))
    assert st.active_layers() == list(g['eval_layers'])
    loss, grad = st.opfunc(st.input)
    tr = st.traces[-1].data
    keys = [k for k in tr if k != 'time']
    assert keys == list(g['eval_trace_keys'])
    got = np.array([tr[k] for k in keys])
    want = g['eval_trace']
    for k, a, b in zip(keys, got, want):
        tol = tol_grad if k.endswith('grad') else tol_loss
        assert abs(a - b) <= tol * abs(b) + 1e-30, (k, a, b)
    assert abs(float(loss) - g['eval_loss']) <= tol_loss * abs(g['eval_loss'])
    assert rel_err(grad.cpu().numpy(), g['eval_grad']) < tol_grad
    norms = st.norms
    for key in g.files:
        if key.startswith('eval_norm_'):
            _, _, kind, layer = key.split('_', 3)
            assert abs(norms[kind][layer] - float(g[key])) <= max(tol_loss, 2e-3 if precision == 'fp16' else 0) * float(g[key])
    lo = st.opfunc(st.input, return_grad=False)
    assert abs(float(lo) - g['eval_loss_only']) <= tol_loss * abs(g['eval_loss_only'])
    assert list(st.traces[-1].data)[-1] == 'loss'


def test_lbfgs_teacher_forced_step(golden, models):
    """Reference optimizer state before its 13th step -> one CUDA step -> same x, trace, new pair."""
    g = golden('small')
    st = _transfer(g, models('fp32'))
    st.input.copy_(dev(g['ck_x'], st.model))
    norms = {k: {} for k in 'cds'}
    for key in g.files:
        if key.startswith('ck_norm_'):
            _, _, kind, layer = key.split('_', 3)
            norms[kind][layer] = float(g[key])
    st.set_norms(norms)
    st.optimizer.load_state(g['ck_S'], g['ck_Y'], list(g['ck_SY']), g['ck_grad'], float(g['ck_loss']))
    st.t = 12
    img, tr = st.step()
    assert rel_err(st.input.cpu().numpy(), g['lbfgs_x'][12]) < 2e-4
    keys = list(g['lbfgs_trace_keys'])
    for k, want in zip(keys, g['lbfgs_trace'][12]):
        tol = 2e-3 if k.endswith('grad') else 5e-4
        assert abs(tr[k] - want) <= tol * abs(want) + 1e-30, (k, tr[k], want)
    S, Y, sy = st.optimizer.export_state()
    assert len(sy) == 10
    assert rel_err(S.cpu().numpy(), g['lbfgs_final_S']) < 5e-4
    np.testing.assert_allclose(sy, g['lbfgs_final_SY'], rtol=5e-3)
    assert rel_err(img, g['lbfgs_image_last']) < 2e-4


def test_lbfgs_short_free_run(golden, models):
    g = golden('small')
    st = _transfer(g, models('fp32'))
    for k in range(3):
        st.step()
        assert rel_err(st.input.cpu().numpy(), g['lbfgs_x'][k]) < 1e-3, k


@pytest.mark.parametrize('precision,min_psnr', [('fp32', 45.0), ('fp16', 35.0)])
def test_adam_free_run(golden, models, precision, min_psnr):
    """Long-horizon image check on the stable optimizer (SURVEY appendix C protocol)."""
    g = golden('small')
    st = _transfer(g, models(precision), 'adam')
    mean = st.model.mean
    for k in range(13):
        img, tr = st.step()
    want_img = (g['adam_x'][12].squeeze() + mean).transpose(1, 2, 0)
    assert psnr(img, want_img) > min_psnr
    if precision == 'fp32':
        assert rel_err(st.input.cpu().numpy(), g['adam_x'][12]) < 2e-3
        assert rel_err(st.optimizer.g1.mean.cpu().numpy(), g['adam_m1']) < 2e-3
        assert rel_err(st.optimizer.g2.mean.cpu().numpy(), g['adam_m2']) < 2e-3
        st.optimizer.resample((60, 84))
        assert rel_err(st.optimizer.x.cpu().numpy(), g['adam_rs_x']) < 2e-3
        assert rel_err(st.optimizer.g2.mean.cpu().numpy(), g['adam_rs_m2']) < 5e-3


@pytest.mark.parametrize('precision,min_psnr', [('fp32', 60.0), ('fp16', 35.0)])
def test_config1_first_steps(golden, models, precision, min_psnr):
    """BASELINE config 1 (256 px, stock YAML, L-BFGS): the head of the reference trajectory."""
    g = golden('config1')
    st = _transfer(g, models(precision))
    keys = list(g['trace_keys'])
    for k in (1, 2):
        img, tr = st.step()
        assert psnr(img, g['image_%03d' % k]) > min_psnr
        if precision == 'fp32':
            for kk, want in zip(keys, g['trace'][k - 1]):
                tol = 5e-3 if kk.endswith('grad') else 1e-3
                assert abs(tr[kk] - want) <= tol * abs(want) + 1e-30, (k, kk, tr[kk], want)
    assert [kk for kk in tr if kk != 'time'] == keys


@pytest.mark.parametrize('precision', ['fp32', 'fp16'])
def test_config1_hundred_lbfgs_steps_stay_finite_and_land_where_the_reference_does(golden, models, precision):
    """BASELINE config 1 end to end: 100 free-running L-BFGS steps.  The reference's fixed-step L-BFGS
    overshoots (loss 2.8e8 -> 9.7e14 at step 7) and recovers; it is chaotic, two correct CPU runs agree to
    ~20-23 dB only.  Stated bound: every loss finite (fp16 stores saturate instead of
    overflowing), final loss within 2x of the reference's, final image >= 15 dB from the reference's."""
    g = golden('config1')
    st = _transfer(g, models(precision))
    keys = list(g['trace_keys'])
    want = g['trace'][-1][keys.index('loss')]
    losses = []
    for k in range(100):
        img, tr = st.step()
        losses.append(tr['loss'])
    assert np.isfinite(losses).all(), losses
    assert 0.5 * want < losses[-1] < 2.0 * want, (losses[-1], want)
    db = psnr(img, g['image_u8_100'])
    print('config1 %s: final loss %.4g (reference %.4g), PSNR vs reference %.1f dB, peak loss %.3g' % (
        precision, losses[-1], want, db, max(losses)))
    assert db > 15.0


@pytest.mark.parametrize('precision,min_psnr', [('fp32', 45.0), ('fp16', 33.0)])
def test_config3_multiscale_adam_follows_the_oracle(golden, models, precision, min_psnr):
    """BASELINE config 3 in miniature: Adam (step 10), three stages joined by the worker's
    SetImages(size, RESAMPLE, RESAMPLE) path (worker.py:154-170, optimizers.py:29-40): x and the first moment
    Lanczos-resampled, the second moment bilinear + clamp, normalisers persist across stages.  The oracle runs
    the same message sequence on the CPU."""
    from oracle.caffe_cpu import CaffeCPUModel
    from oracle.transfer import Transfer, Adam
    g = golden('small')
    weights, params = ast.literal_eval(str(g['weights_repr'])), ast.literal_eval(str(g['params_repr']))
    ora = Transfer(CaffeCPUModel())
    ora.optimizer_cls, ora.step_size = Adam, 10
    ora.set_input(g['x0'])
    ora.set_content(g['content'])
    ora.set_style(g['style'])
    ora.set_weights(weights, params)
    assert ora.start()
    st = _transfer(g, models(precision), 'adam')
    h, w = g['x0'].shape[:2]
    for stage, size in enumerate(((h, w), (int(h * 1.4), int(w * 1.4)), (h * 2, w * 2))):
        if stage:
            for t in (ora, st):
                t.resample_input(size)
                t.resample_content(size)
            assert tuple(st.input.shape[2:]) == size
            assert st.norms['s'].keys() == ora.norms['s'].keys()            # frozen normalisers survive the resize
            for k, v in ora.norms['s'].items():
                assert np.isclose(st.norms['s'][k], v, rtol=2e-3), (k, st.norms['s'][k], v)
        for _ in range(5):
            img_o, tr_o = ora.step()
            img, tr = st.step()
        assert psnr(img, img_o) > min_psnr, (stage, psnr(img, img_o))
        assert np.isclose(tr['loss'], tr_o['loss'], rtol=1e-3 if precision == 'fp32' else 2e-2), (stage, tr['loss'], tr_o['loss'])


def test_gram_matrix_device(models):
    from style_transfer2_b200.worker import gram_matrix
    from oracle.transfer import gram
    m = models('fp32')
    rs = np.random.RandomState(2)
    x = rs.randn(1, 70, 9, 11).astype(np.float32)
    got = gram_matrix(dev(x, m)).cpu().numpy()
    np.testing.assert_allclose(got, gram(x), rtol=1e-4, atol=1e-6)
    with pytest.raises(AssertionError):
        gram_matrix(dev(np.zeros((2, 3, 4, 4)), m))
