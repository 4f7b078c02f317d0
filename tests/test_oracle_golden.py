"""Pins the oracle (oracle/numeric.py, oracle/transfer.py) against vectors produced by the
reference's own Python (tests/golden/*.npz, written by oracle/make_golden.py).  CPU only."""
import ast

import numpy as np
import pytest

from oracle import numeric as nm
from oracle.caffe_cpu import CaffeCPUModel
from oracle.transfer import Transfer, Adam, LBFGS, gram
from conftest import rel_err, psnr


def test_tv_and_pnorm(golden):
    g = golden('numeric')
    x = g['x']
    for beta in ('2', '1.5', '3'):
        n, grad = nm.total_variation(x.copy(), float(beta))
        assert np.isclose(n, g['tv_b%s_norm' % beta], rtol=1e-6)
        np.testing.assert_allclose(grad, g['tv_b%s_grad' % beta], rtol=1e-5, atol=1e-7)
    for p in (2, 6, 3):
        n, grad = nm.p_norm(x.copy(), p)
        assert np.isclose(n, g['pn_p%s_norm' % p], rtol=1e-6)
        np.testing.assert_allclose(grad, g['pn_p%s_grad' % p], rtol=1e-6, atol=1e-9)


def test_level1_and_ema(golden):
    g = golden('numeric')
    a, b = g['a'], g['b']
    assert np.isclose(nm.sdot(a, b), g['dot_ab'], rtol=1e-5)
    np.testing.assert_allclose(nm.saxpy(0.37, a, b.copy()), g['axpy_ab'], rtol=1e-6, atol=1e-7)
    with pytest.raises(ValueError):
        nm.sdot(a, b[:, :2])
    with pytest.raises(ValueError):
        nm.saxpy(1.0, a, b[:, :2])
    ema = nm.EMA(0.9)
    assert ema.value() == 0
    seq = [ema.update(a * (i + 1)) for i in range(4)]
    np.testing.assert_allclose(np.stack(seq), g['ema_seq'], rtol=1e-6)


@pytest.mark.parametrize('tag,hw', [('up2', (74, 106)), ('upsqrt2', (52, 75)), ('down2', (18, 26)),
                                    ('downsqrt2', (26, 37)), ('same', (37, 53))])
def test_resample_matches_pillow(golden, tag, hw):
    g = golden('numeric')
    src = g['rs_in']
    for method in ('lanczos', 'bilinear'):
        got = nm.resample_nchw(src, hw, method)
        want = g['rs_%s_%s' % (method, tag)]
        # Pillow accumulates in fp64 and rounds once; allow 1 ulp of fp32 at 255-scale
        np.testing.assert_allclose(got, want, rtol=0, atol=4e-5)


def test_sizes(golden):
    g = golden('numeric')
    assert [tuple(r) for r in g['scales_300_200']] == nm.scale_pyramid((300, 200), 32)
    assert tuple(g['fit'][0]) == nm.fit_into_square((979, 734), 256, True)
    assert tuple(g['fit'][1]) == nm.fit_into_square((1024, 640), 256, True)
    assert tuple(g['fit'][2]) == nm.fit_into_square((100, 80), 256, False)


def _small_transfer(g, opt='lbfgs'):
    st = Transfer(CaffeCPUModel())
    if opt == 'adam':
        st.optimizer_cls, st.step_size = Adam, 10
    st.set_input(g['x0'])
    st.set_content(g['content'])
    st.set_style(g['style'])
    st.set_weights(ast.literal_eval(str(g['weights_repr'])), ast.literal_eval(str(g['params_repr'])))
    assert st.start()
    return st


def test_objective_single_eval(golden):
    g = golden('small')
    st = _small_transfer(g)
    assert st.active_layers() == list(g['eval_layers'])
    loss, grad = st.opfunc(st.input)
    assert np.isclose(loss, g['eval_loss'], rtol=1e-5)
    assert rel_err(grad, g['eval_grad']) < 1e-5
    tr = st.traces[-1].data
    keys = [k for k in tr if k != 'time']
    assert keys == list(g['eval_trace_keys'])
    np.testing.assert_allclose([tr[k] for k in keys], g['eval_trace'], rtol=2e-5)
    for kind in 'cds':
        for layer, v in st.norms[kind].items():
            assert np.isclose(v, g['eval_norm_%s_%s' % (kind, layer)], rtol=1e-5)
    assert np.isclose(st.opfunc(st.input, return_grad=False), g['eval_loss_only'], rtol=1e-5)
    for layer in ('conv1_1', 'conv3_1', 'conv5_1'):
        np.testing.assert_allclose(st.grams[layer], g['gram_style_' + layer], rtol=1e-5, atol=1e-6)


def test_lbfgs_teacher_forced_step(golden):
    """Load the reference optimizer's complete state before its 13th step, take that one step."""
    g = golden('small')
    st = _small_transfer(g)
    st.input[:] = g['ck_x']
    for key in g.files:
        if key.startswith('ck_norm_'):
            _, _, kind, layer = key.split('_', 3)
            st.norms[kind][layer] = float(g[key])
    o = st.optimizer
    o.S = [s.copy() for s in g['ck_S']]
    o.Y = [y.copy() for y in g['ck_Y']]
    o.SY = list(g['ck_SY'])
    o.grad, o.loss = g['ck_grad'].copy(), float(g['ck_loss'])
    st.t = 12
    img, tr = st.step()
    assert rel_err(st.input, g['lbfgs_x'][12]) < 1e-5
    keys = list(g['lbfgs_trace_keys'])
    np.testing.assert_allclose([tr[k] for k in keys], g['lbfgs_trace'][12], rtol=1e-4)
    assert len(o.S) == 10
    assert rel_err(np.stack(o.S), g['lbfgs_final_S']) < 1e-5
    np.testing.assert_allclose(o.SY, g['lbfgs_final_SY'], rtol=1e-4)
    assert rel_err(img, g['lbfgs_image_last']) < 1e-5


def test_lbfgs_short_free_run(golden):
    g = golden('small')
    st = _small_transfer(g)
    for k in range(3):
        st.step()
        assert rel_err(st.input, g['lbfgs_x'][k]) < 1e-4, k


def test_adam_free_run_and_resample(golden):
    g = golden('small')
    st = _small_transfer(g, 'adam')
    for k in range(13):
        _, tr = st.step()
        assert rel_err(st.input, g['adam_x'][k]) < 1e-4, k
    keys = list(g['adam_trace_keys'])
    np.testing.assert_allclose([tr[k] for k in keys], g['adam_trace'][12], rtol=1e-3)
    assert rel_err(st.optimizer.m1.mean, g['adam_m1']) < 1e-4
    assert rel_err(st.optimizer.m2.mean, g['adam_m2']) < 1e-4
    xr = st.optimizer.resample((60, 84))
    assert rel_err(xr, g['adam_rs_x']) < 1e-6
    assert rel_err(st.optimizer.m1.mean, g['adam_rs_m1']) < 1e-4
    assert rel_err(st.optimizer.m2.mean, g['adam_rs_m2']) < 1e-4


def test_config1_head(golden):
    """BASELINE config 1 (256 px, stock YAML, L-BFGS): first steps of the reference trajectory."""
    g = golden('config1')
    st = Transfer(CaffeCPUModel())
    st.set_input(g['x0'])
    st.set_content(g['content'])
    st.set_style(g['style'])
    st.set_weights(ast.literal_eval(str(g['weights_repr'])), ast.literal_eval(str(g['params_repr'])))
    assert st.start()
    keys = list(g['trace_keys'])
    for k in (1, 2):
        img, tr = st.step()
        assert psnr(img, g['image_%03d' % k]) > 60
        np.testing.assert_allclose([tr[kk] for kk in keys], g['trace'][k - 1], rtol=1e-3)
    assert keys[:4] == ['conv4_2_c_loss', 'conv4_2_c_grad', 'conv1_1_s_loss', 'conv1_1_s_grad']
    assert keys[-3:] == ['loss', 'grad', 'fevals']
