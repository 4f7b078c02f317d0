"""Cross-checks of the restated Caffe layer semantics (oracle/caffe_cpu.py) against independent
implementations: torch autograd with the diff injected *below* each layer's own ReLU, and
closed-form pooled extents.  The reference itself holds no test for this boundary."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.caffe_cpu import CaffeCPUModel, TOPOLOGY, BLOB_NAMES, pool_out, synthetic_weights


def test_blob_names_match_prototxt_order():
    assert BLOB_NAMES[:4] == ['data', 'conv1_1', 'conv1_2', 'pool1']
    assert len(BLOB_NAMES) == 22 and BLOB_NAMES[-1] == 'pool5'
    assert [n for n, k, _ in TOPOLOGY if k == 'pool'] == ['pool%d' % i for i in range(1, 6)]


def test_ceil_mode_extents():
    # cv2.dnn on the reference prototxt gives 75x101 -> 38x51 -> 19x26 -> 10x13 -> 5x7 -> 3x4
    hw = (75, 101)
    seen = []
    for _ in range(5):
        hw = (pool_out(hw[0]), pool_out(hw[1]))
        seen.append(hw)
    assert seen == [(38, 51), (19, 26), (10, 13), (5, 7), (3, 4)]
    m = CaffeCPUModel()
    out = m.forward(np.zeros((1, 3, 75, 101), np.float32))
    assert out['pool5'].shape == (1, 512, 3, 4)
    assert out['conv4_2'].shape == (1, 512, 10, 13)


def _autograd_reference(params, x, diffs):
    """Gradient of sum_l <diff_l, node_l> where node_l is the PRE-ReLU conv output for convX_Y
    blobs (Caffe starts backward at the Convolution layer), the pooled output for poolN, x for data."""
    xt = torch.from_numpy(x).double().requires_grad_(True)
    cur = xt
    total = 0
    if 'data' in diffs:
        total = total + (cur * torch.from_numpy(diffs['data']).double()).sum()
    for name, kind, _ in TOPOLOGY[1:]:
        if kind == 'conv':
            w, b = params[name]
            pre = F.conv2d(cur, torch.from_numpy(w).double(), torch.from_numpy(b).double(), padding=1)
            if name in diffs:
                total = total + (pre * torch.from_numpy(diffs[name]).double()).sum()
            cur = F.relu(pre)
        else:
            cur = F.max_pool2d(cur, 2, 2, ceil_mode=True)
            if name in diffs:
                total = total + (cur * torch.from_numpy(diffs[name]).double()).sum()
    total.backward()
    return xt.grad.numpy()


@pytest.mark.parametrize('layers', [['conv2_1'], ['conv4_2', 'conv1_1', 'conv3_1'],
                                    ['pool2', 'conv3_2', 'data'], ['pool5', 'conv5_4'], ['data']])
def test_backward_matches_autograd_with_unmasked_injection(layers):
    params = synthetic_weights(0)
    m = CaffeCPUModel(params, dtype=torch.float64)
    rs = np.random.RandomState(7)
    x = (rs.rand(1, 3, 21, 27) * 255 - 120).astype(np.float32)
    feats = m.forward(x)
    diffs = {l: rs.randn(*feats[l].shape).astype(np.float32) for l in layers}
    got = m.backward(diffs)
    want = _autograd_reference(params, x, diffs)
    err = np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30)
    assert err < 1e-6, err


def test_injected_diff_is_not_relu_masked():
    """A diff injected at a conv blob where the activation is dead must still propagate."""
    params = synthetic_weights(0)
    m = CaffeCPUModel(params, dtype=torch.float64)
    x = np.zeros((1, 3, 8, 8), np.float32)
    feats = m.forward(x)
    dead = feats['conv1_1'] <= 0
    assert dead.any()
    d = np.zeros_like(feats['conv1_1'])
    d[dead] = 1.0
    assert np.abs(m.backward({'conv1_1': d})).max() > 0


def test_first_max_wins_on_ties():
    m = CaffeCPUModel()
    g = torch.zeros(1, 1, 1, 1) + 1.0
    cur = torch.ones(1, 1, 2, 2)
    _, idx = F.max_pool2d(cur, 2, 2, ceil_mode=True, return_indices=True)
    assert int(idx.reshape(-1)[0]) == 0
    assert F.max_unpool2d(g, idx, 2, 2, output_size=(2, 2)).reshape(-1).tolist() == [1, 0, 0, 0]
