"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol include/st2.h
declares, the wire messages stay pickle-compatible with the reference's names, the network
description matches the reference prototxt, the caffemodel reader round-trips."""
import ctypes
import os
import pickle
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'


def test_library_loads_and_exports_every_declared_symbol():
    from style_transfer2_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'st2.h')).read()
    declared = set(re.findall(r'\b(st2_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.st2_blob_count() == 22
    names = [lib.st2_blob_name(i).decode() for i in range(22)]
    from style_transfer2_b200 import vgg
    assert names == vgg.BLOBS
    assert [lib.st2_blob_channels(i) for i in range(22)] == [t[2] for t in vgg.TOPOLOGY]


def test_no_cpu_fallback_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from style_transfer2_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.st2_ctx_create(0, ctypes.byref(h)) == -1
    assert b'no CPU fallback' in lib.st2_last_error(None)
    from style_transfer2_b200.model import B200Model
    with pytest.raises(RuntimeError):
        B200Model()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'style_transfer2_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, re.M), f


def test_topology_matches_reference_prototxt():
    from style_transfer2_b200 import vgg
    path = os.path.join(REF, 'models', 'vgg19.prototxt')
    if not os.path.exists(path):
        pytest.skip('reference tree not present')
    text = open(path).read()
    assert vgg.check_prototxt(text)
    bad = text.replace('kernel_size: 3', 'kernel_size: 5', 1)
    with pytest.raises(ValueError):
        vgg.check_prototxt(bad)


def test_caffemodel_round_trip(tmp_path):
    from style_transfer2_b200 import vgg
    params = vgg.synthetic_weights(3)
    path = str(tmp_path / 'w.caffemodel')
    vgg.write_caffemodel(path, params)
    back = vgg.read_caffemodel(path)
    assert list(back) == vgg.CONVS
    for k in params:
        np.testing.assert_array_equal(back[k][0], params[k][0])
        np.testing.assert_array_equal(back[k][1], params[k][1])


def test_caffemodel_is_readable_by_opencv_and_matches_oracle(tmp_path):
    """Independent pin of the restated Caffe forward: OpenCV's Caffe importer on the REFERENCE
    prototxt + our synthetic caffemodel vs oracle/caffe_cpu.py."""
    cv2 = pytest.importorskip('cv2')
    proto = os.path.join(REF, 'models', 'vgg19.prototxt')
    if not os.path.exists(proto):
        pytest.skip('reference tree not present')
    from style_transfer2_b200 import vgg
    from oracle.caffe_cpu import CaffeCPUModel
    params = vgg.synthetic_weights(0)
    path = str(tmp_path / 'w.caffemodel')
    vgg.write_caffemodel(path, params)
    net = cv2.dnn.readNetFromCaffe(proto, path)
    rs = np.random.RandomState(0)
    x = (rs.rand(1, 3, 75, 101) * 255 - 120).astype(np.float32)
    net.setInput(x)
    names = ['relu1_1', 'pool1', 'relu3_1', 'relu4_2', 'pool5']
    outs = net.forward(names)
    want = CaffeCPUModel(params).forward(x)
    for n, o in zip(names, outs):
        blob = n.replace('relu', 'conv')
        assert o.shape == want[blob].shape
        err = np.linalg.norm(o - want[blob]) / np.linalg.norm(want[blob])
        assert err < 1e-5, (n, err)


def test_messages_pickle_under_the_reference_module_name():
    from style_transfer2_b200 import messages as m
    saved = sys.modules.get('messages')
    try:
        m.install_as_toplevel()
        msg = m.SetImages(size=(4, 5), input_image=np.zeros((4, 5, 3), np.uint8), reset_state=True)
        blob = pickle.dumps(msg)
        assert b'messages' in blob and b'style_transfer2_b200' not in blob
        back = pickle.loads(blob)
        assert isinstance(back, m.SetImages) and back.size == (4, 5) and back.reset_state
        it = pickle.loads(pickle.dumps(m.Iterate(np.ones((2, 2, 3), np.float32), 7, {'loss': 1.5})))
        assert it.i == 7 and it.trace == {'loss': 1.5}
        assert pickle.loads(pickle.dumps(m.WorkerReady(layers=['data']))).layers == ['data']
    finally:
        if saved is not None:
            sys.modules['messages'] = saved
        else:
            sys.modules.pop('messages', None)
    assert m.SetImages.RESAMPLE == 1
    assert m.SetOptimizer('adam').step_size == 10 and m.SetOptimizer('lbfgs').step_size == 1
    assert m.SetOptimizer('lbfgs', 3).step_size == 3
    with pytest.raises(ValueError):
        m.SetOptimizer('sgd')
    assert m.SetWeights.loss_names == ('content', 'style', 'deepdream')
    assert m.SetWeights.scalar_loss_names == ('tv', 'tv_power', 'p', 'p_power')
    assert m.WorkerReady().layers == []
    assert 'ndarray, shape: (4, 5, 3)' in repr(m.SetImages(input_image=np.zeros((4, 5, 3))))


def test_reference_pickles_are_readable_by_our_messages():
    """A pickle written by the reference's messages.py unpickles into our classes."""
    if not os.path.exists(os.path.join(REF, 'messages.py')):
        pytest.skip('reference tree not present')
    import subprocess
    code = ("import sys, pickle; sys.path.insert(0, %r); sys.dont_write_bytecode = True; sys.argv=['x']; "
            "import worker, messages; sys.stdout.buffer.write(pickle.dumps(["
            "messages.SetWeights({'style': {'conv1_1': 1}}, {'tv': 5, 'tv_power': 2, 'p': 50, 'p_power': 6}), "
            "messages.SetOptimizer('adam'), messages.StartIteration(), "
            "messages.SetImages(size=(3, 4), input_image=messages.SetImages.RESAMPLE)]))" % REF)
    blob = subprocess.run([sys.executable, '-c', code], capture_output=True, check=True).stdout
    from style_transfer2_b200 import messages as m
    saved = sys.modules.get('messages')
    try:
        m.install_as_toplevel()
        w, o, s, i = pickle.loads(blob)
    finally:
        if saved is not None:
            sys.modules['messages'] = saved
        else:
            sys.modules.pop('messages', None)
    assert isinstance(w, m.SetWeights) and w.params['p_power'] == 6
    assert isinstance(o, m.SetOptimizer) and o.optimizer == 'adam' and o.step_size == 10
    assert isinstance(s, m.StartIteration)
    assert isinstance(i, m.SetImages) and i.input_image == m.SetImages.RESAMPLE and i.size == (3, 4)


def test_sizes_match_reference(golden):
    from style_transfer2_b200 import utils
    g = golden('numeric')
    assert [tuple(r) for r in g['scales_300_200']] == utils.scales((300, 200), 32)
    assert tuple(g['fit'][0]) == utils.fit_into_square((979, 734), 256, True)
    assert tuple(g['fit'][2]) == utils.fit_into_square((100, 80), 256, False)


def test_read_config_accepts_the_stock_ini(tmp_path):
    from style_transfer2_b200 import utils
    ini = tmp_path / 'config.ini'
    ini.write_text('[DEFAULT]\ndebug = 0\napp_socket = tcp://127.0.0.1:23898\n'
                   'worker_socket = tcp://127.0.0.1:23899\ngpu = -1\nprototxt = models/vgg19.prototxt\n'
                   'caffemodel = models/vgg19.caffemodel\n')
    cfg = utils.read_config(None, [tmp_path])
    assert cfg['worker_socket'].endswith('23899') and cfg.getint('gpu') == -1
    assert cfg.get('precision', 'fp16') == 'fp16'          # new knobs are optional


def test_job_messages_follow_the_app_sequence_and_pickle():
    """serving.job_messages builds what app.py sends for a fresh job (app.py:244-262): SetImages with a seeded
    random uint8 input + content + style and reset_state, SetWeights, optional SetOptimizer, StartIteration --
    and every message survives the pickle round trip the ZeroMQ transport applies."""
    import pickle
    from style_transfer2_b200 import messages as m, serving
    content = np.zeros((32, 48, 3), np.uint8)
    style = np.ones((20, 30, 3), np.uint8)
    weights = {'content': {'conv4_2': 0.08}, 'style': {'conv1_1': 1.0}, 'deepdream': {}}
    params = {'tv': 5, 'tv_power': 2, 'p': 50, 'p_power': 6}
    msgs = serving.job_messages((32, 48), content, style, weights, params, seed=3)
    assert [type(x) for x in msgs] == [m.SetImages, m.SetWeights, m.StartIteration]
    si = msgs[0]
    assert si.reset_state and si.size == (32, 48) and si.input_image.shape == (32, 48, 3)
    assert si.input_image.dtype == np.uint8
    np.testing.assert_array_equal(si.input_image, np.uint8(np.random.RandomState(3).uniform(0, 255, (32, 48, 3))))
    adam = serving.job_messages(16, content[:16, :16], style, weights, params, optimizer='adam')
    assert [type(x) for x in adam] == [m.SetImages, m.SetWeights, m.SetOptimizer, m.StartIteration]
    assert adam[2].optimizer == 'adam' and adam[2].step_size == 10
    m.install_as_toplevel()                       # what the worker does before it opens its sockets
    for msg in msgs + adam:
        back = pickle.loads(pickle.dumps(msg))
        assert type(back) is type(msg) and sorted(vars(back)) == sorted(vars(msg))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs the CPU port alone (no GPU, no libst2) and prints ONE JSON line with the
    keys the driver reads: impl, metric, value, unit, config, cpu_baseline, e2e with zero copy bytes."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--size', '48',
                          '--steps', '1', '--warmup', '1', '--cpu-budget', '1'], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'style-transfer iterations/sec' and d['unit'] == 'it/s'
    assert d['value'] > 0 and d['higher_is_better'] is True and d['gpu_launches'] == 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config']


def _proto_current(blobs, tail=''):
    """A prototxt in the current dialect for a prefix of the VGG-19 stack (same shape as the reference's file)."""
    out = ['name: "cut"', 'layer { name: "data" type: "Input" top: "data" input_param { shape { dim: 1 dim: 3 dim: 8 dim: 8 } } }']
    prev = 'data'
    for name, kind, c in blobs[1:]:
        if kind == 'conv':
            out.append('layer { name: "%s" type: "Convolution" bottom: "%s" top: "%s" convolution_param { num_output: %d '
                       'pad: 1 kernel_size: 3 } }' % (name, prev, name, c))
            out.append('layer { name: "relu%s" type: "ReLU" bottom: "%s" top: "%s" }' % (name[4:], name, name))
        else:
            out.append('layer { name: "%s" type: "Pooling" bottom: "%s" top: "%s" pooling_param { pool: MAX kernel_size: 2 '
                       'stride: 2 } }' % (name, prev, name))
        prev = name
    return '\n'.join(out) + tail


def _proto_v1(blobs, tail=''):
    """The same in the legacy V1 dialect of the model zoo's VGG deploy files."""
    out = ['name: "VGG_ILSVRC_19_layers"', 'input: "data"', 'input_dim: 10', 'input_dim: 3', 'input_dim: 224', 'input_dim: 224']
    prev = 'data'
    for name, kind, c in blobs[1:]:
        if kind == 'conv':
            out.append('layers { bottom: "%s" top: "%s" name: "%s" type: CONVOLUTION convolution_param { num_output: %d '
                       'pad: 1 kernel_size: 3 } }' % (prev, name, name, c))
            out.append('layers { bottom: "%s" top: "%s" name: "relu%s" type: RELU }' % (name, name, name[4:]))
        else:
            out.append('layers { bottom: "%s" top: "%s" name: "%s" type: POOLING pooling_param { pool: MAX kernel_size: 2 '
                       'stride: 2 } }' % (prev, name, name))
        prev = name
    return '\n'.join(out) + tail


def test_prototxt_prefix_cuts_v1_dialect_and_classifier_tail():
    """(f)1: the prototxt is PARSED into the blob list the engine exposes: the reference's file, a VGG-19 cut earlier,
    the model zoo's V1 deploy dialect, and a classifier tail above pool5 that the style-transfer path never evaluates."""
    from style_transfer2_b200 import vgg
    full = vgg.TOPOLOGY
    blobs, ignored = vgg.net_from_prototxt(_proto_current(full))
    assert blobs == full and ignored == []
    cut = full[:vgg.BLOB_INDEX['conv4_2'] + 1]
    blobs, ignored = vgg.net_from_prototxt(_proto_current(cut))
    assert [b[0] for b in blobs][-1] == 'conv4_2' and len(blobs) == 14
    with pytest.raises(ValueError):
        vgg.check_prototxt(_proto_current(cut))                          # not the reference's exact net
    blobs, _ = vgg.net_from_prototxt(_proto_v1(full))
    assert blobs == full
    tail = ('\nlayers { bottom: "pool5" top: "fc6" name: "fc6" type: INNER_PRODUCT inner_product_param { num_output: 4096 } }'
            '\nlayers { bottom: "fc6" top: "fc6" name: "relu6" type: RELU }'
            '\nlayers { bottom: "fc6" top: "fc6" name: "drop6" type: DROPOUT }'
            '\nlayers { bottom: "fc6" top: "prob" name: "prob" type: SOFTMAX }')
    with pytest.raises(ValueError):
        vgg.net_from_prototxt(_proto_v1(full, tail), strict=True)
    blobs, ignored = vgg.net_from_prototxt(_proto_v1(full, tail), strict=False)
    assert blobs == full and ignored == ['fc6', 'relu6', 'drop6', 'prob']
    bad = _proto_current(full).replace('num_output: 256', 'num_output: 192', 1)
    with pytest.raises(ValueError):
        vgg.net_from_prototxt(bad)
    with pytest.raises(ValueError):
        vgg.net_from_prototxt(_proto_current(full).replace('kernel_size: 3', 'kernel_size: 5', 1))
