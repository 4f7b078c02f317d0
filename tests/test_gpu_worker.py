"""Wire-protocol conformance of the drop-in worker on a GPU: a scripted stand-in for app.py talks to
``Worker`` over real ZeroMQ PUSH/PULL sockets with pickled ``messages.*`` (worker.py:318-409,
app.py:244-262,293-323 for the message sequences)."""
import socket
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('extra', [{}, {'tiles': '2'}], ids=['one-plan', 'two-strips'])
def test_worker_speaks_the_reference_protocol(golden, extra):
    """The same scripted app against the whole-canvas worker and against a worker whose canvas is split into two row
    strips (config key ``tiles``: the tiling scheduler behind the unchanged protocol, incl. the optimizer swap and
    the RESAMPLE scale change on sharded state)."""
    zmq = pytest.importorskip('zmq')
    from style_transfer2_b200 import messages as m
    from style_transfer2_b200 import vgg
    from style_transfer2_b200.worker import Worker
    m.install_as_toplevel()
    g = golden('small')
    cfg = {'worker_socket': 'tcp://127.0.0.1:%d' % _port(), 'app_socket': 'tcp://127.0.0.1:%d' % _port(),
           'gpu': '0', 'precision': 'fp16'}
    cfg.update(extra)
    ctx = zmq.Context.instance()
    app_in = ctx.socket(zmq.PULL)
    app_in.bind(cfg['app_socket'])
    app_out = ctx.socket(zmq.PUSH)
    app_out.connect(cfg['worker_socket'])
    app_in.RCVTIMEO = 120000
    errors = []

    def serve():
        try:
            Worker(cfg).run()
        except Exception:               # pragma: no cover
            import traceback
            errors.append(traceback.format_exc())

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    try:
        ready = app_in.recv_pyobj()
        assert isinstance(ready, m.WorkerReady) and ready.layers == vgg.BLOBS

        app_out.send_pyobj(m.StartIteration())                       # nothing set yet
        assert isinstance(app_in.recv_pyobj(), m.GetImages)

        weights = {'content': {'conv4_2': 0.08}, 'style': {'conv1_1': 1, 'conv2_1': 1, 'conv3_1': 1, 'conv4_1': 1},
                   'deepdream': {}}
        params = {'p': 50, 'p_power': 6, 'tv': 5, 'tv_power': 2}
        app_out.send_pyobj(m.SetWeights(weights, params))
        app_out.send_pyobj(m.SetOptimizer('lbfgs'))
        app_out.send_pyobj(m.SetImages(size=g['x0'].shape[:2], input_image=g['x0'], content_image=g['content'],
                                       style_image=g['style'], reset_state=True))
        app_out.send_pyobj(m.StartIteration())
        seen = []
        for _ in range(3):
            it = app_in.recv_pyobj()
            assert isinstance(it, m.Iterate), (it, errors)
            img = np.float32(it.image)
            assert img.shape == g['x0'].shape and np.isfinite(img).all()
            assert 'loss' in it.trace and it.trace['fevals'] == it.i
            assert all(isinstance(v, (int, float)) for v in it.trace.values())
            seen.append(it.i)
        assert seen == [1, 2, 3]

        app_out.send_pyobj(object())                                   # unknown message: logged, ignored
        app_out.send_pyobj(m.PauseIteration())
        app_out.send_pyobj(m.SetOptimizer('adam'))                     # class change -> reset, i restarts
        h, w = g['x0'].shape[:2]
        app_out.send_pyobj(m.SetImages(size=(h + 8, w + 8), input_image=m.SetImages.RESAMPLE,
                                       content_image=m.SetImages.RESAMPLE))
        app_out.send_pyobj(m.StartIteration())
        while True:
            it = app_in.recv_pyobj()
            if isinstance(it, m.Iterate) and np.float32(it.image).shape == (h + 8, w + 8, 3):
                break
        assert it.i >= 1 and np.isfinite(it.trace['loss'])

        app_out.send_pyobj(m.Shutdown())
        while True:
            last = app_in.recv_pyobj()
            if isinstance(last, m.Shutdown):
                break
        th.join(30)
        assert not th.is_alive() and not errors
    finally:
        app_in.close(0)
        app_out.close(0)


def test_job_scheduler_runs_interleaved_jobs_like_standalone_ones(golden):
    """BASELINE config 5 in miniature: five jobs (different seeds, two canvas sizes, one on Adam) resident on
    one GPU and stepped round-robin give what each job gives when run alone through StyleTransfer."""
    import ast
    from style_transfer2_b200 import optimizers, serving
    from style_transfer2_b200.model import B200Model
    from style_transfer2_b200.worker import StyleTransfer
    g = golden('small')
    weights, params = ast.literal_eval(str(g['weights_repr'])), ast.literal_eval(str(g['params_repr']))
    model = B200Model(precision='fp32')
    h, w = g['x0'].shape[:2]
    jobs, specs = [], []
    for j in range(5):
        size = (h, w) if j % 2 == 0 else (h - 16, w - 16)
        content = g['content'][:size[0], :size[1]]
        opt = 'adam' if j == 3 else 'lbfgs'
        jobs.append(serving.job_messages(size, content, g['style'], weights, params, optimizer=opt, seed=j))
        specs.append((size, content, opt))
    out = serving.JobScheduler(model, max_resident=3).run(jobs, steps=3)
    assert sorted(out) == list(range(5))
    for j, (size, content, opt) in enumerate(specs):
        st = StyleTransfer(model)
        if opt == 'adam':
            st.optimizer_cls, st.step_size = optimizers.AdamOptimizer, 10
        st.set_input(np.uint8(np.random.RandomState(j).uniform(0, 255, size + (3,))))
        st.set_content(content)
        st.set_style(g['style'])
        st.set_weights(weights, params)
        assert st.start()
        for _ in range(3):
            img, tr = st.step()
        it = out[j]['iterate']
        assert it.i == 3 and out[j]['steps'] == 3 and out[j]['latency_s'] > 0
        assert np.abs(np.float32(it.image) - img).max() < 1e-2 * max(1.0, np.abs(img).max())
        assert np.isclose(it.trace['loss'], tr['loss'], rtol=1e-5)
