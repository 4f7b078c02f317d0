"""world_size-2 gloo tests (CPU) of the multi-process host logic used by bench.py --gpus N and the
job-parallel serving path: job sharding, max-over-ranks timing, result gathering."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from style_transfer2_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        jobs = ['job%d' % i for i in range(7)]
        mine = parallel.run_jobs(jobs, lambda j, name: (rank, name.upper()), world, rank)
        assert sorted(mine) == parallel.shard_jobs(7, world, rank)
        ms = parallel.all_max(10.0 + rank)                 # slowest rank decides
        total = parallel.all_sum(len(mine))
        everything = parallel.gather_objects(mine)
        dist.barrier()
        with open(os.path.join(out_dir, 'r%d.txt' % rank), 'w') as f:
            merged = {}
            for part in everything:
                merged.update(part)
            f.write('%r|%r|%r' % (ms, total, sorted(merged.items())))
    finally:
        dist.destroy_process_group()


def test_job_sharding_and_reductions_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    texts = [open(tmp_path / ('r%d.txt' % r)).read() for r in range(2)]
    assert texts[0] == texts[1]
    ms, total, merged = texts[0].split('|')
    assert float(ms) == 11.0 and float(total) == 7.0
    items = eval(merged)
    assert [k for k, _ in items] == list(range(7))
    assert all(v == (k % 2, 'JOB%d' % k) for k, v in items)


def test_shard_jobs_partition():
    for n in (0, 1, 5, 64):
        for world in (1, 2, 3, 8):
            seen = sorted(j for r in range(world) for j in parallel.shard_jobs(n, world, r))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        parallel.shard_jobs(4, 2, 2)


def test_strip_bounds_cover_canvas_and_align():
    for h in (4096, 3071, 1024, 100, 16):
        for world in (1, 2, 4, 8):
            b = parallel.strip_bounds(h, world)
            assert b[0][0] == 0 and len(b) == world
            assert all(s1 == e0 for (_, e0), (s1, _) in zip(b, b[1:]))
            assert all(s % 32 == 0 for s, e in b if e > s)
            assert max(e for _, e in b) == h


def test_strip_bounds_keep_all_five_pools_strip_local():
    """pool1..pool5 are 2x2/2 ceil-mode pools computed per strip: the strips' pooled row counts must add up to the
    whole canvas' at every level (a boundary at an odd multiple of 16 rows would split a pool5 window: 1080 rows
    over 8 strips gave 36 pool5 rows instead of 34)."""
    from oracle.caffe_cpu import pool_out
    for h in (4096, 3071, 1080, 1024, 150, 100):
        for world in (1, 2, 4, 8):
            b = [(s, e) for s, e in parallel.strip_bounds(h, world) if e > s]
            whole, parts = h, [e - s for s, e in b]
            for level in range(5):
                whole = pool_out(whole)
                parts = [pool_out(n) for n in parts]
                assert sum(parts) == whole, (h, world, level)


def test_single_process_fallbacks():
    assert parallel.all_max(3.5) == 3.5
    assert parallel.all_sum(2) == 2.0
    assert parallel.gather_objects({'a': 1}) == [{'a': 1}]


def _strip_worker(rank, world, port, out_dir):
    """The reduction protocol of the row-strip tiling (tiled.py / DESIGN section 6) on the CPU: every rank holds a
    strip of rows, swaps ONE circular halo row of x with its neighbours, computes its partial sums with the oracle's
    own formulas and all-reduces them -- the totals must be the whole canvas' values."""
    import numpy as np
    from oracle import numeric as nm
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(5)
        H, W, C = 48, 20, 6
        x = rs.randn(1, 3, H, W) * 40
        F = np.maximum(rs.randn(1, C, H, W), 0)            # a post-ReLU "feature map" and its content target
        Fc = np.maximum(rs.randn(1, C, H, W), 0)
        r0, r1 = parallel.strip_bounds(H, world)[rank]
        up, dn = (rank - 1) % world, (rank + 1) % world    # circular: the TV term wraps around the canvas
        mine = torch.from_numpy(np.ascontiguousarray(x[:, :, r0:r1]))
        halo_dn, halo_up = torch.empty(1, 3, 1, W, dtype=mine.dtype), torch.empty(1, 3, 1, W, dtype=mine.dtype)
        ops = [dist.P2POp(dist.isend, mine[:, :, :1].contiguous(), up), dist.P2POp(dist.isend, mine[:, :, -1:].contiguous(), dn),
               dist.P2POp(dist.irecv, halo_dn, dn), dist.P2POp(dist.irecv, halo_up, up)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        # TV partial: forward differences need the row BELOW the strip's last row (circular)
        xs = np.concatenate([mine.numpy(), halo_dn.numpy()], axis=2) / 255.0
        dw = xs[:, :, :-1] - np.roll(xs[:, :, :-1], -1, axis=3)
        dh = xs[:, :, :-1] - xs[:, :, 1:]
        tv_part = np.sum(dw ** 2 + dh ** 2 + 1e-8)          # beta = 2
        Fs = F[0, :, r0:r1].reshape(C, -1)
        sums = torch.tensor([tv_part, np.sum((F[:, :, r0:r1] - Fc[:, :, r0:r1]) ** 2), np.sum(F[:, :, r0:r1] ** 2)],
                            dtype=torch.float64)
        gram = torch.from_numpy(Fs @ Fs.T)                  # un-normalised strip Gram sum
        dist.all_reduce(sums)
        dist.all_reduce(gram)
        want_tv, _ = nm.total_variation(x / 255.0, 2)
        Fw = F[0].reshape(C, -1)
        assert np.isclose(sums[0].item(), want_tv, rtol=1e-12)
        assert np.isclose(sums[1].item(), np.sum((F - Fc) ** 2), rtol=1e-12)
        assert np.isclose(sums[2].item(), np.sum(F ** 2), rtol=1e-12)
        np.testing.assert_allclose(gram.numpy() / (C * H * W), Fw @ Fw.T / (C * H * W), rtol=1e-12)
        with open(os.path.join(out_dir, 's%d.txt' % rank), 'w') as f:
            f.write('ok %d %d' % (r0, r1))
    finally:
        dist.destroy_process_group()


def test_strip_reductions_reproduce_the_whole_canvas_world2(tmp_path):
    port = _free_port()
    mp.spawn(_strip_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / ('s%d.txt' % r)).read() for r in range(2)] == ['ok 0 32', 'ok 32 48']


class _FakeTransfer:
    """Records what the worker loop drives (the real ones need a GPU)."""

    def __init__(self):
        self.is_running = False
        self.log = []
        self.t = 0

    def check_consistency(self):
        return True

    def start(self):
        self.is_running = True
        self.log.append('start')
        return True

    def pause(self):
        self.is_running = False
        self.log.append('pause')

    def set_weights(self, weights, params):
        self.log.append(('weights', sorted(weights)))

    def set_optimizer_class(self, cls, step_size):
        self.log.append(('optimizer', cls.__name__, step_size))

    def step_async(self):
        self.t += 1
        self.log.append(('step', self.t))
        t = self.t

        class Handle:
            def result(self_inner):
                import numpy as np
                return (np.zeros((2, 2, 3), np.float32) if self.rank0 else None), {'loss': 1.0, 'fevals': t}
        h = Handle()
        h.t = t
        return h


def _worker_loop_rank(rank, world, port, zport_in, zport_out, out_dir):
    """Rank 0 owns the ZeroMQ sockets and broadcasts what arrives; every rank replays the same messages and takes the
    same steps (style_transfer2_b200.worker.Worker with ``gpus = ...``), here with a recording stand-in transfer."""
    import pickle
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from style_transfer2_b200 import messages as m
        from style_transfer2_b200 import worker as wk
        m.install_as_toplevel()
        w = wk.Worker.__new__(wk.Worker)
        w.rank, w.world, w._ctl = rank, world, dist.new_group(backend='gloo')
        w.run_should_stop = False
        w.pickle_protocol = pickle.DEFAULT_PROTOCOL
        w.transfer = _FakeTransfer()
        w.transfer.rank0 = rank == 0
        if rank == 0:
            import zmq
            w._zmq = zmq
            w.ctx = zmq.Context.instance()
            w.sock_in = w.ctx.socket(zmq.PULL)
            w.sock_out = w.ctx.socket(zmq.PUSH)
            w.sock_in.bind('tcp://127.0.0.1:%d' % zport_in)
            w.sock_out.connect('tcp://127.0.0.1:%d' % zport_out)
        else:
            w._zmq, w.ctx, w.sock_in, w.sock_out = None, None, None, wk._NullSocket()
        w.run()
        with open(os.path.join(out_dir, 'w%d.txt' % rank), 'w') as f:
            f.write(repr(w.transfer.log))
    finally:
        dist.destroy_process_group()


def test_worker_replays_messages_on_every_rank_world2(tmp_path):
    zmq = pytest.importorskip('zmq')
    from style_transfer2_b200 import messages as m
    m.install_as_toplevel()
    port, zin, zout = _free_port(), _free_port(), _free_port()
    ctx = zmq.Context.instance()
    app_in = ctx.socket(zmq.PULL)
    app_in.bind('tcp://127.0.0.1:%d' % zout)
    app_in.RCVTIMEO = 60000
    app_out = ctx.socket(zmq.PUSH)
    app_out.connect('tcp://127.0.0.1:%d' % zin)
    procs = mp.spawn(_worker_loop_rank, args=(2, port, zin, zout, str(tmp_path)), nprocs=2, join=False)
    try:
        app_out.send_pyobj(m.SetWeights({'content': {}, 'style': {}, 'deepdream': {}}, {}))
        app_out.send_pyobj(m.SetOptimizer('adam'))
        app_out.send_pyobj(m.StartIteration())
        seen = []
        while len(seen) < 3:
            it = app_in.recv_pyobj()
            assert isinstance(it, m.Iterate)
            seen.append(it.i)
        assert seen == [1, 2, 3]
        app_out.send_pyobj(m.PauseIteration())
        app_out.send_pyobj(m.Shutdown())
        while not isinstance(app_in.recv_pyobj(), m.Shutdown):
            pass
        procs.join(60)
    finally:
        app_in.close(0)
        app_out.close(0)
    logs = [eval(open(tmp_path / ('w%d.txt' % r)).read()) for r in range(2)]
    assert logs[0] == logs[1]                                   # same messages, same steps, in the same order
    assert logs[0][:3] == [('weights', ['content', 'deepdream', 'style']), ('optimizer', 'AdamOptimizer', 10), 'start']
    assert ('step', 3) in logs[0] and logs[0][-1] == 'pause'
