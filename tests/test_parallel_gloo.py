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
            assert all(s % 16 == 0 for s, e in b if e > s)
            assert max(e for _, e in b) == h


def test_single_process_fallbacks():
    assert parallel.all_max(3.5) == 3.5
    assert parallel.all_sum(2) == 2.0
    assert parallel.gather_objects({'a': 1}) == [{'a': 1}]
