#!/usr/bin/env python3
"""Iterations/sec of the WHOLE drop-in worker as app.py sees it (SURVEY 8d "second figure"): a scripted app talks
to ``Worker`` over real ZeroMQ PUSH/PULL sockets; every iterate is deprocessed, copied to the host, pickled
(H x W x 3 fp32 = 12.6 MB at 1024^2), sent, received and unpickled (worker.py:351-353, app.py:293-323).

    python tools/worker_bench.py [--size 1024] [--iterates 100] [--transport ipc|tcp] [--tiles P | --gpus 0,1,..]

``--gpus 0,1,...`` starts the worker the way app.py does -- ``python -m style_transfer2_b200.worker <config>`` as a
child process -- with a config file carrying ``gpus = ...``: the worker re-launches itself as one process per GPU and
serves ONE canvas in row strips (rank 0 owns the sockets).  ``--tiles P`` keeps one process and splits the canvas
into P strips on one GPU.
"""
import argparse
import json
import os
import socket
import sys
import tempfile
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zmq

import bench
from style_transfer2_b200 import messages as m
from style_transfer2_b200.worker import Worker


def _port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=1024)
    ap.add_argument('--iterates', type=int, default=100)
    ap.add_argument('--transport', default='ipc', choices=['ipc', 'tcp'])
    ap.add_argument('--tiles', type=int, default=1)
    ap.add_argument('--gpus', default='')
    args = ap.parse_args()
    m.install_as_toplevel()
    if args.transport == 'ipc':
        d = tempfile.mkdtemp()
        cfg = {'worker_socket': 'ipc://%s/worker' % d, 'app_socket': 'ipc://%s/app' % d}
    else:
        cfg = {'worker_socket': 'tcp://127.0.0.1:%d' % _port(), 'app_socket': 'tcp://127.0.0.1:%d' % _port()}
    cfg.update(gpu='0', precision='fp16')
    if args.tiles > 1:
        cfg['tiles'] = str(args.tiles)
    ctx = zmq.Context.instance()
    app_in = ctx.socket(zmq.PULL)
    app_in.bind(cfg['app_socket'])
    app_out = ctx.socket(zmq.PUSH)
    app_out.connect(cfg['worker_socket'])
    app_in.RCVTIMEO = 300000
    child = None
    if args.gpus:
        import subprocess
        cfg['gpus'] = args.gpus
        path = os.path.join(tempfile.mkdtemp(), 'worker.ini')
        with open(path, 'w') as f:
            f.write('[DEFAULT]\n' + ''.join('%s = %s\n' % kv for kv in cfg.items()))
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        child = subprocess.Popen([sys.executable, '-m', 'style_transfer2_b200.worker', path], cwd=root)
        th = threading.Thread(target=child.wait, daemon=True)
    else:
        th = threading.Thread(target=lambda: Worker(cfg).run(), daemon=True)
    th.start()
    assert isinstance(app_in.recv_pyobj(), m.WorkerReady)
    content, style, x0 = bench.load_images(args.size)
    app_out.send_pyobj(m.SetWeights(bench.WEIGHTS, bench.PARAMS))
    app_out.send_pyobj(m.SetImages(size=x0.shape[:2], input_image=x0, content_image=content, style_image=style,
                                   reset_state=True))
    app_out.send_pyobj(m.StartIteration())
    for _ in range(10):                                   # warm-up
        it = app_in.recv_pyobj()
    t0 = time.perf_counter()
    nbytes = 0
    for _ in range(args.iterates):
        it = app_in.recv_pyobj()
        nbytes += np.asarray(it.image).nbytes
    dt = time.perf_counter() - t0
    app_out.send_pyobj(m.Shutdown())
    th.join(30)
    print(json.dumps({'what': 'whole worker over ZeroMQ (%s), every iterate delivered to the app' % args.transport,
                      'placement': ('row strips over GPUs %s, one process per GPU' % args.gpus) if args.gpus else
                                   ('%d row strips in one process' % args.tiles if args.tiles > 1 else 'one plan'),
                      'canvas': [args.size, args.size], 'iterates': args.iterates, 'iterations_per_s': args.iterates / dt,
                      'ms_per_iterate': 1e3 * dt / args.iterates, 'iterate_bytes': nbytes // args.iterates,
                      'last_i': it.i, 'last_loss': it.trace['loss']}))


if __name__ == '__main__':
    main()
