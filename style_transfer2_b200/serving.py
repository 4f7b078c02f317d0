"""Job-parallel serving (BASELINE config 5): many independent style-transfer jobs,
sharded over the GPUs of a node (one process per GPU), several resident per GPU.

The reference serves one job per ``app.py`` + ``worker.py`` pair and a router hands each browser session
to a free pair (router.py:72-84); a job is defined entirely by the message sequence the app sends its
worker (app.py:244-262, 177-228): ``SetImages`` + ``SetWeights`` (+ ``SetOptimizer``) +
``StartIteration``.  Here the same message sequences are the unit of work: ``JobScheduler`` feeds each one
through the worker's own ``process_message`` dispatch into a private ``StyleTransfer`` and then steps the
resident jobs round-robin (one iteration per job per turn, so latency is shared fairly), all on the
device, fetching an iterate only when a job reports or finishes.  Jobs share nothing but the packed
weights, so there is no data-path collective; ``parallel.shard_jobs`` assigns job j to rank j % world.
"""
import time

import numpy as np
import torch

from . import parallel
from .messages import Iterate, SetImages, SetOptimizer, SetWeights, StartIteration
from .worker import StyleTransfer, Worker


def job_messages(size, content, style, weights, params, optimizer='lbfgs', step_size=None, seed=0):
    """The message sequence app.py sends for a fresh job (app.py:244-262): random initial image
    (app.py:251), content, style, weights, optimizer, start."""
    h, w = (size, size) if np.isscalar(size) else size
    x0 = np.uint8(np.random.RandomState(seed).uniform(0, 255, (h, w, 3)))
    msgs = [SetImages((h, w), x0, content, style, reset_state=True), SetWeights(weights, params)]
    if optimizer != 'lbfgs' or step_size is not None:
        msgs.append(SetOptimizer(optimizer, SetOptimizer.step_sizes[optimizer] if step_size is None else step_size))
    msgs.append(StartIteration())
    return msgs


class _Resident:
    def __init__(self, index, transfer):
        self.index, self.transfer = index, transfer
        self.started = time.perf_counter()
        self.steps = 0


class JobScheduler:
    """Runs this rank's share of ``jobs`` (lists of messages) for ``steps`` iterations each."""

    def __init__(self, model, max_resident=8):
        self.model = model
        self.max_resident = max_resident
        self._dispatch = Worker.process_message          # the reference-compatible dispatch, no sockets needed

    def _admit(self, index, msgs):
        tr = StyleTransfer(self.model, private_plans=True)
        shim = type('Shim', (), {})()
        shim.transfer = tr
        shim.sock_out = type('Null', (), {'send_pyobj': staticmethod(lambda obj: None)})()
        for m in msgs:
            self._dispatch(shim, m)
        if not tr.is_running:
            raise RuntimeError('job %d did not start (inconsistent images?)' % index)
        return _Resident(index, tr)

    def run(self, jobs, steps, world_size=1, rank=0, fetch_final=True):
        """Returns {job index: {'iterate': Iterate | None, 'latency_s': float, 'steps': int}}."""
        mine = parallel.shard_jobs(len(jobs), world_size, rank)
        pending = list(mine)
        resident, done = [], {}
        dev = self.model.engine.device
        while pending or resident:
            while pending and len(resident) < self.max_resident:
                j = pending.pop(0)
                resident.append(self._admit(j, jobs[j]))
            for r in list(resident):
                r.transfer.step(fetch=False)
                r.steps += 1
                if r.steps >= steps:
                    it = None
                    if fetch_final:
                        image = np.array(r.transfer.image())
                        it = Iterate(image, r.transfer.t, dict(r.transfer.traces[-1].data))
                    else:
                        torch.cuda.current_stream(dev).synchronize()
                    done[r.index] = {'iterate': it, 'latency_s': time.perf_counter() - r.started, 'steps': r.steps}
                    r.transfer.close()
                    resident.remove(r)
        return done
