"""CPU oracle for the style_transfer2 worker hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``style_transfer2_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker or the timed CPU arm -- never as a product path.

What it is: a plain NumPy / torch-CPU restatement of the reference's worker
algorithm (``/root/reference/worker.py``, ``optimizers.py``, ``utils.py``),
function by function, each citing the reference ``file:line`` it follows.  The
one third-party piece whose arithmetic is *not* in ``/root/reference`` is BVLC
Caffe (un-pinned: ``config.ini:7`` just points at a checkout); its layer
semantics (cross-correlation 3x3 pad 1 with bias, in-place ReLU, ceil-mode 2x2
max-pool with first-max arg-max, ``net.backward(start=, end=)`` on layer names)
are restated in ``oracle/caffe_cpu.py`` from the published algorithm.

Parity pin status
-----------------
* objective / optimizers / numeric utils: PINNED -- ``oracle/make_golden.py``
  imports the reference's own Python from ``/root/reference`` *unmodified* and
  stores its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
  checks this restatement against those vectors.
* Caffe layer semantics: the reference ships no test, fixture or golden vector
  for them and Caffe itself is not installable offline, so that part is
  "parity unpinned" by the reference; it is cross-checked against OpenCV's
  independent Caffe importer (``cv2.dnn.readNetFromCaffe`` on the reference
  prototxt, forward) and torch autograd / finite differences (backward).
"""
