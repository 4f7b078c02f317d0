"""The steady-state L-BFGS iteration replayed from CUDA graphs (StyleTransfer._graph_step) against the same job
enqueued kernel by kernel: same iterates, same traces, and every state change drops the graphs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _job(use_graphs, size=96):
    import bench
    st, _ = bench.build_job(size, 'fp32', prefill=0)
    st.use_graphs = use_graphs
    return st


def _run(st, n):
    out = []
    for _ in range(n):
        img, tr = st.step()
        out.append((img, dict(tr)))
    return out


def _same(a, b, tol=2e-4):
    for (ia, ta), (ib, tb) in zip(a, b):
        assert list(ta) == list(tb)
        for k in ta:
            if k == 'time':
                continue
            assert ta[k] == pytest.approx(tb[k], rel=tol, abs=1e-12), k
        assert np.abs(ia - ib).max() <= tol * 255


def test_graph_replay_follows_the_eager_trajectory():
    eager, graphed = _job(False), _job(True)
    a, b = _run(eager, 14), _run(graphed, 14)
    assert eager._graphs is None
    assert graphed._graphs is not None and graphed.engine.graph_launches > 0
    assert [t['fevals'] for _, t in b] == list(range(1, 15))
    _same(a, b)
    # the L-BFGS history the graphs maintained is the eager one
    sa, sb = eager.optimizer.syk, graphed.optimizer.syk
    assert len(sa) == len(sb) == 10
    np.testing.assert_allclose(sa, sb, rtol=1e-3)


def test_state_changes_drop_the_graphs_and_the_job_goes_on():
    import bench
    eager, graphed = _job(False), _job(True)
    _run(eager, 8), _run(graphed, 8)
    assert graphed._graphs is not None
    weights = {k: dict(v) for k, v in bench.WEIGHTS.items()}
    weights['style'] = {k: 2.0 * v for k, v in weights['style'].items()}
    for st in (eager, graphed):
        st.set_weights(weights, bench.PARAMS)
    assert graphed._graphs is None
    a, b = _run(eager, 8), _run(graphed, 8)
    assert graphed._graphs is not None            # captured again once the new objective was in steady state
    _same(a, b)
    for st in (eager, graphed):
        st.resample_input((64, 80))
        st.resample_content((64, 80))
    assert graphed._graphs is None
    _same(_run(eager, 6), _run(graphed, 6))


def test_pipelined_steps_and_direct_evaluations_mix_with_replay():
    eager, graphed = _job(False), _job(True)
    outs = []
    for st in (eager, graphed):
        hs = []
        for i in range(10):
            h = st.step_async()
            if i == 6:
                st.opfunc(st.input, return_grad=False)        # e.g. a GetImages-time evaluation between two steps
            hs.append((np.array(h.result()[0]), dict(h.result()[1])))
        outs.append(hs)
    assert graphed._graphs is not None
    _same(*outs)
