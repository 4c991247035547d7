"""Node view, fused Op, local service registry and the process launcher (CPU / gloo)."""
import numpy as np
import pytest

from pytensor_federated_b200 import LogpGradServiceClient
from pytensor_federated_b200._graph_backend import at, function, grad
from pytensor_federated_b200.federation import NodeFederation, launch_federation
from pytensor_federated_b200.models import LinregShards, make_demo_data
from pytensor_federated_b200.parallel import FederatedEngine

pytestmark = pytest.mark.timeout(300)


def _three_node_engine():
    x, y, sigma = make_demo_data()
    return FederatedEngine(LinregShards([x, x, x], [y, y + 1.0, y - 0.5], [sigma] * 3), backend="collective")


def test_node_view_matches_per_node_models():
    eng = _three_node_engine()
    fed = NodeFederation(eng)
    x, y, sigma = make_demo_data()
    single = FederatedEngine(LinregShards([x], [y + 1.0], [sigma]), backend="collective")
    want = single.logp_grad(np.array(0.3), np.array(0.6))
    got = fed.evaluate_node(1, np.array(0.3), np.array(0.6))
    np.testing.assert_allclose(got[0], want[0], rtol=1e-12)
    np.testing.assert_allclose(got[1], want[1], rtol=1e-12)
    res = fed.evaluate_nodes({0: (0.1, 0.5), 2: (0.2, 0.4)})
    assert set(res) == {0, 2} and fed.n_launches == 2


def test_fused_federated_ops_use_one_launch_per_model_evaluation():
    """demo_model.py topology: 3 remote calls with offset intercepts -> ONE engine launch."""
    fed = NodeFederation(_three_node_engine())
    ops = fed.node_ops()
    icpt = at.vector("intercept")
    slope = at.scalar("slope")
    total = None
    for i, (op, off) in enumerate(zip(ops, np.linspace(-1.5, 1.5, 3))):
        logp, *_ = op(icpt[i] + off, slope)
        total = logp if total is None else total + logp
    fn = function([icpt, slope], [total, *grad(total, [icpt, slope])])
    kinds = [type(n.op).__name__ for n in fn.maker.fgraph.toposort()]
    assert kinds.count("ParallelAsyncOp") == 1 and "FederatedLogpGradOp" not in kinds
    fed.n_launches = 0
    val, g_ic, g_slope = fn(np.array([0.1, 0.2, 0.3]), 0.5)
    assert fed.n_launches == 1
    # oracle: per-node evaluation
    want, want_ic, want_slope = 0.0, [], 0.0
    for i, off in enumerate(np.linspace(-1.5, 1.5, 3)):
        lp, (da, db) = fed.evaluate_node(i, [0.1, 0.2, 0.3][i] + off, 0.5)
        want += lp
        want_ic.append(da)
        want_slope += db
    np.testing.assert_allclose(val, want, rtol=1e-12)
    np.testing.assert_allclose(g_ic, want_ic, rtol=1e-12)
    np.testing.assert_allclose(g_slope, want_slope, rtol=1e-12)
    # unfused compile mode: one launch per node
    slow = function([icpt, slope], total, mode="FAST_COMPILE")
    fed.n_launches = 0
    slow(np.array([0.1, 0.2, 0.3]), 0.5)
    assert fed.n_launches == 3


def test_nodes_are_reachable_through_the_reference_client_api():
    fed = NodeFederation(_three_node_engine())
    addresses = fed.register_services("gpu", 10)
    try:
        client = LogpGradServiceClient(hosts_and_ports=[addresses[1]])
        logp, grads = client.evaluate(np.array(0.3), np.array(0.6))
        want = fed.evaluate_node(1, 0.3, 0.6)
        np.testing.assert_allclose(logp, want[0])
        np.testing.assert_allclose(grads, want[1])
        del client
    finally:
        fed.shutdown()


def _build(rank, world, dev):
    rng = np.random.default_rng(40 + rank)
    x = rng.normal(size=64)
    y = 2.0 - 0.3 * x + rng.normal(scale=0.5, size=64)
    return LinregShards([x], [y], [0.5], local_ids=[rank], n_shards_total=world, device=dev)


def test_launch_federation_spawns_peer_nodes():
    import scipy.stats

    with launch_federation(_build, 3, device_type="cpu", backend="collective") as eng:
        assert eng.world == 3 and eng.is_root
        logp, da, db = eng.evaluate(np.array(2.0), np.array(-0.3))
        fed = NodeFederation(eng)
        per = fed.evaluate_nodes({r: (2.0, -0.3) for r in range(3)})
    want = 0.0
    for rank in range(3):
        rng = np.random.default_rng(40 + rank)
        x = rng.normal(size=64)
        y = 2.0 - 0.3 * x + rng.normal(scale=0.5, size=64)
        lp = scipy.stats.norm.logpdf(y, 2.0 - 0.3 * x, 0.5).sum()
        np.testing.assert_allclose(per[rank][0], lp, rtol=1e-12)
        want += lp
    np.testing.assert_allclose(logp, want, rtol=1e-12)
