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


def test_fused_federated_ops_need_no_event_loop(monkeypatch):
    """The fused engine call blocks anyway, so a ParallelAsyncOp made only of federated Ops calls it directly
    (`perform_fused_sync`): no asyncio machinery per model evaluation; errors surface as on the async path."""
    from pytensor_federated_b200 import op_async

    fed = NodeFederation(_three_node_engine())
    ops = fed.node_ops()
    a, b = at.scalar("a"), at.scalar("b")
    total = ops[0](a, b)[0] + ops[1](a + 1.0, b)[0] + ops[2](a - 1.0, b)[0]
    fn = function([a, b], [total, *grad(total, [a, b])])

    def no_loop():
        raise AssertionError("the synchronous fast path must not touch the event loop")

    monkeypatch.setattr(op_async, "get_useful_event_loop", no_loop)
    fed.n_launches = 0
    val, da, db = fn(0.2, 0.4)
    assert fed.n_launches == 1
    want = sum(fed.evaluate_node(i, 0.2 + off, 0.4)[0] for i, off in enumerate((0.0, 1.0, -1.0)))
    np.testing.assert_allclose(val, want, rtol=1e-12)

    def boom(requests):
        raise RuntimeError("node 1 is on fire")

    monkeypatch.setattr(fed, "evaluate_nodes", boom)
    with pytest.raises(RuntimeError, match="on fire"):
        fn(0.2, 0.4)


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


def _glm_nodes(n_chains=1, n_nodes=3):
    import torch

    from pytensor_federated_b200.models import GlmShards

    torch.manual_seed(4)
    Xs = [torch.randn(150 + 20 * i, 8).to(torch.bfloat16) for i in range(n_nodes)]
    ys = [(torch.rand(X.shape[0]) < 0.45).float() for X in Xs]
    per_node = GlmShards(Xs, ys, groups=[0, 1, 0][:n_nodes], n_groups=2, n_chains=n_chains, node_ids=list(range(n_nodes)),
                         n_nodes=n_nodes, kernel="tc")
    return Xs, ys, per_node


def test_glm_nodes_keep_their_own_outputs_and_parameters():
    """One Op per GLM node (reference pattern) on a model that keeps per-node output blocks: shared parameters
    use one chain, distinct parameters one chain each, all in one launch."""
    from pytensor_federated_b200.models import GlmShards
    from pytensor_federated_b200.parallel import FederationError

    Xs, ys, model = _glm_nodes(n_chains=3)
    fed = NodeFederation(FederatedEngine(model, backend="collective"))
    assert fed.n_nodes == 3
    rng = np.random.default_rng(0)
    ic, beta = rng.normal(size=2) * 0.2, rng.normal(size=8) * 0.3
    # (1) every node at the same parameters == the node's own single-shard model; their sum == the pooled model
    res = fed.evaluate_nodes({i: (ic, beta) for i in range(3)})
    assert fed.n_launches == 1
    total = 0.0
    for i in range(3):
        single = FederatedEngine(GlmShards([Xs[i]], [ys[i]], groups=[[0, 1, 0][i]], n_groups=2, kernel="simt"), backend="collective")
        want = single.logp_grad(ic, beta)
        np.testing.assert_allclose(res[i][0], want[0], rtol=2e-6)
        np.testing.assert_allclose(res[i][1][0], want[1][0], rtol=2e-5, atol=1e-4)
        np.testing.assert_allclose(res[i][1][1], want[1][1], rtol=2e-5, atol=1e-4)
        total += float(want[0])
    pooled = FederatedEngine(GlmShards(Xs, ys, groups=[0, 1, 0], n_groups=2, kernel="simt"), backend="collective")
    np.testing.assert_allclose(pooled.evaluate(ic, beta)[0], total, rtol=2e-6)
    # (2) three different parameter vectors -> three chains, still one launch
    betas = [beta, beta * 0.5, -beta]
    res = fed.evaluate_nodes({i: (ic, betas[i]) for i in range(3)})
    assert fed.n_launches == 2
    for i in range(3):
        single = FederatedEngine(GlmShards([Xs[i]], [ys[i]], groups=[[0, 1, 0][i]], n_groups=2, kernel="simt"), backend="collective")
        np.testing.assert_allclose(res[i][0], single.logp_grad(ic, betas[i])[0], rtol=2e-6)
    # (3) more distinct vectors than chains is an error, not a silent re-launch
    _, _, small = _glm_nodes(n_chains=1)
    fed1 = NodeFederation(FederatedEngine(small, backend="collective"))
    with pytest.raises(FederationError, match="n_chains"):
        fed1.evaluate_nodes({0: (ic, beta), 1: (ic, beta * 2)})
    with pytest.raises(FederationError, match="does not exist"):
        fed1.evaluate_node(7, ic, beta)
    # a model without per-node blocks cannot be viewed per node
    with pytest.raises(FederationError, match="node_ids"):
        NodeFederation(pooled)


def test_ode_nodes_take_their_own_parameters_through_federated_ops():
    from pytensor_federated_b200.models import OdeShards, synth_lv_shard

    shards = [synth_lv_shard(12, 5, seed=s, device="cpu") for s in range(2)]
    args = lambda idx: ([shards[i][0] for i in idx], [shards[i][1] for i in idx], [shards[i][2] for i in idx], [shards[i][3] for i in idx])
    model = OdeShards(*args([0, 1]), node_ids=[0, 1], n_nodes=2)
    assert model.n_theta_words == 8 and model.n_vals == 10
    eng = FederatedEngine(model, backend="collective")
    th0, th1 = np.array([1.0, 0.4, 0.8, 0.2]), np.array([0.9, 0.45, 0.75, 0.25])
    # the ComputeFunc view: one vector for everybody -> summed gradient; one row per node -> per-node gradients
    logp, grad_sum = eng.evaluate(th0)
    logp2, grad_rows = eng.evaluate(np.stack([th0, th0]))
    np.testing.assert_allclose(logp, logp2, rtol=1e-12)
    np.testing.assert_allclose(grad_rows.sum(0), grad_sum, rtol=1e-10)
    # node view with different parameters per node
    fed = NodeFederation(eng)
    res = fed.evaluate_nodes({0: (th0,), 1: (th1,)})
    for i, th in enumerate((th0, th1)):
        single = FederatedEngine(OdeShards(*args([i])), backend="collective")
        want = single.logp_grad(th)
        np.testing.assert_allclose(res[i][0], want[0], rtol=1e-10)
        np.testing.assert_allclose(res[i][1][0], want[1][0], rtol=1e-9)
    # federated Ops: log-potentials of both nodes in one graph, gradients w.r.t. both parameter vectors
    ops = fed.node_ops()
    a, b = at.vector("theta_a"), at.vector("theta_b")
    total = ops[0](a)[0] + ops[1](b)[0]
    fn = function([a, b], [total, *grad(total, [a, b])])
    fed.n_launches = 0
    val, ga, gb = fn(th0, th1)
    assert fed.n_launches == 1
    np.testing.assert_allclose(val, float(res[0][0]) + float(res[1][0]), rtol=1e-12)
    np.testing.assert_allclose(ga, res[0][1][0], rtol=1e-12)
    np.testing.assert_allclose(gb, res[1][1][0], rtol=1e-12)
    # the whole federation as ONE Op: a [n_nodes, n_params] matrix in, summed logp and the matrix of gradients out
    lp, (g_all,) = fed.all_nodes_func()(np.stack([th0, th1]))
    np.testing.assert_allclose(lp, val, rtol=1e-12)
    np.testing.assert_allclose(g_all, np.stack([ga, gb]), rtol=1e-12)


def test_replicated_gpu_nodes_balance_and_fail_over():
    """Replicated-shard mode: two engines hold the same data; clients spread over them by load and a lost
    replica is replaced transparently (reference semantics: service.py:239-275, :407-416)."""
    from pytensor_federated_b200 import service
    from pytensor_federated_b200.federation import register_replicas

    x, y, sigma = make_demo_data()
    engines = [FederatedEngine(LinregShards([x], [y], [sigma]), backend="collective") for _ in range(2)]
    addresses = register_replicas(engines, host="replica", first_port=10)
    try:
        c1 = LogpGradServiceClient(hosts_and_ports=addresses)
        c2 = LogpGradServiceClient(hosts_and_ports=addresses)
        want = engines[0].logp_grad(np.array(0.4), np.array(1.2))
        for c in (c1, c2):
            logp, grads = c.evaluate(np.array(0.4), np.array(1.2))
            np.testing.assert_allclose(logp, want[0], rtol=1e-12)
        ports = sorted(int(service._privates[service.thread_pid_id(c._client)].channel._port) for c in (c1, c2))
        assert ports == [10, 11]                          # the second client went to the idle replica
        # lose the replica c1 is connected to: its next call lands on the survivor, same answer
        lost = int(service._privates[service.thread_pid_id(c1._client)].channel._port) - 10
        engines[lost].shutdown()
        logp, grads = c1.evaluate(np.array(0.4), np.array(1.2))
        np.testing.assert_allclose(logp, want[0], rtol=1e-12)
        np.testing.assert_allclose(grads, want[1], rtol=1e-12)
        assert int(service._privates[service.thread_pid_id(c1._client)].channel._port) - 10 == 1 - lost
        # both gone: a connection-level error, like a fleet of dead servers
        engines[1 - lost].shutdown()
        with pytest.raises((service.StreamTerminatedError, TimeoutError)):
            c1.evaluate(np.array(0.4), np.array(1.2), retries=1)
        del c1, c2
    finally:
        for host, port in addresses:
            service.unregister_local_node(host, port)


def test_whole_federation_as_one_op_matches_the_per_node_ops():
    """`all_nodes_op`: vector of node intercepts + shared slope in, summed logp and gradients out — the same numbers
    as one Op per node, from a graph whose size does not depend on the number of nodes."""
    fed = NodeFederation(_three_node_engine())
    offs = np.linspace(-1.5, 1.5, 3)
    icpt, slope = at.vector("intercept"), at.scalar("slope")
    # one Op per node (reference pattern)
    total = None
    for i, op in enumerate(fed.node_ops()):
        logp, *_ = op(icpt[i] + offs[i], slope)
        total = logp if total is None else total + logp
    per_node = function([icpt, slope], [total, *grad(total, [icpt, slope])])
    # one Op for the federation
    logp_all, *_ = fed.all_nodes_op()(icpt + at.as_tensor(offs), slope)
    whole = function([icpt, slope], [logp_all, *grad(logp_all, [icpt, slope])])
    assert len(whole.maker.fgraph.toposort()) < len(per_node.maker.fgraph.toposort()) / 3
    args = (np.array([0.1, 0.2, 0.3]), 0.5)
    fed.n_launches = 0
    got = whole(*args)
    assert fed.n_launches == 1
    for g, w in zip(got, per_node(*args)):
        np.testing.assert_allclose(g, w, rtol=1e-12)
    f = fed.all_nodes_func()
    lp, (da, db) = f(np.array([0.1, 0.2, 0.3]) + offs, np.array([0.5, 0.5, 0.5]))
    np.testing.assert_allclose(lp, got[0], rtol=1e-12)
    assert da.shape == (3,) and db.shape == (3,)
    np.testing.assert_allclose(db.sum(), got[2], rtol=1e-12)
