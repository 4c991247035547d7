"""The two demo CLIs end to end over real gRPC (CPU): node pool + hierarchical model."""
import os
import subprocess
import sys
import time

import numpy as np
import pytest

from _helpers import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(300)


@pytest.fixture(scope="module")
def node_pool():
    ports = [free_port() for _ in range(2)]
    proc = subprocess.Popen(
        [sys.executable, os.path.join(ROOT, "demo_node.py"), "--bind", "127.0.0.1", "--ports", ",".join(map(str, ports)),
         "--device", "cpu"],
        cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
    )
    from pytensor_federated_b200 import service
    from pytensor_federated_b200.utils import get_useful_event_loop

    deadline = time.time() + 120
    loop = get_useful_event_loop()
    while time.time() < deadline:
        loads = loop.run_until_complete(service.get_loads_async([("127.0.0.1", p) for p in ports], timeout=1))
        if all(l is not None for l in loads):
            break
        time.sleep(0.5)
    else:
        proc.terminate()
        pytest.fail("demo_node workers did not come up")
    yield ports
    proc.terminate()
    proc.wait(20)


@pytest.mark.parametrize("use_async", [False, True])
def test_hierarchical_demo_over_grpc(node_pool, use_async):
    sys.path.insert(0, ROOT)
    import demo_model
    import scipy.stats

    from pytensor_federated_b200.models import make_demo_data

    ops, client = demo_model.remote_ops_grpc("127.0.0.1", node_pool, 3, use_async)
    res = demo_model.run_model(ops, 3, tune=150, draws=100)
    x, y, _ = make_demo_data()
    slope = scipy.stats.linregress(x, y).slope
    names = ["intercept_mu", "intercept[0]", "intercept[1]", "intercept[2]", "slope"]
    assert abs(np.median(res.samples[:, names.index("slope")]) - slope) < 0.1
    assert res.divergences == 0


# ---- the black-box node model itself (reference: test_demo_node.py:10-109) ------------------------


def _import_demo_node():
    sys.path.insert(0, ROOT)
    import demo_node

    return demo_node


def test_blackbox_logp_and_gradient_match_closed_form():
    import scipy.stats

    from pytensor_federated_b200.models import make_demo_data

    x, y, sigma = make_demo_data()
    blackbox = _import_demo_node().LinearModelBlackbox(x, y, sigma, device="cpu")
    a, b = 1.3, 0.45
    logp, grads = blackbox(np.array(a), np.array(b))
    resid = y - (a + b * x)
    np.testing.assert_allclose(logp, scipy.stats.norm.logpdf(y, a + b * x, sigma).sum(), rtol=1e-12)
    np.testing.assert_allclose(grads[0], resid.sum() / sigma**2, rtol=1e-10)
    np.testing.assert_allclose(grads[1], (resid * x).sum() / sigma**2, rtol=1e-10)
    assert logp.shape == () and all(g.shape == () for g in grads)


def test_blackbox_delay_pads_the_call():
    from pytensor_federated_b200.models import make_demo_data

    x, y, sigma = make_demo_data()
    blackbox = _import_demo_node().LinearModelBlackbox(x, y, sigma, delay=0.3, device="cpu")
    blackbox(np.array(1.0), np.array(1.0))  # warm-up
    t0 = time.perf_counter()
    blackbox(np.array(1.0), np.array(1.0))
    assert 0.3 <= time.perf_counter() - t0 < 1.0


def test_blackbox_in_a_graph_equals_the_native_model():
    """A LogpGradOp around the black box vs the same likelihood written in the graph IR: same logp,
    same gradients, same MAP (reference: test_linear_model_equivalence / ..._findmap)."""
    from pytensor_federated_b200 import LogpGradOp
    from pytensor_federated_b200._graph_backend import at, function, grad
    from pytensor_federated_b200.models import make_demo_data
    from pytensor_federated_b200.sampling import find_map

    x, y, sigma = make_demo_data()
    blackbox = _import_demo_node().LinearModelBlackbox(x, y, sigma, device="cpu")

    def build(native: bool):
        a, b = at.scalar("a"), at.scalar("b")
        if native:
            z = (at.as_tensor(y) - (a + b * at.as_tensor(x))) / sigma
            logp = (-0.5 * z * z).sum() - len(x) * np.log(sigma * np.sqrt(2 * np.pi))
        else:
            logp = LogpGradOp(blackbox)(a, b)[0]
        prior = -0.5 * (a * a + b * b) / 100.0
        total = logp + prior
        return function([a, b], [total] + list(grad(total, [a, b])))

    f_native, f_blackbox = build(True), build(False)
    for point in [(0.0, 0.0), (1.5, 0.5), (-2.0, 3.0)]:
        got, want = f_blackbox(*point), f_native(*point)
        for g, w in zip(got, want):
            np.testing.assert_allclose(g, w, rtol=1e-9, atol=1e-9)

    def as_logp_dlogp(f):
        def fn(theta):
            lp, da, db = f(theta[0], theta[1])
            return float(lp), np.array([float(da), float(db)])

        return fn

    map_native, info1 = find_map(as_logp_dlogp(f_native), np.zeros(2))
    map_blackbox, info2 = find_map(as_logp_dlogp(f_blackbox), np.zeros(2))
    assert info1["converged"] and info2["converged"]
    np.testing.assert_allclose(map_blackbox, map_native, atol=1e-5)
    assert abs(map_native[0] - 1.5) < 0.6 and abs(map_native[1] - 0.5) < 0.15
