"""Gradient-free end to end: ``LogpServiceClient`` -> ``LogpOp`` / ``AsyncLogpOp`` -> Metropolis, in one
process and in worker processes, over real gRPC (reference: ``test_wrapper_ops.py:68-118`` and ``:291-317``)."""
import functools

import numpy as np
import pytest

from pytensor_federated_b200 import AsyncLogpOp, LogpOp, LogpServiceClient
from pytensor_federated_b200._graph_backend import at
from pytensor_federated_b200.sampling import Model, metropolis_sample, sample_parallel

from _helpers import ServerProcess, slope_posterior_logp, straight_line_blackbox

pytestmark = pytest.mark.timeout(600)
GOLDEN = -1511.41423640139   # logp(intercept=0.4, slope=1.2) of the reference's fixture


@pytest.fixture(scope="module")
def node():
    with ServerProcess(func="blackbox_logp") as server:
        yield server


def test_metropolis_recovers_a_correlated_gaussian_and_resumes():
    mu, sd = np.array([1.0, -2.0]), np.array([1.0, 0.1])
    target = lambda t: float(-0.5 * np.sum(((t - mu) / sd) ** 2))
    res = metropolis_sample(target, np.zeros(2), draws=6000, tune=3000, seed=3)
    assert res.n_logp_evals == 1 + 9000 and 0.15 < res.accept_rate < 0.6
    np.testing.assert_allclose(res.samples.mean(0), mu, atol=0.12)
    np.testing.assert_allclose(res.samples.std(0), sd, rtol=0.15)
    assert res.inv_mass[0] > 10 * res.inv_mass[1]              # the proposal learned the scales
    more = metropolis_sample(target, draws=500, resume=res)    # continues with the tuned proposal
    assert more.step_size == res.step_size and more.samples.shape == (500, 2)
    with pytest.raises(ValueError, match="finite"):
        metropolis_sample(lambda t: -np.inf, np.zeros(1))


def test_blackbox_golden_value_locally_and_over_grpc(node):
    np.testing.assert_allclose(straight_line_blackbox(0.4, 1.2), GOLDEN, rtol=1e-12)
    client = LogpServiceClient("127.0.0.1", node.port)
    result = client(0.4, 1.2)
    assert isinstance(result, np.ndarray) and result.shape == ()
    np.testing.assert_allclose(result, GOLDEN, rtol=1e-12)
    del client


@pytest.mark.parametrize("use_async", [False, True], ids=["LogpOp", "AsyncLogpOp"])
def test_metropolis_through_the_op_in_process(node, use_async):
    """cores = 1: three chains one after the other, the Op inside a compiled graph (no gradient exists)."""
    client = LogpServiceClient("127.0.0.1", node.port)
    op = AsyncLogpOp(client.evaluate_async) if use_async else LogpOp(client)
    m = Model()
    slope = m.Normal("slope", 0.0, 2.0)
    m.Potential("L", op(at.constant(0.5), slope))
    np.testing.assert_allclose(m.logp(np.array([1.2])) + 0.5 * (1.2 / 2) ** 2 + np.log(2.0) + 0.9189385332046727,
                               straight_line_blackbox(0.5, 1.2), rtol=1e-10)
    results, columns = m.sample_metropolis(draws=300, tune=200, chains=3, seed=1234)
    assert len(results) == 3 and columns["slope"].shape == (300, 3)
    np.testing.assert_allclose(np.median(columns["slope"]), 2.0, atol=0.1)
    del op, client


@pytest.mark.parametrize("use_async", [False, True], ids=["LogpOp", "AsyncLogpOp"])
def test_metropolis_chains_in_worker_processes(node, use_async):
    """cores = 4: every chain lives in its own process with its own connection to the node."""
    factory = functools.partial(slope_posterior_logp, node.port, use_async)
    results = sample_parallel(factory, np.zeros(1), chains=3, cores=4, sampler="metropolis", draws=300, tune=200, seed=1234)
    assert len(results) == 3
    pooled = np.concatenate([r.samples[:, 0] for r in results])
    np.testing.assert_allclose(np.median(pooled), 2.0, atol=0.1)
    # different seeds, different chains
    assert not np.array_equal(results[0].samples, results[1].samples)
