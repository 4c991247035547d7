"""Inference drivers on analytic targets and on a federated linear model (CPU)."""
import numpy as np
import pytest

from pytensor_federated_b200 import LogpGradOp
from pytensor_federated_b200.models import LinregShards, make_demo_data
from pytensor_federated_b200.parallel import FederatedEngine
from pytensor_federated_b200.sampling import Model, find_map, hmc_sample, nuts_sample
from pytensor_federated_b200.sampling.mcmc import effective_sample_size

pytestmark = pytest.mark.timeout(300)


def _gauss(mu, sd):
    mu, sd = np.asarray(mu, float), np.asarray(sd, float)

    def f(x):
        z = (x - mu) / sd
        return float(-0.5 * np.sum(z * z)), -z / sd

    return f


def test_find_map_quadratic():
    x, info = find_map(_gauss([1.0, -2.0, 0.5], [1.0, 0.1, 3.0]), np.zeros(3))
    np.testing.assert_allclose(x, [1.0, -2.0, 0.5], atol=1e-5)
    assert info["converged"]


@pytest.mark.parametrize("sampler", [hmc_sample, nuts_sample])
def test_samplers_recover_gaussian_moments(sampler):
    mu, sd = np.array([1.0, -2.0, 0.5]), np.array([1.0, 0.2, 3.0])
    res = sampler(_gauss(mu, sd), np.zeros(3), draws=1500, tune=600, seed=3)
    assert res.accept_rate > 0.55
    assert np.all(np.abs(res.samples.mean(0) - mu) < 4 * sd / np.sqrt(200))
    np.testing.assert_allclose(res.samples.std(0), sd, rtol=0.2)
    assert effective_sample_size(res.samples[:, 0]) > 100
    assert set(res.summary(["a", "b", "c"])) == {"a", "b", "c"}


def test_model_builder_gradient_matches_finite_differences():
    m = Model()
    mu = m.Normal("mu", 0.0, 2.0)
    a = m.Normal("a", mu, 0.5, size=3)
    m.Potential("pot", (a * a).sum() * -0.1)
    theta = np.array([0.3, 0.1, -0.2, 0.4])
    lp, g = m.logp_dlogp(theta)
    eps = 1e-6
    for i in range(4):
        d = np.zeros(4)
        d[i] = eps
        fd = (m.logp_dlogp(theta + d)[0] - m.logp_dlogp(theta - d)[0]) / (2 * eps)
        np.testing.assert_allclose(g[i], fd, rtol=1e-5, atol=1e-7)
    assert m.names() == ["mu", "a[0]", "a[1]", "a[2]"]


def test_federated_linear_model_map_and_nuts():
    """The reference demo end to end, minus PyMC: MAP + NUTS through a LogpGradOp."""
    x, y, sigma = make_demo_data()
    eng = FederatedEngine(LinregShards([x], [y], [sigma]), backend="collective")
    op = LogpGradOp(eng.logp_grad)
    m = Model()
    intercept = m.Normal("intercept", 0.0, 10.0)
    slope = m.Normal("slope", 0.0, 10.0)
    logp, *_ = op(intercept, slope)
    m.Potential("likelihood", logp)
    import scipy.stats

    mle = scipy.stats.linregress(x, y)
    theta_map, info = find_map(m.logp_dlogp, np.zeros(2))
    np.testing.assert_allclose(theta_map, [mle.intercept, mle.slope], atol=2e-3)
    res = nuts_sample(m.logp_dlogp, theta_map, draws=300, tune=300, seed=1)
    np.testing.assert_allclose(np.median(res.samples[:, 1]), mle.slope, atol=0.1)
    assert res.divergences == 0
    # one engine evaluation per logp+gradient evaluation (merge pass), not two
    before = eng.n_evals
    m.logp_dlogp(theta_map)
    assert eng.n_evals == before + 1


def test_checkpoint_and_resume_continue_the_same_chain(tmp_path):
    from pytensor_federated_b200.sampling import SamplerResult

    target = _gauss([0.5, -1.0], [1.0, 0.3])
    full = nuts_sample(target, np.zeros(2), draws=120, tune=200, seed=5)
    first = nuts_sample(target, np.zeros(2), draws=60, tune=200, seed=5)
    first.save(tmp_path / "chain")
    restored = SamplerResult.load(tmp_path / "chain")
    assert restored.step_size == first.step_size
    np.testing.assert_array_equal(restored.inv_mass, first.inv_mass)
    second = nuts_sample(target, resume=restored, draws=60)
    # identical random stream + identical adaptation => the resumed half reproduces the uninterrupted run
    np.testing.assert_allclose(np.concatenate([first.samples, second.samples]), full.samples, rtol=0, atol=1e-12)


def test_batched_hmc_runs_chains_in_lock_step():
    from pytensor_federated_b200.sampling import hmc_sample_batched

    mu, sd = np.array([1.0, -2.0, 0.5]), np.array([1.0, 0.2, 3.0])
    calls = []

    def batch(theta):
        calls.append(theta.shape)
        z = (theta - mu) / sd
        return -0.5 * np.sum(z * z, axis=1), -z / sd

    K = 6
    res = hmc_sample_batched(batch, np.zeros((K, 3)), draws=600, tune=400, n_leapfrog=12, seed=2)
    assert res.samples.shape == (600, K, 3) and set(calls) == {(K, 3)}
    assert res.n_batched_evals == 1 + (600 + 400) * 12          # ONE batched call per leapfrog step, not K
    pooled = res.samples.reshape(-1, 3)
    assert np.all(np.abs(pooled.mean(0) - mu) < 4 * sd / np.sqrt(300))
    np.testing.assert_allclose(pooled.std(0), sd, rtol=0.15)
    assert np.all(res.rhat() < 1.1) and np.all(res.accept_rate > 0.5)


def test_batched_hmc_on_a_multichain_glm_engine():
    """K chains of a logistic GLM through the collective (CPU) engine: one evaluate per leapfrog."""
    import torch

    from pytensor_federated_b200.models import GlmShards
    from pytensor_federated_b200.sampling import glm_batch_fn, hmc_sample_batched

    torch.manual_seed(0)
    X = torch.randn(400, 4).to(torch.bfloat16)
    beta_true = np.array([0.8, -0.5, 0.0, 0.3])
    y = (torch.rand(400) < torch.sigmoid(X.float() @ torch.tensor(beta_true, dtype=torch.float32) + 0.2)).float()
    K = 4
    eng = FederatedEngine(GlmShards([X], [y], n_chains=K, kernel="simt"), backend="collective")
    res = hmc_sample_batched(glm_batch_fn(eng, 1), np.zeros((K, 5)), draws=150, tune=150, n_leapfrog=8, step_size=0.05, seed=1)
    assert eng.n_evals == res.n_batched_evals
    post = res.samples.reshape(-1, 5).mean(0)
    np.testing.assert_allclose(post[1:], beta_true, atol=0.35)


def test_hierarchical_glm_with_vector_inputs_through_logp_grad_op():
    """Partial pooling over shard intercepts: vector-valued Op inputs/gradients through the graph IR."""
    import torch

    from pytensor_federated_b200.models import GlmShards

    torch.manual_seed(2)
    true_ic = np.array([-0.6, 0.1, 0.8])
    beta_true = np.array([0.7, -0.4])
    Xs, ys = [], []
    for g in range(3):
        X = torch.randn(600, 2)
        p = torch.sigmoid(X @ torch.tensor(beta_true, dtype=torch.float32) + true_ic[g])
        Xs.append(X.to(torch.bfloat16))
        ys.append((torch.rand(600) < p).float())
    eng = FederatedEngine(GlmShards(Xs, ys, groups=[0, 1, 2], n_groups=3, kernel="simt"), backend="collective")
    op = LogpGradOp(eng.logp_grad)
    m = Model()
    mu = m.Normal("mu", 0.0, 1.0)
    ic = m.Normal("intercept", mu, 0.7, size=3)
    beta = m.Normal("beta", 0.0, 1.0, size=2)
    logp, *_ = op(ic, beta)
    m.Potential("likelihood", logp)
    theta0 = np.zeros(m.dim)
    lp, g = m.logp_dlogp(theta0)
    eps = 1e-3  # theta is float32 in the engine's mailbox -> coarse finite differences
    for i in range(m.dim):
        d = np.zeros(m.dim)
        d[i] = eps
        fd = (m.logp_dlogp(theta0 + d)[0] - m.logp_dlogp(theta0 - d)[0]) / (2 * eps)
        np.testing.assert_allclose(g[i], fd, rtol=2e-2, atol=0.5)
    theta_map, info = find_map(m.logp_dlogp, theta0)
    point = m.point(theta_map)
    np.testing.assert_allclose(point["beta"], beta_true, atol=0.25)
    assert np.all(np.diff(point["intercept"]) > 0)  # ordering of the group intercepts is recovered
    res = nuts_sample(m.logp_dlogp, theta_map, draws=80, tune=80, seed=3)
    assert res.divergences == 0 and res.accept_rate > 0.6


def test_half_normal_and_model_conveniences():
    """Unknown noise scale with a HalfNormal prior (log transform + Jacobian), Model.find_map / .sample."""
    from pytensor_federated_b200._graph_backend import at

    rng = np.random.default_rng(0)
    data = rng.normal(1.5, 0.6, size=400)
    m = Model()
    mu = m.Normal("mu", 0.0, 10.0)
    sigma = m.HalfNormal("sigma", 5.0)
    z = (at.as_tensor(data) - mu) / sigma
    m.Potential("lik", (-0.5 * z * z).sum() - 400.0 * at.log(sigma))
    # gradient of the transformed density vs finite differences
    theta = np.array([1.0, np.log(0.8)])
    lp, g = m.logp_dlogp(theta)
    for i in range(2):
        d = np.zeros(2)
        d[i] = 1e-6
        fd = (m.logp_dlogp(theta + d)[0] - m.logp_dlogp(theta - d)[0]) / 2e-6
        np.testing.assert_allclose(g[i], fd, rtol=1e-5)
    point, info = m.find_map()
    assert info["converged"] and abs(point["mu"] - data.mean()) < 0.05 and abs(point["sigma"] - data.std()) < 0.05
    res, cols = m.sample(draws=300, tune=300, seed=4)
    assert set(cols) == {"mu", "sigma_log__", "sigma"} and np.all(cols["sigma"] > 0)
    assert abs(np.median(cols["sigma"]) - data.std()) < 0.06 and res.divergences == 0


def test_diagnostics_on_known_processes():
    from pytensor_federated_b200.sampling import effective_sample_size, split_rhat, summarize

    rng = np.random.default_rng(0)
    iid = rng.normal(size=(2000, 4))
    assert abs(split_rhat(iid) - 1.0) < 0.01
    assert 6000 < effective_sample_size(iid) < 10000            # ~ number of draws
    # AR(1) with phi = 0.9: ESS ~ N (1 - phi) / (1 + phi) = N / 19
    ar = np.zeros((4000, 4))
    eps = rng.normal(size=ar.shape)
    for t in range(1, ar.shape[0]):
        ar[t] = 0.9 * ar[t - 1] + eps[t]
    ess = effective_sample_size(ar)
    assert 16000 / 19 / 1.6 < ess < 16000 / 19 * 1.6
    # chains stuck in different places
    assert split_rhat(iid + np.array([0.0, 0.0, 3.0, 3.0])) > 1.5
    # a trend inside every chain is caught by the split
    assert split_rhat(iid + np.linspace(0, 3, 2000)[:, None]) > 1.2
    table = summarize({"a": iid, "v": rng.normal(2.0, 0.5, size=(500, 2, 3))})
    assert set(table) == {"a", "v[0]", "v[1]", "v[2]"}
    assert abs(table["v[1]"]["mean"] - 2.0) < 0.1 and abs(table["v[1]"]["sd"] - 0.5) < 0.05
    assert table["a"]["q3"] < -1.5 and table["a"]["q97"] > 1.5
    with pytest.raises(ValueError):
        split_rhat(np.zeros((3, 2)))


def test_model_sample_multiple_chains_converge():
    from pytensor_federated_b200._graph_backend import at
    from pytensor_federated_b200.sampling import summarize

    m = Model()
    mu = m.Normal("mu", 0.0, 10.0)
    data = np.random.default_rng(1).normal(0.7, 1.0, size=50)
    z = at.as_tensor(data) - mu
    m.Potential("lik", (-0.5 * z * z).sum())
    results, cols = m.sample(draws=300, tune=300, chains=3, seed=10)
    assert len(results) == 3 and cols["mu"].shape == (300, 3)
    row = summarize(cols)["mu"]
    assert row["rhat"] < 1.03 and row["ess"] > 150
    assert abs(row["mean"] - data.mean()) < 0.08 and abs(row["sd"] - 1 / np.sqrt(50)) < 0.04


def test_chains_in_worker_processes_over_replicas(monkeypatch):
    """Replica parallelism: one chain per process, each with its own balanced connection to a fleet of
    identical servers (reference: ``pm.sample(cores=...)`` in test_wrapper_ops.py:305-317)."""
    import functools

    from _helpers import ServerProcess, remote_gaussian_logp_dlogp
    from pytensor_federated_b200.sampling import sample_parallel, summarize

    monkeypatch.setenv("B200FED_CONNECT_SLEEP", "0,0")
    servers = [ServerProcess(func="gaussian").start() for _ in range(2)]
    try:
        factory = functools.partial(remote_gaussian_logp_dlogp, [("127.0.0.1", s.port) for s in servers])
        results = sample_parallel(factory, np.zeros(2), chains=2, cores=2, sampler="hmc", draws=150, tune=150, seed=5)
    finally:
        for s in servers:
            s.terminate()
    assert len(results) == 2 and all(r.samples.shape == (150, 2) for r in results)
    assert not np.array_equal(results[0].samples, results[1].samples)      # different seeds
    draws = np.stack([r.samples for r in results], axis=1)                  # [draws, chains, dim]
    table = summarize({"theta": draws})
    for row in table.values():
        assert abs(row["mean"] - 1.0) < 0.15 and abs(row["sd"] - 0.5) < 0.12 and row["rhat"] < 1.1
    with pytest.raises(ValueError):
        sample_parallel(factory, np.zeros(2), sampler="gibbs")


def test_constrained_priors_sample_their_own_distribution():
    """Exponential (log scale) and Uniform (logit scale): with no likelihood the draws follow the prior."""
    from pytensor_federated_b200.sampling import summarize

    m = Model()
    m.Exponential("rate", lam=2.0)
    m.Uniform("u", lower=-1.0, upper=3.0)
    theta = np.array([0.3, -0.4])
    lp, g = m.logp_dlogp(theta)
    for i in range(2):
        d = np.zeros(2)
        d[i] = 1e-6
        fd = (m.logp_dlogp(theta + d)[0] - m.logp_dlogp(theta - d)[0]) / 2e-6
        np.testing.assert_allclose(g[i], fd, rtol=1e-5, atol=1e-7)
    _, cols = m.sample(draws=600, tune=400, chains=2, start=np.zeros(2), seed=3)
    assert set(cols) == {"rate_log__", "u_interval__", "rate", "u"}
    assert np.all(cols["rate"] > 0) and np.all((cols["u"] > -1.0) & (cols["u"] < 3.0))
    table = summarize({"rate": cols["rate"], "u": cols["u"]})
    assert abs(table["rate"]["mean"] - 0.5) < 0.08 and abs(table["rate"]["sd"] - 0.5) < 0.12      # Exp(2)
    assert abs(table["u"]["mean"] - 1.0) < 0.2 and abs(table["u"]["sd"] - 4 / np.sqrt(12)) < 0.15   # U(-1, 3)
    with pytest.raises(ValueError):
        m.Uniform("bad", 1.0, 1.0)


def test_batched_chains_checkpoint_and_resume(tmp_path):
    """Save after 60 draws, resume for 40 more: identical to one 100-draw run (same adaptation, same RNG stream)."""
    from pytensor_federated_b200.sampling import BatchedResult, hmc_sample_batched

    def fn(theta):
        return -0.5 * np.sum(theta * theta / np.array([1.0, 4.0]), axis=1), -theta / np.array([1.0, 4.0])

    x0 = np.zeros((3, 2))
    full = hmc_sample_batched(fn, x0, draws=100, tune=150, seed=11)
    first = hmc_sample_batched(fn, x0, draws=60, tune=150, seed=11)
    first.save(str(tmp_path / "ckpt"))
    loaded = BatchedResult.load(str(tmp_path / "ckpt"))
    np.testing.assert_array_equal(loaded.samples, first.samples)
    np.testing.assert_array_equal(loaded.inv_mass, first.inv_mass)
    rest = hmc_sample_batched(fn, None, draws=40, resume=loaded)
    np.testing.assert_allclose(np.concatenate([first.samples, rest.samples]), full.samples, rtol=0, atol=0)
    np.testing.assert_array_equal(rest.step_size, first.step_size)


def test_hmc_chain_resumes_bit_identically(tmp_path):
    from pytensor_federated_b200.sampling import SamplerResult

    def fn(theta):
        return float(-0.5 * np.sum(theta * theta)), -theta

    full = hmc_sample(fn, np.zeros(3), draws=80, tune=120, seed=2)
    first = hmc_sample(fn, np.zeros(3), draws=50, tune=120, seed=2)
    first.save(str(tmp_path / "chain"))
    rest = hmc_sample(fn, None, draws=30, resume=SamplerResult.load(str(tmp_path / "chain")))
    np.testing.assert_array_equal(np.concatenate([first.samples, rest.samples]), full.samples)


def test_glm_batch_fn_tiles_more_chains_than_the_kernel_takes_per_launch():
    """7 chains on engines with capacity 1 / 3 / 7: 7 / 3 / 1 launches, same numbers (padding rows dropped)."""
    import torch

    from pytensor_federated_b200.models import GlmShards
    from pytensor_federated_b200.sampling import glm_batch_fn

    torch.manual_seed(0)
    X = torch.randn(300, 8).to(torch.bfloat16)
    y = (torch.rand(300) < 0.5).float()
    theta = np.random.default_rng(0).normal(size=(7, 9)) * 0.2
    results = {}
    for cap, launches in ((1, 7), (3, 3), (7, 1)):
        eng = FederatedEngine(GlmShards([X], [y], n_chains=cap, kernel="simt" if cap == 1 else "tc"), backend="collective")
        results[cap] = glm_batch_fn(eng, 1)(theta)
        assert eng.n_evals == launches
        assert results[cap][0].shape == (7,) and results[cap][1].shape == (7, 9)
    for cap in (3, 7):
        np.testing.assert_allclose(results[cap][0], results[1][0], rtol=1e-5)
        np.testing.assert_allclose(results[cap][1], results[1][1], rtol=1e-4, atol=1e-4)
