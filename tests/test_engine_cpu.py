"""Engine plumbing on CPU: collective backend, single process and 2-rank gloo."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from pytensor_federated_b200.models import GlmShards, LinregShards, OdeShards, make_demo_data, synth_lv_shard
from pytensor_federated_b200.parallel import FederatedEngine, FederationError

pytestmark = pytest.mark.timeout(300)


def _scipy_linreg(x, y, sigma, a, b):
    import scipy.stats

    return scipy.stats.norm.logpdf(y, loc=a + b * x, scale=sigma).sum()


def test_linreg_reference_matches_scipy_and_finite_differences():
    x, y, sigma = make_demo_data()
    model = LinregShards([x], [y], [sigma])
    eng = FederatedEngine(model, backend="collective")
    logp, da, db = eng.evaluate(np.array(0.4), np.array(1.2))
    assert logp.shape == () and da.shape == () and db.shape == ()
    np.testing.assert_allclose(logp, _scipy_linreg(x, y, sigma, 0.4, 1.2), rtol=1e-12)
    eps = 1e-6
    fd_a = (_scipy_linreg(x, y, sigma, 0.4 + eps, 1.2) - _scipy_linreg(x, y, sigma, 0.4 - eps, 1.2)) / (2 * eps)
    fd_b = (_scipy_linreg(x, y, sigma, 0.4, 1.2 + eps) - _scipy_linreg(x, y, sigma, 0.4, 1.2 - eps)) / (2 * eps)
    np.testing.assert_allclose([da, db], [fd_a, fd_b], rtol=1e-6)
    # speculative launches belong to the fused backend; elsewhere the switch is a no-op that says so
    assert eng.set_speculative(500.0) is False and eng.speculative is False
    eng.shutdown()
    with pytest.raises(FederationError):
        eng.evaluate(np.array(0.0), np.array(0.0))


def test_linreg_per_shard_parameters():
    rng = np.random.default_rng(1)
    xs = [rng.normal(size=20) for _ in range(3)]
    ys = [rng.normal(size=20) for _ in range(3)]
    model = LinregShards(xs, ys, [0.5, 1.0, 2.0])
    eng = FederatedEngine(model, backend="collective")
    a = np.array([0.1, 0.2, 0.3])
    logp, da, db = eng.evaluate(a, np.array(0.7))
    assert da.shape == (3,) and db.shape == ()
    expected = sum(_scipy_linreg(x, y, s, ai, 0.7) for x, y, s, ai in zip(xs, ys, [0.5, 1.0, 2.0], a))
    np.testing.assert_allclose(logp, expected, rtol=1e-12)


def test_concurrent_evaluate_keeps_every_callers_shapes_and_values():
    """``evaluate`` is called from several service threads (``offload=True``, ``DynamicBatcher``): packing,
    evaluation and unpacking of the shared buffers must be atomic and the input shapes per call."""
    import threading

    rng = np.random.default_rng(5)
    xs = [rng.normal(size=30) for _ in range(4)]
    ys = [rng.normal(size=30) for _ in range(4)]
    model = LinregShards(xs, ys, [1.0] * 4)
    eng = FederatedEngine(model, backend="collective")
    vec = np.array([0.1, 0.2, 0.3, 0.4])
    want_scalar = [np.array(v) for v in eng.evaluate(np.array(0.5), np.array(0.25))]
    want_vector = [np.array(v) for v in eng.evaluate(vec, vec[::-1].copy())]
    assert want_scalar[1].shape == () and want_vector[1].shape == (4,)
    errors = []

    def worker(scalar: bool):
        try:
            for _ in range(150):
                got = eng.evaluate(np.array(0.5), np.array(0.25)) if scalar else eng.evaluate(vec, vec[::-1].copy())
                want = want_scalar if scalar else want_vector
                for g, w in zip(got, want):
                    assert g.shape == w.shape
                    np.testing.assert_array_equal(g, w)
        except Exception as ex:  # noqa: BLE001 - reported by the main thread
            errors.append(ex)

    threads = [threading.Thread(target=worker, args=(i % 2 == 0,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[0]
    raw = eng.evaluate_raw([vec, vec])
    raw2 = eng.evaluate_raw([np.array(0.0), np.array(0.0)])
    assert raw is not raw2 and not np.array_equal(raw, raw2)   # fresh copies, not views of one buffer


def test_glm_reference_gradient_matches_autograd():
    torch.manual_seed(0)
    X = torch.randn(300, 16).to(torch.bfloat16)
    y = (torch.rand(300) < 0.5).float()
    for family in ("logistic", "poisson", "gaussian"):
        model = GlmShards([X[:100], X[100:]], [y[:100], y[100:]], groups=[0, 1], n_groups=2, family=family)
        ic = np.array([0.2, -0.1])
        beta = np.linspace(-0.2, 0.2, 16)
        logp, d_ic, d_beta = FederatedEngine(model, backend="collective").evaluate(ic, beta)
        t_ic = torch.tensor(ic, requires_grad=True)
        t_b = torch.tensor(beta, requires_grad=True)
        eta = X.double() @ t_b + torch.cat([t_ic[0].expand(100), t_ic[1].expand(200)])
        if family == "logistic":
            ll = (y.double() * eta - torch.nn.functional.softplus(eta)).sum()
        elif family == "poisson":
            ll = (y.double() * eta - torch.exp(eta)).sum()
        else:
            ll = (-0.5 * (y.double() - eta) ** 2 - 0.918938533204672742).sum()
        ll.backward()
        np.testing.assert_allclose(logp, ll.item(), rtol=2e-5)
        np.testing.assert_allclose(d_ic, t_ic.grad.numpy(), rtol=2e-4, atol=1e-4)
        np.testing.assert_allclose(d_beta, t_b.grad.numpy(), rtol=2e-4, atol=1e-4)


def test_ode_reference_gradient_matches_finite_differences():
    t, y0, obs, sigma = synth_lv_shard(5, 6, seed=3, device="cpu")
    model = OdeShards([t], [y0], [obs], [sigma], substeps=4)
    th = np.array([0.9, 0.45, 0.75, 0.25])
    # the oracle itself is float64 end to end (theta only becomes float32 in the mailbox)
    logp, grad = model.reference([th])
    eps = 1e-6
    for k in range(4):
        d = np.zeros(4)
        d[k] = eps
        fd = (model.reference([th + d])[0] - model.reference([th - d])[0]) / (2 * eps)
        np.testing.assert_allclose(grad[k], fd, rtol=1e-5, atol=1e-6)
    logp_e, grad_e = FederatedEngine(model, backend="collective").evaluate(th)
    np.testing.assert_allclose(logp_e, logp, rtol=1e-5)
    np.testing.assert_allclose(grad_e, grad, rtol=1e-3, atol=1e-2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, out_queue):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(10 + rank)  # every rank has its own private shard
    x, y = rng.normal(size=50), rng.normal(size=50)
    model = LinregShards([x], [y], [0.8], local_ids=[rank], n_shards_total=world)
    eng = FederatedEngine(model, backend="collective")
    if rank == 0:
        logp, da, db = eng.evaluate(np.array(0.3), np.array(-0.2))
        logp2, *_ = eng.evaluate(np.array(0.0), np.array(0.0))
        eng.shutdown()
        out_queue.put((float(logp), float(da), float(db), float(logp2)))
    else:
        served = eng.serve()
        eng.shutdown()
        out_queue.put(("served", served))
    dist.destroy_process_group()


def _gloo_phased_worker(rank, world, port, out_queue):
    """bench.py's structure: peers serve bounded phases, then everybody shuts down."""
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(10 + rank)
    x, y = rng.normal(size=50), rng.normal(size=50)
    eng = FederatedEngine(LinregShards([x], [y], [0.8], local_ids=[rank], n_shards_total=world), backend="collective")
    if rank == 0:
        for _ in range(3):
            eng.evaluate(np.array(0.3), np.array(-0.2))
        dist.barrier()
        for _ in range(2):
            eng.evaluate(np.array(0.1), np.array(0.2))
        dist.barrier()
    else:
        assert eng.serve(max_epochs=3) == 3
        dist.barrier()
        assert eng.serve(max_epochs=2) == 2
        dist.barrier()
    eng.shutdown()  # must not dead-lock although the peers are not inside serve()
    out_queue.put(("done", rank))
    dist.destroy_process_group()


def test_bounded_serve_phases_then_shutdown_do_not_deadlock():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_phased_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    done = sorted(q.get(timeout=120)[1] for _ in procs)
    assert done == [0, 1, 2]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0


def test_two_rank_gloo_federation_sums_private_shards():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    served = [r for r in results if r[0] == "served"][0]
    root = [r for r in results if r[0] != "served"][0]
    assert served == ("served", 2)
    expected = 0.0
    for rank in range(2):
        rng = np.random.default_rng(10 + rank)
        x, y = rng.normal(size=50), rng.normal(size=50)
        expected += _scipy_linreg(x, y, 0.8, 0.3, -0.2)
    np.testing.assert_allclose(root[0], expected, rtol=1e-12)


def test_fp8_tile_scale_packing_matches_tmem_word_order():
    """pack_tile_scales: words 0-7 feed MMA #1 (row group q, feature blocks 4g..4g+3), words 8-15 MMA #2
    (row groups 0..3 of feature block 4h+qq); absent blocks are 2^0 (csrc/glm_fp8.cu)."""
    import torch

    from pytensor_federated_b200.models.glm import pack_tile_scales

    for P in (128, 256):
        nfb, tiles = P // 32, 3
        s = torch.randint(100, 150, (4 * tiles, nfb), dtype=torch.uint8)
        pk = pack_tile_scales(s, P).view(tiles, 16, 4)
        for t in range(tiles):
            rows = s[4 * t : 4 * t + 4]
            for g in range(2):
                for q in range(4):
                    want = [int(rows[q, 4 * g + j]) if 4 * g < nfb else 127 for j in range(4)]
                    assert pk[t, g * 4 + q].tolist() == want
                    want = [int(rows[j, 4 * g + q]) if 4 * g < nfb else 127 for j in range(4)]
                    assert pk[t, 8 + g * 4 + q].tolist() == want
    with pytest.raises(ValueError):
        pack_tile_scales(torch.zeros(6, 8, dtype=torch.uint8), 256)


@pytest.mark.parametrize("family", ["logistic", "poisson", "gaussian"])
def test_glm_shards_agree_with_the_same_model_written_in_the_graph_ir(family):
    """Two independent implementations of the federated GLM: ``GlmShards`` (eager PyTorch partials summed by
    the engine) and the likelihood written with ``at.dot`` in the graph IR, differentiated symbolically."""
    from pytensor_federated_b200._graph_backend import BACKEND, at, function, grad

    if BACKEND != "builtin":
        pytest.skip("written against the built-in IR's op names")
    rng = np.random.default_rng(3)
    P, groups = 5, [0, 1, 1]
    Xs = [rng.normal(size=(n, P)) for n in (40, 25, 31)]
    beta_true = rng.normal(size=P) * 0.3
    ys = []
    for X in Xs:
        eta = X @ beta_true
        ys.append({"logistic": (rng.random(X.shape[0]) < 1 / (1 + np.exp(-eta))).astype(float),
                   "poisson": rng.poisson(np.exp(eta)).astype(float),
                   "gaussian": eta + rng.normal(size=X.shape[0])}[family])
    model = GlmShards([torch.tensor(X) for X in Xs], [torch.tensor(y) for y in ys], groups=groups, n_groups=2, family=family)
    engine = FederatedEngine(model, backend="collective")
    ic_v, beta_v = np.array([0.2, -0.1]), rng.normal(size=P) * 0.2
    logp, d_ic, d_beta = engine.evaluate(ic_v, beta_v)
    engine.shutdown()

    ic, beta = at.vector("ic"), at.vector("beta")
    total = None
    for X, y, g in zip(Xs, ys, groups):
        eta = at.dot(at.as_tensor(X), beta) + ic[g]
        yv = at.as_tensor(y)
        if family == "logistic":
            ll = yv * eta - at.softplus(eta)
        elif family == "poisson":
            ll = yv * eta - at.exp(eta)
        else:
            ll = -0.5 * at.sqr(yv - eta) - 0.918938533204672742
        total = ll.sum() if total is None else total + ll.sum()
    f = function([ic, beta], [total] + list(grad(total, [ic, beta])))
    want = f(ic_v, beta_v)
    np.testing.assert_allclose(logp, want[0], rtol=1e-5)
    np.testing.assert_allclose(d_ic, want[1], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(d_beta, want[2], rtol=1e-4, atol=1e-4)


def test_fp8_shards_on_the_cpu_oracle_path_equal_dense_shards_of_the_dequantised_matrix():
    from pytensor_federated_b200.models import Fp8GlmShards, dequantize_block_fp8

    torch.manual_seed(9)
    Xs = [torch.randn(n, 128) * torch.exp(torch.randn(1, 128)) for n in (200, 130)]
    ys = [(torch.rand(n) < 0.4).float() for n in (200, 130)]
    fp8 = Fp8GlmShards.from_dense(Xs, ys, groups=[0, 1], n_groups=2, family="poisson")
    assert [tuple(s.shape) for s in fp8._kernel_scales] == [(2, 64), (2, 64)]       # 16 words per 128-row tile
    dense = GlmShards([dequantize_block_fp8(X, s) for X, s in zip(fp8.Xs, fp8.scales)], ys, groups=[0, 1], n_groups=2,
                      family="poisson")
    ic, beta = np.array([0.1, -0.2]), (np.random.default_rng(2).normal(size=128) * 0.01)
    with FederatedEngine(fp8, backend="collective") as a, FederatedEngine(dense, backend="collective") as b:
        for u, v in zip(a.evaluate(ic, beta), b.evaluate(ic, beta)):
            np.testing.assert_allclose(u, v, rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        Fp8GlmShards.from_dense(Xs, ys, n_chains=4)


def test_custom_family_compiles_for_sm100a_and_has_a_cpu_oracle_path(tmp_path):
    """The user's CUDA snippet is cross-compiled here (no GPU needed); the eager oracle runs on the CPU."""
    import shutil

    from pytensor_federated_b200.models import CustomFamily

    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("nvcc not available")
    student_t = CustomFamily(
        "const float d = y - eta; ll = -2.5f * log1pf(d * d * 0.25f); r = 5.f * d / (4.f + d * d);",
        torch_fn=lambda y, eta: (-2.5 * torch.log1p((y - eta) ** 2 / 4), 5 * (y - eta) / (4 + (y - eta) ** 2)),
    )
    assert student_t.launcher_address() != 0                      # nvcc -> .so -> symbol
    assert student_t.compile() is student_t.compile()             # cached
    with pytest.raises(RuntimeError, match="nvcc rejected"):
        CustomFamily("ll = this is not CUDA;").compile()

    rng = np.random.default_rng(0)
    X = torch.tensor(rng.normal(size=(60, 4)), dtype=torch.float32)
    y = torch.tensor(rng.standard_t(4, size=60), dtype=torch.float32)
    model = GlmShards([X], [y], family=student_t)
    with FederatedEngine(model, backend="collective") as eng:
        beta = rng.normal(size=4) * 0.1
        logp, d_ic, d_beta = eng.evaluate(np.array([0.05]), beta)
        eps = 1e-3
        for j in range(4):
            bp, bm = beta.copy(), beta.copy()
            bp[j] += eps
            bm[j] -= eps
            fd = (eng.evaluate(np.array([0.05]), bp)[0] - eng.evaluate(np.array([0.05]), bm)[0]) / (2 * eps)
            np.testing.assert_allclose(d_beta[j], fd, rtol=2e-3, atol=2e-3)


def test_user_defined_ode_system_oracle_matches_scipy_and_compiles_for_sm100a():
    """A user-supplied right-hand side (SIR epidemic): the eager RK4 oracle agrees with SciPy's adaptive
    solver, its autograd gradient with finite differences, and the CUDA snippet cross-compiles into the fused
    kernel (the GPU test runs it)."""
    import scipy.integrate

    from pytensor_federated_b200.models import LOTKA_VOLTERRA, OdeSystem, synth_ode_shard

    sir = OdeSystem(
        "const auto inf = th[0] * y[0] * y[1]; dy[0] = -inf; dy[1] = inf - th[1] * y[1]; dy[2] = th[1] * y[1];",
        lambda y, th, t: (-th[0] * y[0] * y[1], th[0] * y[0] * y[1] - th[1] * y[1], th[1] * y[1]),
        n_states=3, n_params=2, name="sir",
    )
    theta = np.array([1.8, 0.5])
    y0 = np.array([[0.95, 0.9, 0.8], [0.05, 0.1, 0.2], [0.0, 0.0, 0.0]])
    t, y0_t, obs, sigma = synth_ode_shard(sir, theta, y0, 12, seed=4, device="cpu", sigma=0.02, t_end=6.0, substeps=16)
    model = OdeShards([t], [y0_t], [obs], [sigma], substeps=16, system=sir)
    assert model.n_theta_words == 2 and model.n_vals == 3
    logp, grad = FederatedEngine(model, backend="collective").evaluate(theta)
    # SciPy oracle of the log-likelihood (adaptive RK45, tight tolerances)
    def scipy_logp(th):
        total = 0.0
        for i in range(y0.shape[1]):
            sol = scipy.integrate.solve_ivp(lambda tt, y: [-th[0] * y[0] * y[1], th[0] * y[0] * y[1] - th[1] * y[1], th[1] * y[1]],
                                            (0.0, float(t[-1])), y0[:, i], t_eval=t.numpy().astype(np.float64), rtol=1e-10, atol=1e-12)
            r = obs[:, :, i].numpy().astype(np.float64) - sol.y.T
            total += float(np.sum(-0.5 * r * r / sigma**2 - np.log(sigma) - 0.918938533204672742))
        return total
    np.testing.assert_allclose(logp, scipy_logp(theta), rtol=2e-5)
    eps = 1e-5
    fd = [(scipy_logp(theta + eps * np.eye(2)[k]) - scipy_logp(theta - eps * np.eye(2)[k])) / (2 * eps) for k in range(2)]
    np.testing.assert_allclose(grad, fd, rtol=2e-4, atol=1e-3)
    # the snippets compile into the fused kernel (cross-compilation needs no GPU)
    assert sir.compile().b200_launch_ode_custom is not None
    ns, npar = C.c_int(), C.c_int()
    sir.compile().b200_ode_generic_dims(C.byref(ns), C.byref(npar))
    assert (ns.value, npar.value) == (3, 2)
    with pytest.raises(RuntimeError, match="nvcc rejected"):
        OdeSystem("dy[0] = undefined_symbol;", None, n_states=1, n_params=1).compile()
    assert LOTKA_VOLTERRA.n_params == 4
