"""The built-in graph IR: autodiff of every primitive vs finite differences, merge, FunctionGraph."""
import numpy as np
import pytest

from pytensor_federated_b200.graph import core as G


def _fd(f, x, eps=1e-6):
    x = np.asarray(x, dtype=np.float64)
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        d = np.zeros_like(x)
        d[it.multi_index] = eps
        g[it.multi_index] = (f(x + d) - f(x - d)) / (2 * eps)
    return g


CASES = {
    "add-mul-broadcast": lambda v, s: ((v * s + v) * 2.0).sum(),
    "sub-div": lambda v, s: ((v - s) / (v * v + 1.0)).sum(),
    "exp-log": lambda v, s: (G.exp(v * 0.3) + G.log(v * v + 1.5) * s).sum(),
    "sqr-sqrt-pow": lambda v, s: (G.sqrt(v * v + 2.0) + (v ** 2) * s + (v * v + 1.0) ** 1.5).sum(),
    "sigmoid-softplus": lambda v, s: (G.sigmoid(v) * s + G.softplus(v - s)).sum(),
    "neg-index-stack": lambda v, s: (-v[0] * v[2] + G.stack([v[1], s, v[0]]).sum() * s),
    "sum-axis": lambda v, s: ((G.as_tensor(np.ones((2, 3))) * v).sum(axis=0) * v).sum() * s,
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_gradients_match_finite_differences(name):
    v, s = G.vector("v"), G.scalar("s")
    cost = CASES[name](v, s)
    gv, gs = G.grad(cost, [v, s])
    fn = G.function([v, s], [cost, gv, gs])
    v0, s0 = np.array([0.3, -1.2, 2.0]), 0.7
    c, dv, ds = fn(v0, s0)
    f_cost = G.function([v, s], cost, mode="FAST_COMPILE")
    np.testing.assert_allclose(dv, _fd(lambda x: float(f_cost(x, s0)), v0), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ds, _fd(lambda x: float(f_cost(v0, float(x))), np.array(s0)), rtol=1e-5, atol=1e-7)
    assert np.isfinite(c)


def test_grad_of_unused_input_is_zero_and_disconnected_outputs_are_flagged():
    a, b = G.scalar("a"), G.scalar("b")
    g = G.grad(a * a, [a, b])
    fn = G.function([a, b], g)
    assert [float(x) for x in fn(3.0, 5.0)] == [6.0, 0.0]
    with pytest.raises(ValueError):
        G.grad(a * a, [b], disconnected_inputs="raise")
    with pytest.raises(TypeError):
        G.grad(G.vector() * 2.0, [a])  # non-scalar cost


def test_merge_pass_deduplicates_and_function_graph_validates_inputs():
    a = G.scalar("a")
    e1 = G.exp(a * 2.0)
    e2 = G.exp(a * 2.0)
    fn = G.function([a], e1 + e2)
    ops = [type(n.op).__name__ + ":" + getattr(n.op, "name", "") for n in fn.maker.fgraph.toposort()]
    assert ops.count("Elemwise:exp") == 1  # merged
    assert float(fn(0.5)) == pytest.approx(2 * np.exp(1.0))
    b = G.scalar("b")
    with pytest.raises(ValueError):
        G.function([a], a + b)  # b is not an input
    with pytest.raises(TypeError):
        fn(np.array([1.0, 2.0]))  # wrong ndim
    with pytest.raises(TypeError):
        list(a)  # symbolic variables are not iterable


def test_replace_validate_rolls_back_on_failure():
    a, b = G.scalar("a"), G.scalar("b")
    out = a * 2.0
    fg = G.FunctionGraph([a], [out], clone=True)
    fg.attach_feature(G.ReplaceValidate())
    before = [n.op for n in fg.toposort()]
    with pytest.raises(ValueError):
        fg.replace_all_validate([(fg.outputs[0].owner.inputs[0], b)])  # b is foreign -> invalid graph
    assert [n.op for n in fg.toposort()] == before


def test_dot_transpose_and_more_elementwise_gradients():
    """dot (all vector/matrix combinations), transpose, tanh, log1p, abs, maximum, mean vs finite differences."""
    from pytensor_federated_b200.graph import core as at

    rng = np.random.default_rng(0)

    def num_grad(fn, v, eps=1e-6):
        g = np.zeros_like(v)
        for idx in np.ndindex(*v.shape):
            vp, vm = v.copy(), v.copy()
            vp[idx] += eps
            vm[idx] -= eps
            g[idx] = (fn(vp) - fn(vm)) / (2 * eps)
        return g

    A, B, x, b = at.matrix("A"), at.matrix("B"), at.vector("x"), at.vector("b")
    cost = (at.sum(at.tanh(at.dot(A, x) + b)) + at.mean(at.log1p(at.abs(x))) + at.sum(at.maximum(x, 0.1))
            + at.sum(at.sqr(at.dot(A, B))) + at.sum(at.sqr(at.dot(x, B))) + at.dot(x, x) + at.sum(at.transpose(A) * 0.5))
    f = at.function([A, B, x, b], [cost] + list(at.grad(cost, [A, B, x, b])))
    vals = [rng.normal(size=(3, 2)), rng.normal(size=(2, 4)), rng.normal(size=2), rng.normal(size=3)]
    out = f(*vals)
    for i in range(4):
        def scalar(v, i=i):
            args = list(vals)
            args[i] = v
            return float(f(*args)[0])

        np.testing.assert_allclose(out[1 + i], num_grad(scalar, vals[i]), atol=2e-5)
    g = at.function([x], [at.mean(x), at.mean(at.stack([x, x]), axis=0)])
    m, m0 = g(np.array([1.0, 3.0]))
    assert m == 2.0 and m0.tolist() == [1.0, 3.0]
    with pytest.raises(TypeError):
        at.dot(at.scalar("s"), x)


def test_identity_eliminator_removes_autodiff_noise_without_touching_values_shapes_or_dtypes():
    """`g + zeros_like(scalar)`, `1.0 * g`, `g + 0.0` disappear in FAST_RUN; anything that could broadcast or
    change the dtype stays."""
    v, s = G.vector("v"), G.scalar("s")
    cost = (v * s).sum() + s * s + (v[1] + 2.0) * s
    grads = G.grad(cost, [v, s])
    fast = G.function([v, s], [cost, *grads])
    slow = G.function([v, s], [cost, *grads], mode=G.Mode("FAST_RUN").excluding("identities"))
    n_fast, n_slow = len(fast.maker.fgraph.toposort()), len(slow.maker.fgraph.toposort())
    assert n_fast < n_slow
    for args in ((np.array([0.3, -1.2, 2.0]), 0.7), (np.array([1.0, 2.0]), -3.0)):
        for a, b in zip(fast(*args), slow(*args)):
            assert isinstance(a, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape
            np.testing.assert_array_equal(a, b)

    # not an identity: a zero VECTOR broadcasts the scalar, a float constant promotes an integer operand
    out = s + G.zeros_like(v)
    f = G.function([v, s], out)
    assert any(isinstance(n.op, G.ZerosLike) for n in f.maker.fgraph.toposort())
    np.testing.assert_array_equal(f(np.zeros(3), 2.0), [2.0, 2.0, 2.0])
    i = G.scalar("i", dtype="int64")
    g = G.function([i], i * 1.0)
    assert g(np.int64(3)).dtype == np.float64 and len(g.maker.fgraph.toposort()) == 1


def test_compiled_function_runs_scalars_as_numpy_scalars_but_hands_out_arrays():
    """Statically scalar arithmetic runs on NumPy scalars inside the generated code; Ops that go through
    `perform` (user functions, federated Ops) and the caller still see ndarrays, and the generated function agrees
    with the interpreter loop it replaces."""
    seen = []

    def spy(a, b):
        seen.append((type(a), type(b), np.shape(a), np.shape(b)))
        return np.asarray(a * b)

    op = G.FromFunctionOp(spy, [G.TensorType("float64", ()), G.TensorType("float64", (None,))], [G.TensorType("float64", (None,))])
    s, v = G.scalar("s"), G.vector("v")
    t = (s * 2.0 + 1.0) / 3.0                 # scalar thunks
    out = op(t, v * t)                        # perform path fed by a scalar
    total = out.sum() + t
    fn = G.function([s, v], [total, t, out])
    assert fn._run is not None
    res = fn(1.5, np.array([1.0, 2.0]))
    assert all(isinstance(r, np.ndarray) for r in res) and res[0].shape == () and res[1].shape == ()
    assert seen == [(np.ndarray, np.ndarray, (), (2,))]
    tv = (1.5 * 2.0 + 1.0) / 3.0
    np.testing.assert_allclose(res[2], tv * np.array([1.0, 2.0]) * tv)
    np.testing.assert_allclose(res[0], res[2].sum() + tv)
    # the interpreter fallback computes the same thing
    generated, fn._run = fn._run, None
    again = fn(1.5, np.array([1.0, 2.0]))
    fn._run = generated
    for a, b in zip(res, again):
        assert isinstance(b, np.ndarray) and a.shape == b.shape and a.dtype == b.dtype
        np.testing.assert_array_equal(a, b)
    with pytest.raises(TypeError):
        fn(1.5)


def test_compiled_function_edge_cases_inputs_as_outputs_constants_and_mixed_dtypes():
    s, v = G.scalar("s"), G.vector("v")
    f = G.function([s, v], [s, v, s * 2.0, G.as_tensor(3.0), s * 2.0])
    out = f(1.5, np.array([1.0, 2.0]))
    assert [np.asarray(o).tolist() for o in out] == [1.5, [1.0, 2.0], 3.0, 3.0, 3.0]
    assert isinstance(G.function([s], s)(2.0), np.ndarray)
    i = G.scalar("i", dtype="int64")
    r = G.function([i, s], [i + 1, i / 2, i * s, -i])(3, 0.5)
    assert [o.dtype.kind for o in r] == ["i", "f", "f", "i"] and [o.tolist() for o in r] == [4, 1.5, 1.5, -3]
    x = G.scalar("x", dtype="float32")
    r = G.function([x, s], [x * x, x + s])(np.float32(1.5), 0.25)
    assert r[0].dtype == np.float32 and r[1].dtype == np.float64 and r[1] == 1.75


def _random_expression(rng, leaves, depth):
    """A random expression over scalar and vector leaves (broadcasting, reductions, constants that look like
    identities) — food for the differential test below."""
    if depth == 0 or rng.random() < 0.15:
        leaf = leaves[rng.integers(len(leaves))]
        return leaf if rng.random() < 0.8 else G.as_tensor(float(rng.choice([0.0, 1.0, 2.0, -0.5])))
    kind = rng.integers(9)
    a = _random_expression(rng, leaves, depth - 1)
    if kind in (0, 1, 2, 3):
        b = _random_expression(rng, leaves, depth - 1)
        return [G.add, G.sub, G.mul, lambda x, y: x / (y * y + 1.5)][kind](a, b)
    if kind == 4:
        return -a
    if kind == 5:
        return G.tanh(a)
    if kind == 6:
        return G.exp(a * 0.1)
    if kind == 7:
        return a.sum() if a.type.ndim else a * 1.0
    return a + G.zeros_like(leaves[rng.integers(len(leaves))]) if rng.random() < 0.5 else a + 0.0


@pytest.mark.parametrize("seed", range(25))
def test_generated_code_identities_and_scalar_paths_agree_with_the_plain_interpreter(seed):
    """Differential test of the compiled Function: FAST_RUN (identity elimination, generated straight-line code,
    NumPy-scalar arithmetic) against the interpreter loop on the unsimplified graph, values AND gradients."""
    rng = np.random.default_rng(seed)
    s, t, v = G.scalar("s"), G.scalar("t"), G.vector("v")
    # elements of v used one by one: their gradients are IncSubtensor terms (merged into one ScatterAdd)
    expr = _random_expression(rng, [s, t, v, v[0], v[2], v[0] * v[1], v[0:3]], depth=5)
    cost = expr.sum() if expr.type.ndim else expr
    outputs = [cost, expr, *G.grad(cost, [s, t, v])]
    fast = G.function([s, t, v], outputs)
    plain = G.function([s, t, v], outputs, mode=G.Mode("FAST_RUN").excluding("identities", "merge_inc_subtensor"))
    plain._run = None                                    # the interpreter loop
    args = (0.7, -1.3, np.array([0.2, -0.4, 1.1]))
    for a, b in zip(fast(*args), plain(*args)):
        assert isinstance(a, np.ndarray) and a.shape == b.shape and a.dtype == b.dtype
        np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-13)


def test_inc_subtensor_terms_are_merged_into_one_scatter_add():
    """d/dv of sum_k f_k(v[k]): one ScatterAdd instead of one zeros vector + one vector add per element."""
    v, s = G.vector("v"), G.scalar("s")
    cost = G.exp(v[0] * s) + v[1] * v[1] + G.tanh(v[2]) * s + v[1] * 3.0 + (v * v).sum()
    gv, gs = G.grad(cost, [v, s])
    fast = G.function([v, s], [cost, gv, gs])
    slow = G.function([v, s], [cost, gv, gs], mode=G.Mode("FAST_RUN").excluding("merge_inc_subtensor"))
    kinds = [type(n.op).__name__ for n in fast.maker.fgraph.toposort()]
    assert kinds.count("ScatterAdd") == 1 and "IncSubtensorZeros" not in kinds
    assert len(fast.maker.fgraph.toposort()) < len(slow.maker.fgraph.toposort())
    args = (np.array([0.3, -1.2, 2.0, 0.5]), 0.7)
    for a, b in zip(fast(*args), slow(*args)):
        np.testing.assert_allclose(a, b, rtol=1e-14)
    # ScatterAdd is differentiable itself (second use of the gradient graph)
    g2 = G.grad(gv.sum(), v)
    np.testing.assert_allclose(G.function([v, s], g2)(*args), G.function([v, s], g2, mode="FAST_COMPILE")(*args), rtol=1e-12)
