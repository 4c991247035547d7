"""LogpOp / LogpGradOp / ArraysToArraysOp and async twins with a mock client (CPU)."""
import asyncio

import numpy as np
import pytest

from pytensor_federated_b200 import (
    ArraysToArraysOp,
    AsyncArraysToArraysOp,
    AsyncLogpGradOp,
    AsyncLogpOp,
    LogpGradOp,
    LogpOp,
    op_async,
)
from pytensor_federated_b200._graph_backend import at, function, grad


class _MockClient:
    """Quadratic log-potential with manual gradients; counts calls."""

    def __init__(self, data=(1.0, 2.0, 3.0)) -> None:
        self.data = np.asarray(data)
        self.calls = 0

    def logp(self, a, b):
        self.calls += 1
        return np.asarray(-np.sum((self.data - a - b * self.data) ** 2))

    def logp_grad(self, a, b):
        self.calls += 1
        r = self.data - a - b * self.data
        return np.asarray(-np.sum(r**2)), [np.asarray(2 * np.sum(r)), np.asarray(2 * np.sum(r * self.data))]

    async def logp_async(self, a, b):
        await asyncio.sleep(0.01)
        return self.logp(a, b)

    async def logp_grad_async(self, a, b):
        await asyncio.sleep(0.01)
        return self.logp_grad(a, b)


def test_logp_op_make_node_and_perform():
    client = _MockClient()
    op = LogpOp(client.logp)
    a, b = at.scalar(), at.scalar()
    node = op.make_node(a, b)
    assert len(node.inputs) == 2 and len(node.outputs) == 1 and node.outputs[0].type.ndim == 0
    # plain numbers are accepted (reference issue #24)
    assert len(op.make_node(1.0, np.array(2.0)).inputs) == 2
    storage = [[None]]
    op.perform(node, [np.array(0.5), np.array(0.1)], storage)
    assert storage[0][0] == client.logp(0.5, 0.1)
    assert op(a, b).eval({a: 0.5, b: 0.1}) == storage[0][0]


def test_logp_grad_op_make_node_perform_and_eval():
    client = _MockClient()
    op = LogpGradOp(client.logp_grad)
    a, v = at.scalar(), at.vector()
    node = op.make_node(a, v)
    assert len(node.outputs) == 3
    assert node.outputs[1].type == a.type and node.outputs[2].type == v.type
    assert len(op.make_node(0.3, 0.4).outputs) == 3
    a, b = at.scalar(), at.scalar()
    node = op.make_node(a, b)
    storage = [[None], [None], [None]]
    op.perform(node, [np.array(0.5), np.array(0.1)], storage)
    want = client.logp_grad(0.5, 0.1)
    assert storage[0][0] == want[0] and storage[1][0] == want[1][0] and storage[2][0] == want[1][1]
    logp, da, db = op(a, b)
    np.testing.assert_allclose(da.eval({a: 0.5, b: 0.1}), want[1][0])


def test_grad_uses_the_federated_gradient_with_one_call():
    client = _MockClient()
    op = LogpGradOp(client.logp_grad)
    a, b = at.scalar("a"), at.scalar("b")
    logp, *_ = op(a, b)
    cost = 3.0 * logp + a * a
    ga, gb = grad(cost, [a, b])
    fn = function([a, b], [cost, ga, gb])
    client.calls = 0
    c, da, db = fn(0.5, 0.1)
    assert client.calls == 1  # forward and gradient share ONE remote evaluation (merge pass)
    w_logp, (w_da, w_db) = client.logp_grad(0.5, 0.1)
    np.testing.assert_allclose(c, 3 * w_logp + 0.25)
    np.testing.assert_allclose(da, 3 * w_da + 1.0)
    np.testing.assert_allclose(db, 3 * w_db)


def test_grad_refuses_second_derivatives():
    op = LogpGradOp(_MockClient().logp_grad)
    a, b = at.scalar(), at.scalar()
    logp, da, db = op(a, b)
    with pytest.raises(ValueError, match="Can't propagate gradients wrt parameter 1"):
        grad(logp + da, [a])


def test_async_ops_fuse_and_match_sync_results():
    client = _MockClient()
    a, b = at.scalar("a"), at.scalar("b")
    sync_total = LogpGradOp(client.logp_grad)(a, b)[0] + LogpGradOp(client.logp_grad)(a + 1.0, b)[0]
    aop = AsyncLogpGradOp(client.logp_grad_async)
    async_total = aop(a, b)[0] + aop(a + 1.0, b)[0]
    f_sync = function([a, b], [sync_total, *grad(sync_total, [a, b])])
    f_async = function([a, b], [async_total, *grad(async_total, [a, b])])
    kinds = [type(n.op).__name__ for n in f_async.maker.fgraph.toposort()]
    assert kinds.count("ParallelAsyncOp") == 1 and "AsyncLogpGradOp" not in kinds
    np.testing.assert_allclose(f_sync(0.2, 0.3), f_async(0.2, 0.3))
    lop = AsyncLogpOp(client.logp_async)
    np.testing.assert_allclose(lop(a, b).eval({a: 0.2, b: 0.3}), client.logp(0.2, 0.3))


def test_arrays_to_arrays_ops():
    def compute(x, y):
        return [x + y, x * y]

    async def compute_async(x, y):
        await asyncio.sleep(0.01)
        return [x + y, x * y]

    vec = at.vector().type
    for cls, fn in ((ArraysToArraysOp, compute), (AsyncArraysToArraysOp, compute_async)):
        op = cls(fn, [vec, vec], [vec, vec])
        x = at.vector()
        s, p = op(x, np.array([1.0, 2.0]))  # non-Variable input is coerced
        np.testing.assert_allclose(s.eval({x: [3.0, 4.0]}), [4.0, 6.0])
        np.testing.assert_allclose(p.eval({x: [3.0, 4.0]}), [3.0, 8.0])
        with pytest.raises(ValueError):
            op(x)
    assert issubclass(AsyncArraysToArraysOp, op_async.AsyncOp)
