"""AsyncOp / ParallelAsyncOp / fuse_asyncs — timing and structure (CPU).

Strategy of the reference's ``test_op_async.py`` (cooperative sleeps + wall-clock windows,
structural assertions on the rewritten graph), run on whichever graph backend is active.
"""
import asyncio
import time

import numpy as np
import pytest

from pytensor_federated_b200 import op_async
from pytensor_federated_b200._graph_backend import BACKEND, FunctionGraph, at, function


class _AsyncDelay(op_async.AsyncOp):
    """Passes its scalar input through after sleeping cooperatively."""

    def __init__(self, delay: float, fail: bool = False) -> None:
        self.delay = delay
        self.fail = fail
        super().__init__()

    def make_node(self, x):
        x = at.as_tensor(x)
        from pytensor_federated_b200._graph_backend import Apply

        return Apply(self, [x], [x.type()])

    async def perform_async(self, node, inputs, output_storage):
        await asyncio.sleep(self.delay)
        if self.fail:
            raise RuntimeError("remote node exploded")
        output_storage[0][0] = np.asarray(inputs[0])


def _timed(fn, *args):
    t0 = time.perf_counter()
    out = fn(*args)
    return out, time.perf_counter() - t0


def test_async_op_perform_blocks_on_the_coroutine():
    op = _AsyncDelay(0.3)
    x = at.scalar()
    node = op.make_node(x)
    storage = [[None]]
    _, dt = _timed(op.perform, node, [np.array(1.5)], storage)
    assert storage[0][0] == 1.5
    assert 0.3 <= dt < 0.6


def test_async_from_function_op():
    async def double(a):
        await asyncio.sleep(0.2)
        return 2 * a

    op = op_async.AsyncFromFunctionOp(double, [at.scalar().type], [at.scalar().type])
    x = at.scalar()
    out = op(x)
    (res), dt = _timed(lambda: out.eval({x: 4.0}))
    assert res == 8.0
    assert 0.2 <= dt < 0.5


def test_parallel_async_op_validation_and_concurrency():
    x, y = at.scalar(), at.scalar()
    a1 = _AsyncDelay(0.4).make_node(x)
    a2 = _AsyncDelay(0.3).make_node(y)
    with pytest.raises(ValueError, match="not an `AsyncOp`"):
        op_async.ParallelAsyncOp([a1, (x + y).owner])
    pop = op_async.ParallelAsyncOp([a1, a2])
    with pytest.raises(ValueError, match="expected 2 for 2"):
        pop.make_node(x)
    node = pop.make_node(x, y)
    assert len(node.outputs) == 2
    storage = [[None], [None]]
    _, dt = _timed(pop.perform, node, [np.array(1.0), np.array(2.0)], storage)
    assert [float(s[0]) for s in storage] == [1.0, 2.0]
    assert 0.4 <= dt < 0.65  # max(0.4, 0.3) = 0.4, not the sum 0.7


def test_parallel_async_op_reraises_child_errors():
    """The reference swallows these (return_exceptions=True, results never inspected)."""
    x, y = at.scalar(), at.scalar()
    pop = op_async.ParallelAsyncOp([_AsyncDelay(0.05).make_node(x), _AsyncDelay(0.01, fail=True).make_node(y)])
    node = pop.make_node(x, y)
    with pytest.raises(RuntimeError, match="exploded"):
        pop.perform(node, [np.array(1.0), np.array(2.0)], [[None], [None]])


def test_find_and_fuse_independent_applies():
    a, b = at.scalar("a"), at.scalar("b")
    d1 = _AsyncDelay(0.01)(a)
    d2 = _AsyncDelay(0.01)(b)
    d3 = _AsyncDelay(0.01)(a + b)
    total = d1 + d2 + d3
    fg = FunctionGraph([a, b], [total], clone=True)
    found = op_async.find_parallelizable_applies(fg, op_async.AsyncOp)
    assert len(found) == 3
    # other op classes are not collected
    assert op_async.find_parallelizable_applies(fg, op_async.ParallelAsyncOp) == []
    op_async.parallelize_async_applies(fg, found)
    owners = {id(n) for n in fg.toposort() if isinstance(n.op, op_async.ParallelAsyncOp)}
    assert len(owners) == 1
    assert not any(isinstance(n.op, _AsyncDelay) for n in fg.toposort())


def test_layered_fusion_follows_dependencies():
    a = at.scalar("a")
    l1a = _AsyncDelay(0.01)(a)
    l1b = _AsyncDelay(0.01)(a * 2.0)
    l2a = _AsyncDelay(0.01)(l1a + l1b)
    l2b = _AsyncDelay(0.01)(l1a - l1b)
    fg = FunctionGraph([a], [l2a + l2b], clone=True)
    op_async.parallelize_all_async_applies(fg)
    fused = [n for n in fg.toposort() if isinstance(n.op, op_async.ParallelAsyncOp)]
    assert len(fused) == 2  # one per dependency level
    assert all(len(n.op.applies) == 2 for n in fused)


def test_compile_modes_sequential_vs_fused_timing():
    a = at.scalar("a")
    l1a = _AsyncDelay(0.3)(a)
    l1b = _AsyncDelay(0.2)(a * 2.0)
    l2a = _AsyncDelay(0.3)(l1a + l1b)
    l2b = _AsyncDelay(0.2)(l1a - l1b)
    out = l2a + l2b
    slow = function([a], out, mode="FAST_COMPILE")
    fast = function([a], out)  # default mode runs fuse_asyncs
    r1, t_slow = _timed(slow, 1.0)
    r2, t_fast = _timed(fast, 1.0)
    assert float(r1) == float(r2) == (1.0 + 2.0) + (1.0 - 2.0)
    assert 1.0 <= t_slow < 1.5  # 0.3 + 0.2 + 0.3 + 0.2
    assert 0.6 <= t_fast < 0.95  # max(0.3, 0.2) per level; sequential would be 1.0


def test_fuse_asyncs_is_registered_once():
    from pytensor_federated_b200._graph_backend import optdb

    assert "fuse_asyncs" in optdb
    assert isinstance(optdb["fuse_asyncs"], op_async.AsyncFusionOptimizer) or BACKEND == "pytensor"
    if BACKEND == "builtin":
        # the module guards its registration; a second unguarded one must be rejected
        with pytest.raises(ValueError, match="already registered"):
            optdb.register("fuse_asyncs", op_async.AsyncFusionOptimizer(), "fast_run", position=90)
        from pytensor_federated_b200.graph.core import Mode

        a = at.scalar()
        out = _AsyncDelay(0.0)(a) + _AsyncDelay(0.0)(a * 2.0)
        unfused = function([a], out, mode=Mode("FAST_RUN").excluding("fuse_asyncs"))
        assert not any(isinstance(n.op, op_async.ParallelAsyncOp) for n in unfused.maker.fgraph.toposort())
