"""Asynchronous Ops and the graph rewrite that runs independent ones concurrently.

Capabilities of ``/root/reference/pytensor_federated/op_async.py``:

* ``AsyncOp`` — an Op whose computation is a coroutine (``perform_async``); the synchronous
  ``perform`` drives it on a re-entrant event loop (``:16-34``).
* ``AsyncFromFunctionOp`` — wraps ``async def fn(*arrays)`` (``:37-65``).
* ``ParallelAsyncOp`` — one node that awaits the coroutines of several child applies at once
  (``:68-132``).
* ``find_parallelizable_applies`` / ``parallelize_async_applies`` /
  ``parallelize_all_async_applies`` and the ``fuse_asyncs`` rewriter registered for
  ``fast_run`` at position 90 (``:135-234``).

Differences on purpose:

* Exceptions raised by a child are **re-raised** after all children finished (the reference
  gathers with ``return_exceptions=True`` and never looks at the results, ``:127-130``, so a
  failed remote call silently leaves stale values in the output storage).
* Children that share a *fusable engine* (a GPU federation) are evaluated by one fused launch
  instead of N awaits — see :class:`FusableAsyncOp`.
"""
from __future__ import annotations

import asyncio
from typing import Any, Callable, List, Optional, Sequence

from ._graph_backend import (
    Apply,
    FromFunctionOp,
    FunctionGraph,
    GraphRewriter,
    Op,
    ReplaceValidate,
    Variable,
    apply_depends_on,
    optdb,
)
from .utils import get_useful_event_loop

OutputStorageType = List[List[Optional[Any]]]


class AsyncOp(Op):
    """Base class: implement :meth:`perform_async`; ``perform`` blocks on it."""

    def perform(self, node: Apply, inputs: Sequence[Any], output_storage: OutputStorageType) -> None:
        loop = get_useful_event_loop()
        loop.run_until_complete(self.perform_async(node, inputs, output_storage))

    async def perform_async(self, node: Apply, inputs: Sequence[Any], output_storage: OutputStorageType) -> None:
        raise NotImplementedError()


class AsyncFromFunctionOp(AsyncOp, FromFunctionOp):
    """Async twin of ``FromFunctionOp``: ``fn`` is a coroutine function of arrays.

    ``AsyncOp.perform`` wins over ``FromFunctionOp.perform`` by MRO.
    """

    def __init__(self, fn: Callable, itypes: Sequence, otypes: Sequence, infer_shape: Optional[Callable] = None):
        self._async_fn = fn
        super().__init__(fn, itypes, otypes, infer_shape)

    async def perform_async(self, node: Apply, inputs: Sequence[Any], output_storage: OutputStorageType) -> None:
        outs = await self._async_fn(*inputs)
        if not isinstance(outs, (list, tuple)):
            outs = (outs,)
        if len(outs) != len(output_storage):
            raise ValueError(f"Function returned {len(outs)} outputs, the Op declares {len(output_storage)}.")
        for cell, value in zip(output_storage, outs):
            cell[0] = value


class ParallelAsyncOp(AsyncOp):
    """Runs the ``perform_async`` of several ``AsyncOp`` apply nodes concurrently.

    Inputs/outputs are the concatenation of the children's inputs/outputs, in order.
    """

    def __init__(self, applies: Sequence[Apply]) -> None:
        applies = tuple(applies)
        for a, apply in enumerate(applies):
            if not isinstance(apply.op, AsyncOp):
                raise ValueError(
                    f"The owner of apply node {a} is not an `AsyncOp`. "
                    "All apply nodes given to `ParallelAsyncOp` must be owned by an `AsyncOp`."
                )
        self.applies = applies
        super().__init__()

    def make_node(self, *inputs: Variable) -> Apply:
        nin_exp = sum(a.nin for a in self.applies)
        if len(inputs) != nin_exp:
            raise ValueError(
                f"Unexpected number of inputs to `ParallelAsyncOp` {self}. "
                f"Got {len(inputs)} inputs but expected {nin_exp} for {len(self.applies)} apply nodes."
            )
        outputs = [out.type() for app in self.applies for out in app.outputs]
        return Apply(self, list(inputs), outputs)

    def _slices(self, inputs, output_storage):
        ifrom = ofrom = 0
        for apply in self.applies:
            yield apply, inputs[ifrom : ifrom + apply.nin], output_storage[ofrom : ofrom + apply.nout]
            ifrom += apply.nin
            ofrom += apply.nout

    async def perform_async(self, node: Apply, inputs: Sequence[Any], output_storage: OutputStorageType) -> None:
        parts = list(self._slices(list(inputs), output_storage))
        pending = []
        # children that can be answered by one fused multi-GPU launch are grouped first
        groups = {}
        for apply, ins, outs in parts:
            key = apply.op.fusion_key() if isinstance(apply.op, FusableAsyncOp) else None
            if key is None:
                pending.append(apply.op.perform_async(apply, ins, outs))
            else:
                groups.setdefault(key, []).append((apply, ins, outs))
        for members in groups.values():
            pending.append(members[0][0].op.perform_fused(members))
        results = await asyncio.gather(*pending, return_exceptions=True)
        errors = [r for r in results if isinstance(r, BaseException)]
        if errors:
            raise errors[0]

    def perform(self, node: Apply, inputs: Sequence[Any], output_storage: OutputStorageType) -> None:
        loop = get_useful_event_loop()
        loop.run_until_complete(self.perform_async(node, inputs, output_storage))


class FusableAsyncOp(AsyncOp):
    """An ``AsyncOp`` whose siblings can be evaluated together by one engine call.

    ``fusion_key()`` returns a hashable key (e.g. the id of a :class:`FederatedEngine`) or ``None``;
    ``perform_fused(members)`` receives ``[(apply, inputs, output_storage), ...]`` of all siblings
    with the same key inside one :class:`ParallelAsyncOp`.
    """

    def fusion_key(self):
        return None

    async def perform_fused(self, members) -> None:
        await asyncio.gather(*[a.op.perform_async(a, i, o) for a, i, o in members])


def find_parallelizable_applies(fg: FunctionGraph, op_cls: type) -> List[Apply]:
    """Finds ≥ 2 apply nodes of ``op_cls`` that do not depend on each other.

    Walks the graph in topological order and collects independent nodes; a dependent node
    ends the collection (or restarts it, when only one node was collected so far).  Repeated
    application therefore fuses a dependency chain level by level.
    """
    applies: List[Apply] = []
    for apply in fg.toposort():
        if not isinstance(apply.op, op_cls):
            continue
        if not any(apply_depends_on(apply, a) for a in applies):
            applies.append(apply)
        elif len(applies) == 1:
            applies = [apply]
        else:
            break
    return applies if len(applies) > 1 else []


def parallelize_async_applies(fg: FunctionGraph, applies: Sequence[Apply]) -> None:
    """Replaces ``applies`` by a single :class:`ParallelAsyncOp` node, in place."""
    inputs: List[Variable] = []
    old_outputs: List[Variable] = []
    for apply in applies:
        inputs.extend(apply.inputs)
        old_outputs.extend(apply.outputs)
    new_outputs = ParallelAsyncOp(applies=applies).make_node(*inputs).outputs
    replace_all = getattr(fg, "replace_all_validate", fg.replace_all)
    replace_all(list(zip(old_outputs, new_outputs)))


def parallelize_all_async_applies(fg: FunctionGraph) -> None:
    """Fuses until no two independent ``AsyncOp`` applies remain.

    ``ParallelAsyncOp`` is itself an ``AsyncOp``; a fused node is never fused again with the
    nodes it already contains because those are gone from the graph, so this terminates.
    """
    applies = find_parallelizable_applies(fg, AsyncOp)
    while applies:
        parallelize_async_applies(fg, applies)
        applies = find_parallelizable_applies(fg, AsyncOp)


class AsyncFusionOptimizer(GraphRewriter):
    """Graph rewriter that parallelises ``AsyncOp.perform_async`` calls."""

    def add_requirements(self, fgraph: FunctionGraph) -> None:
        fgraph.attach_feature(ReplaceValidate())

    def apply(self, fgraph: FunctionGraph) -> None:
        parallelize_all_async_applies(fgraph)


if "fuse_asyncs" not in optdb:
    optdb.register("fuse_asyncs", AsyncFusionOptimizer(), "fast_run", position=90)

__all__ = [
    "AsyncOp",
    "AsyncFromFunctionOp",
    "ParallelAsyncOp",
    "FusableAsyncOp",
    "find_parallelizable_applies",
    "parallelize_async_applies",
    "parallelize_all_async_applies",
    "AsyncFusionOptimizer",
]
