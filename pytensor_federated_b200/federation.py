"""Node-level view of a GPU federation, and launch helpers.

In the reference a *node* is a server process with a private dataset; a model that pools N nodes
creates N clients/Ops and lets the graph add the log-potentials
(``/root/reference/demo_model.py:17-36``, README "distributed" diagram).  Here the N nodes are the
N data shards resident on the GPUs of one NVSwitch domain, and :class:`NodeFederation` gives each
of them the reference's call signatures:

* ``fed.evaluate_node(i, intercept, slope) -> (logp, [grads])`` — one node's ``LogpGradFunc``;
* ``fed.evaluate_nodes({i: inputs, ...})`` — several nodes in ONE fused launch (what a fused
  ``ParallelAsyncOp`` of :class:`~pytensor_federated_b200.wrapper_ops.FederatedLogpGradOp`
  children calls);
* ``fed.node_ops()`` — one ``FederatedLogpGradOp`` per node for building models;
* ``fed.register_services("gpu")`` — makes ``LogpGradServiceClient("gpu", i)`` resolve to node i
  without sockets (the reference's client API, the NVLink data plane).
"""
from __future__ import annotations

import contextlib
import os
import socket
import threading
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .models.linreg import LinregShards
from .parallel.engine import FederatedEngine, FederationError


class NodeFederation:
    """Per-node access to an engine whose model keeps per-shard results (``LinregShards``)."""

    def __init__(self, engine: FederatedEngine) -> None:
        if not isinstance(engine.model, LinregShards):
            raise FederationError("NodeFederation needs a model with per-shard outputs (LinregShards)")
        self.engine = engine
        self.n_nodes = engine.model.n_shards_total
        self._intercepts = np.zeros(self.n_nodes)
        self._slopes = np.zeros(self.n_nodes)
        self._lock = threading.Lock()   # parameter staging + launch + un-staging of the result form one unit
        self.n_launches = 0

    # -- evaluation ----------------------------------------------------------------------------
    def evaluate_nodes(self, requests: Dict[int, Sequence[np.ndarray]]) -> Dict[int, Tuple[np.ndarray, List[np.ndarray]]]:
        """``{node: (intercept, slope)} -> {node: (logp, [d_intercept, d_slope])}``, one launch."""
        with self._lock:
            for node, (a, b) in requests.items():
                self._intercepts[node] = float(np.asarray(a))
                self._slopes[node] = float(np.asarray(b))
            raw = self.engine.evaluate_raw([self._intercepts, self._slopes])
            self.n_launches += 1
            per = LinregShards.per_shard(raw)
            return {
                node: (np.array(per[node, 0]), [np.array(per[node, 1]), np.array(per[node, 2])])
                for node in requests
            }

    def evaluate_node(self, node: int, intercept, slope) -> Tuple[np.ndarray, List[np.ndarray]]:
        return self.evaluate_nodes({node: (intercept, slope)})[node]

    def logp_grad_func(self, node: int) -> Callable:
        """The node as a plain ``LogpGradFunc`` (usable with the generic ``LogpGradOp``)."""
        return lambda intercept, slope: self.evaluate_node(node, intercept, slope)

    def compute_func(self, node: int) -> Callable:
        """The node as a ``ComputeFunc``: ``(logp, d_intercept, d_slope)``."""

        def compute(intercept, slope):
            logp, grads = self.evaluate_node(node, intercept, slope)
            return (logp, *grads)

        return compute

    # -- graph integration ---------------------------------------------------------------------
    def node_ops(self):
        from .wrapper_ops import FederatedLogpGradOp

        return [FederatedLogpGradOp(self, i) for i in range(self.n_nodes)]

    # -- reference client API ------------------------------------------------------------------
    def register_services(self, host: str = "gpu", first_port: int = 0) -> List[Tuple[str, int]]:
        from . import service

        addresses = []
        for i in range(self.n_nodes):
            service.register_local_node(host, first_port + i, self.compute_func(i), name=f"{host}:{first_port + i}")
            addresses.append((host, first_port + i))
        self._registered = addresses
        return addresses

    def unregister_services(self) -> None:
        from . import service

        for host, port in getattr(self, "_registered", []):
            service.unregister_local_node(host, port)
        self._registered = []

    def shutdown(self) -> None:
        self.unregister_services()
        self.engine.shutdown()


def free_port() -> int:
    """An unused TCP port on 127.0.0.1 (rendezvous of a freshly launched federation)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _peer_main(rank: int, world: int, port: int, build_model, backend: str, device_type: str, timeout: float) -> None:
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if device_type == "cuda":
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = build_model(rank, world, dev)
        eng = FederatedEngine(model, backend=backend, device=dev, timeout=timeout)
        eng.serve()
        eng.shutdown()
    finally:
        dist.destroy_process_group()


@contextlib.contextmanager
def launch_federation(build_model: Callable, n_nodes: int, *, device_type: Optional[str] = None,
                      backend: str = "auto", timeout: float = 3600.0):
    """Starts ``n_nodes - 1`` peer processes (one per GPU) and yields the root's engine.

    ``build_model(rank, world, device) -> ShardModel`` builds each node's private shard model and
    must be picklable (module-level function).  The calling process is rank 0, the client.  On
    exit the peers are drained and joined.  ``device_type="cpu"`` runs the same topology over
    gloo with the collective backend (plumbing tests).
    """
    import torch
    import torch.distributed as dist
    import torch.multiprocessing as mp

    if device_type is None:
        device_type = "cuda" if torch.cuda.is_available() else "cpu"
    if device_type == "cuda" and torch.cuda.device_count() < n_nodes:
        raise FederationError(f"{n_nodes} nodes requested but only {torch.cuda.device_count()} GPUs are visible")
    port = free_port()
    ctx = mp.get_context("spawn")
    procs = [
        ctx.Process(target=_peer_main, args=(r, n_nodes, port, build_model, backend, device_type, timeout), daemon=True)
        for r in range(1, n_nodes)
    ]
    for p in procs:
        p.start()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    owns_pg = False
    engine = None
    try:
        if n_nodes > 1:
            if device_type == "cuda":
                torch.cuda.set_device(0)
                dist.init_process_group("nccl", rank=0, world_size=n_nodes, device_id=torch.device("cuda", 0))
            else:
                dist.init_process_group("gloo", rank=0, world_size=n_nodes)
            owns_pg = True
        dev = torch.device("cuda", 0) if device_type == "cuda" else torch.device("cpu")
        engine = FederatedEngine(build_model(0, n_nodes, dev), backend=backend, device=dev, timeout=timeout)
        yield engine
    finally:
        if engine is not None:
            engine.shutdown()
        if owns_pg:
            dist.destroy_process_group()
        for p in procs:
            p.join(60)
            if p.is_alive():
                p.terminate()


__all__ = ["NodeFederation", "launch_federation", "free_port"]
