"""Dynamic batching of concurrent requests on a node (pytensor_federated_b200/batching.py)."""
import asyncio

import numpy as np
import pytest

from pytensor_federated_b200 import ArraysToArraysServiceClient, service
from pytensor_federated_b200.batching import DynamicBatcher, stacked_compute_func
from pytensor_federated_b200.rpc import Server
from pytensor_federated_b200.utils import get_useful_event_loop

pytestmark = pytest.mark.timeout(120)


def _chain_evaluator(calls):
    """Stand-in for a K-chain engine: logp and gradient of N(0, I) per chain, counts launches."""

    def evaluate(theta):                       # [K, D]
        calls.append(theta.shape[0])
        return [-0.5 * np.sum(theta * theta, axis=1), -theta]

    return evaluate


def test_concurrent_requests_share_one_batched_call():
    calls = []
    batcher = DynamicBatcher(stacked_compute_func(_chain_evaluator(calls), max_batch=4), max_batch=4, max_delay=0.05)

    async def main():
        thetas = [np.full(3, float(i)) for i in range(6)]
        results = await asyncio.gather(*[batcher(t) for t in thetas])
        await batcher.close()
        return thetas, results

    thetas, results = get_useful_event_loop().run_until_complete(main())
    for theta, (logp, grad) in zip(thetas, results):
        assert logp == -0.5 * np.sum(theta * theta)
        np.testing.assert_array_equal(grad, -theta)
    assert calls == [4, 4]                      # 6 requests -> a full batch + a padded batch of 2
    assert batcher.n_batches == 2 and batcher.n_requests == 6


def test_errors_reach_every_request_of_the_batch_and_the_worker_survives():
    def flaky(requests):
        if any(r[0][0] < 0 for r in requests):
            raise ValueError("negative input")
        return [[r[0] * 2] for r in requests]

    batcher = DynamicBatcher(flaky, max_batch=8, max_delay=0.02)

    async def main():
        bad = await asyncio.gather(batcher(np.array([1.0])), batcher(np.array([-1.0])), return_exceptions=True)
        good = await batcher(np.array([3.0]))
        await batcher.close()
        return bad, good

    bad, good = get_useful_event_loop().run_until_complete(main())
    assert all(isinstance(b, ValueError) for b in bad)
    np.testing.assert_array_equal(good[0], [6.0])
    with pytest.raises(ValueError):
        DynamicBatcher(flaky, max_batch=0)
    with pytest.raises(ValueError):
        stacked_compute_func(lambda x: [x], max_batch=2)([(np.zeros(1),)] * 3)


def test_batching_node_over_grpc_serves_several_clients_with_fewer_launches(monkeypatch):
    """Four clients (think: four chains) against ONE node: the node answers them from shared launches."""
    monkeypatch.setenv("B200FED_CONNECT_SLEEP", "0,0")
    calls = []
    batcher = DynamicBatcher(stacked_compute_func(_chain_evaluator(calls), max_batch=4), max_batch=4, max_delay=0.05)
    loop = get_useful_event_loop()
    server = Server([service.ArraysToArraysService(batcher)])
    port = loop.run_until_complete(server.start("127.0.0.1", 0))
    clients = [ArraysToArraysServiceClient("127.0.0.1", port) for _ in range(4)]

    async def round_trip(step):
        return await asyncio.gather(*[c.evaluate_async(np.full(2, float(10 * step + i))) for i, c in enumerate(clients)])

    try:
        for step in range(3):
            answers = loop.run_until_complete(round_trip(step))
            for i, (logp, grad) in enumerate(answers):
                theta = np.full(2, float(10 * step + i))
                assert logp == -0.5 * np.sum(theta * theta)
                np.testing.assert_array_equal(grad, -theta)
    finally:
        del clients
        loop.run_until_complete(batcher.close())
        loop.run_until_complete(server.close(None))
    assert len(calls) < 12                      # 12 requests, far fewer launches
    assert batcher.n_requests == 12


def test_coroutine_compute_function_behind_an_in_process_node():
    batcher = DynamicBatcher(lambda reqs: [[r[0] + 1] for r in reqs], max_batch=2, max_delay=0.0)
    service.register_local_node("batched", 0, batcher)
    try:
        client = ArraysToArraysServiceClient("batched", 0)
        np.testing.assert_array_equal(client.evaluate(np.array([1.0, 2.0]))[0], [2.0, 3.0])
        np.testing.assert_array_equal(client.evaluate(np.array([5.0]))[0], [6.0])   # cached fast path
    finally:
        service.unregister_local_node("batched", 0)
        get_useful_event_loop().run_until_complete(batcher.close())
