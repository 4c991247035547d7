"""Transport tests: server wrapper, load probes, balancing, pickling, failover.

Every "node" is a child process on 127.0.0.1, an unbound port plays a dead node,
``terminate()`` injects failures — the strategy of the reference's
``test_service.py:109-283`` without its fixed sleeps and fixed ports.
"""
import asyncio
import multiprocessing
import os
import pickle
import time
from unittest import mock

import numpy as np
import pytest

from pytensor_federated_b200 import LogpServiceClient, service
from pytensor_federated_b200.npproto.utils import ndarray_from_numpy, ndarray_to_numpy
from pytensor_federated_b200.rpc import GetLoadResult, InputArrays, OutputArrays, Server
from pytensor_federated_b200.utils import get_useful_event_loop

from _helpers import (
    ProductTester,
    ServerProcess,
    free_port,
    product_func,
    run_product_queries,
    start_fleet,
)

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(180)


def test_run_compute_func_decodes_and_encodes():
    def compute(a, b):
        return np.sum(a + b), np.prod(a + b), np.prod(a - b)

    a = np.array([1, 2])
    b = np.array([3, 4.5])
    request = InputArrays(items=[ndarray_from_numpy(a), ndarray_from_numpy(b)], uuid="req-1")
    response = service._run_compute_func(request, compute)
    assert isinstance(response, OutputArrays)
    assert response.uuid == "req-1"
    np.testing.assert_array_equal(ndarray_to_numpy(response.items[0]), 10.5)
    np.testing.assert_array_equal(ndarray_to_numpy(response.items[1]), 4 * 6.5)
    np.testing.assert_array_equal(ndarray_to_numpy(response.items[2]), -2 * -2.5)


@mock.patch("psutil.getloadavg", return_value=[0.1, 0.2, 0.3])
@mock.patch("psutil.cpu_count", return_value=3)
def test_determine_load(mock_cpu_count, mock_getloadavg):
    svc = service.ArraysToArraysService(product_func)
    # primed exactly once at construction so that psutil starts monitoring
    mock_getloadavg.assert_called_once_with()
    mock_cpu_count.assert_called_once_with()
    svc._n_clients = 3
    load = svc.determine_load()
    assert load.n_clients == 3
    assert load.percent_cpu == 0.1 / 3 * 100
    assert 0 < load.percent_ram < 100


def test_gpu_load_is_optional_and_never_serialised():
    svc = service.ArraysToArraysService(product_func, gpu_index=0)
    load = svc.determine_load()
    assert load.gpu is None or (0 <= load.gpu[0] <= 100 and 0 <= load.gpu[1] <= 100)
    with mock.patch.object(service, "gpu_load", return_value=(37.0, 12.5)):
        load = svc.determine_load()
    assert load.gpu == (37.0, 12.5)
    decoded = GetLoadResult.FromString(bytes(load))
    assert decoded.gpu is None and decoded.n_clients == load.n_clients


def test_stream_counts_clients_even_on_errors():
    svc = service.ArraysToArraysService(product_func)

    async def requests(fail):
        yield InputArrays(items=[ndarray_from_numpy(np.array(2))], uuid="a")
        if fail:
            raise RuntimeError("client vanished")

    async def drive(fail):
        seen = []
        try:
            async for out in svc.evaluate_stream(requests(fail)):
                seen.append(svc._n_clients)
        except RuntimeError:
            pass
        return seen

    loop = get_useful_event_loop()
    assert loop.run_until_complete(drive(False)) == [1]
    assert svc._n_clients == 0
    assert loop.run_until_complete(drive(True)) == [1]
    assert svc._n_clients == 0  # the reference leaks a count here


def test_client_argchecking():
    client = service.ArraysToArraysServiceClient("localhost", 9999)
    with pytest.raises(ValueError, match="must be >= 0"):
        client.evaluate(np.array(0), np.array(1), retries=-1)


def test_client_pickles_without_connection_state():
    client = service.ArraysToArraysServiceClient(hosts_and_ports=[("127.0.0.1", 1), ("127.0.0.1", 2)])
    clone = pickle.loads(pickle.dumps(client))
    assert clone._hosts_and_ports == [("127.0.0.1", 1), ("127.0.0.1", 2)]
    assert service.thread_pid_id(clone) != service.thread_pid_id(client)


def test_local_node_fast_path_needs_no_sockets():
    calls = []

    def compute(a, b):
        calls.append((a, b))
        return (np.asarray(a * b),)

    node = service.register_local_node("gpu", 0, compute)
    try:
        client = service.ArraysToArraysServiceClient("gpu", 0)
        (out,) = client.evaluate(np.array(3.0), np.array(4.0))
        assert out == 12.0 and node.n_clients == 1
        (out,) = client(np.array(2.0), np.array(4.0))  # cached, loop-free path
        assert out == 8.0 and len(calls) == 2
        loads = get_useful_event_loop().run_until_complete(
            service.get_loads_async([("gpu", 0), ("127.0.0.1", free_port())], timeout=1)
        )
        assert loads[0].n_clients == 1 and loads[1] is None
        del client
        assert node.n_clients == 0
    finally:
        service.unregister_local_node("gpu", 0)


def test_load_probes_and_balancing():
    servers = start_fleet([3, 4, 2])
    offline = ("127.0.0.1", free_port())
    try:
        addresses = [offline] + [("127.0.0.1", s.port) for s in servers]
        loop = get_useful_event_loop()
        loads = loop.run_until_complete(service.get_loads_async(addresses, timeout=3))
        assert loads[0] is None
        assert [l.n_clients for l in loads[1:]] == [3, 4, 2]
        assert all(isinstance(l, GetLoadResult) for l in loads[1:])

        client = service.ArraysToArraysServiceClient(hosts_and_ports=addresses)
        result = client.evaluate(np.array(2), np.array(3))
        assert isinstance(result, list) and isinstance(result[0], np.ndarray)
        assert result[0] == 6
        cid = service.thread_pid_id(client)
        assert int(service._privates[cid].channel._port) == servers[2].port  # fewest clients
        # unary mode on the same connection
        assert client.evaluate(np.array(5), np.array(3), use_stream=False)[0] == 15
        # the real counter: one open stream on that server
        after = loop.run_until_complete(service.get_load_async("127.0.0.1", servers[2].port))
        assert after.n_clients == 3
        del client
    finally:
        for s in servers:
            s.terminate()


@pytest.mark.parametrize("eval_on_main", [False, True])
@pytest.mark.parametrize("mp_start_method", ["spawn", "fork"])
def test_client_multiprocessing(eval_on_main, mp_start_method):
    if mp_start_method == "fork" and eval_on_main:
        pytest.skip("grpc's C core does not survive fork() once initialised in the parent; use spawn")
    with ServerProcess() as server:
        if mp_start_method == "fork":
            # fork() is only safe while gRPC has never been initialised in the parent, which cannot be
            # guaranteed inside a long pytest session -> run the scenario in a fresh interpreter
            import os
            import subprocess
            import sys

            here = os.path.dirname(os.path.abspath(__file__))
            env = dict(os.environ, PYTHONPATH=os.pathsep.join([here, os.path.dirname(here)]))
            res = subprocess.run(
                [sys.executable, "-c", f"import _helpers; _helpers.fork_pool_scenario({server.port})"],
                env=env, capture_output=True, text=True, timeout=150,
            )
            assert res.returncode == 0 and "FORK-OK" in res.stdout, res.stderr[-2000:]
            return
        ctx = multiprocessing.get_context(mp_start_method)
        client = service.ArraysToArraysServiceClient("127.0.0.1", server.port)
        tester = ProductTester(client)
        if eval_on_main:
            assert tester.run(n=2)
        with ctx.Pool(processes=3) as pool:
            assert all(pool.map(run_product_queries, [client] * 4))
        with ctx.Pool(processes=3) as pool:
            assert all(pool.map(tester.run, [25] * 4))


def test_client_failover():
    s_busy = ServerProcess(n_clients=5).start()
    s_idle = ServerProcess(n_clients=2).start()
    try:
        client = service.ArraysToArraysServiceClient(
            hosts_and_ports=[("127.0.0.1", s_busy.port), ("127.0.0.1", s_idle.port)]
        )
        cid = service.thread_pid_id(client)
        assert cid not in service._privates
        assert client.evaluate(np.array(2), np.array(3))[0] == 6
        channel1 = service._privates[cid].channel
        assert int(channel1._port) == s_idle.port

        # killing the node is not noticed until the next call ...
        s_idle.terminate()
        assert int(service._privates[cid].channel._port) == s_idle.port
        # ... which fails over to the surviving replica with one retry
        assert client.evaluate(np.array(2), np.array(4), retries=1)[0] == 8
        assert channel1.closed
        assert int(service._privates[cid].channel._port) == s_busy.port

        s_busy.terminate()
        with pytest.raises(TimeoutError, match="None of 2 servers responded"):
            client.evaluate(np.array(2), np.array(4))
    finally:
        s_busy.terminate()
        s_idle.terminate()


def test_single_dead_server_raises_after_retries():
    client = service.ArraysToArraysServiceClient("127.0.0.1", free_port())
    with pytest.raises(service.StreamTerminatedError):
        client.evaluate(np.array(1), retries=0)


def test_a_failing_compute_function_is_reported_once_not_retried():
    """A compute function that raises on the node is an application error: the client must surface it
    (with the server's message) after ONE execution, not treat it as a lost connection and re-run it."""
    calls = []

    def picky(a):
        calls.append(float(a))
        if a < 0:
            raise ValueError("negative input rejected")
        return [np.sqrt(a)]

    loop = get_useful_event_loop()
    server = Server([service.ArraysToArraysService(picky)])
    port = loop.run_until_complete(server.start("127.0.0.1", 0))
    client = service.ArraysToArraysServiceClient("127.0.0.1", port)
    try:
        assert client.evaluate(np.array(4.0))[0] == 2.0
        for use_stream in (True, False):
            calls.clear()
            with pytest.raises(service.RemoteComputeError, match="negative input rejected") as err:
                loop.run_until_complete(client.evaluate_async(np.array(-1.0), use_stream=use_stream, retries=2))
            assert calls == [-1.0]
            assert not isinstance(err.value, service.StreamTerminatedError)
            # the connection is re-established transparently for the next (valid) call
            assert client.evaluate(np.array(9.0))[0] == 3.0
    finally:
        del client
        loop.run_until_complete(server.close(None))


def test_concurrent_async_clients_share_a_loop():
    """Fan-out as the fused graph node does it: N clients, one gather."""
    servers = start_fleet([0, 0, 0])
    try:
        clients = [service.ArraysToArraysServiceClient("127.0.0.1", s.port) for s in servers]

        async def fan_out():
            coros = [c.evaluate_async(np.array(i + 1), np.array(10)) for i, c in enumerate(clients)]
            return await asyncio.gather(*coros)

        results = get_useful_event_loop().run_until_complete(fan_out())
        assert [int(r[0]) for r in results] == [10, 20, 30]
        del clients
    finally:
        for s in servers:
            s.terminate()


def test_prometheus_metrics_of_a_serving_node():
    """Counters, latency histogram and the client gauge; scraped over HTTP like Prometheus would."""
    import urllib.request

    pytest.importorskip("prometheus_client")
    from pytensor_federated_b200.metrics import ServiceMetrics, metrics_from_env
    from pytensor_federated_b200.rpc import InputArrays

    assert metrics_from_env() is None          # opt-in only
    metrics = ServiceMetrics(port=0, addr="127.0.0.1")
    calls = {"n": 0}

    def flaky(a):
        calls["n"] += 1
        if calls["n"] == 3:
            raise RuntimeError("boom")
        return [a * 2]

    svc = service.ArraysToArraysService(flaky, metrics=metrics)
    loop = get_useful_event_loop()
    request = InputArrays.from_arrays([np.arange(3.0)], uuid="u")

    async def drive():
        async def requests():
            for _ in range(2):
                yield request

        gen = svc.evaluate_stream(requests())
        first = await gen.__anext__()
        assert svc._n_clients == 1 and "b200fed_clients 1.0" in metrics.render()
        await gen.__anext__()
        with pytest.raises(StopAsyncIteration):
            await gen.__anext__()
        with pytest.raises(RuntimeError):
            await svc.evaluate(request)
        return first

    first = loop.run_until_complete(drive())
    np.testing.assert_array_equal(first.arrays[0], [0.0, 2.0, 4.0])
    try:
        text = urllib.request.urlopen(f"http://127.0.0.1:{metrics.port}/metrics", timeout=5).read().decode()
    finally:
        metrics.close()
    assert "b200fed_evaluations_total 3.0" in text and "b200fed_errors_total 1.0" in text
    assert "b200fed_clients 0.0" in text and "b200fed_compute_seconds_count 3.0" in text


def test_per_attempt_timeout_detects_a_stalled_node(monkeypatch):
    """A node that accepts the call but does not answer: ``timeout`` turns it into ``TimeoutError`` and the
    stale answer can never be mistaken for the reply to a later request (fresh stream + uuid check)."""
    monkeypatch.setenv("B200FED_CONNECT_SLEEP", "0,0")
    with ServerProcess(func="slow_product") as srv:
        client = service.ArraysToArraysServiceClient("127.0.0.1", srv.port)
        (out,) = client.evaluate(np.array(3), np.array(4), timeout=5.0)
        assert out == 12
        t0 = time.perf_counter()
        with pytest.raises(TimeoutError):
            client.evaluate(np.array(99), np.array(2), timeout=0.25, retries=0)
        assert time.perf_counter() - t0 < 1.5
        # the node finishes its stalled computation, then serves the next request correctly
        (out,) = client.evaluate(np.array(5), np.array(6), timeout=10.0)
        assert out == 30
        logp_client = LogpServiceClient("127.0.0.1", srv.port)
        assert logp_client.evaluate(np.array(2), np.array(8), timeout=10.0, retries=1) == 16


def test_stalled_replica_is_quarantined_and_the_call_fails_over(monkeypatch):
    monkeypatch.setenv("B200FED_CONNECT_SLEEP", "0,0")
    service._quarantine.clear()
    stalling = ServerProcess(func="slow_product", n_clients=0).start()     # preferred by the balancer
    healthy = ServerProcess(func="product", n_clients=5).start()
    try:
        client = service.ArraysToArraysServiceClient(
            hosts_and_ports=[("127.0.0.1", stalling.port), ("127.0.0.1", healthy.port)])
        (out,) = client.evaluate(np.array(99), np.array(2), timeout=1.0, retries=1)
        assert out == 198
        cid = service.thread_pid_id(client)
        assert int(service._privates[cid].channel._port) == healthy.port
        assert ("127.0.0.1", stalling.port) in service._quarantine
        # with every replica quarantined the balancer falls back to the full list instead of giving up
        service._quarantine[("127.0.0.1", healthy.port)] = time.monotonic() + 60
        other = service.ArraysToArraysServiceClient(
            hosts_and_ports=[("127.0.0.1", stalling.port), ("127.0.0.1", healthy.port)])
        assert other.evaluate(np.array(2), np.array(2))[0] == 4
    finally:
        service._quarantine.clear()
        stalling.terminate()
        healthy.terminate()


def test_sigterm_drains_the_request_in_flight(tmp_path, monkeypatch):
    """A node started with service.serve() finishes the running request on SIGTERM, then exits cleanly."""
    import signal
    import subprocess
    import sys
    import textwrap

    monkeypatch.setenv("B200FED_CONNECT_SLEEP", "0,0")
    script = tmp_path / "node.py"
    script.write_text(textwrap.dedent(
        """
        import asyncio, sys, time
        import numpy as np
        sys.path.insert(0, %r)
        from pytensor_federated_b200 import service

        def slow_square(a):
            time.sleep(0.8)
            return [a * a]

        asyncio.new_event_loop().run_until_complete(
            service.serve(slow_square, "127.0.0.1", int(sys.argv[1]), offload=True, ready=lambda p: print("ready", flush=True)))
        """
    ) % str(ROOT_DIR))
    port = free_port()
    proc = subprocess.Popen([sys.executable, str(script), str(port)], stdout=subprocess.PIPE, text=True)
    try:
        assert proc.stdout.readline().strip() == "ready"
        client = service.ArraysToArraysServiceClient("127.0.0.1", port)
        loop = get_useful_event_loop()

        async def call_and_terminate():
            task = asyncio.ensure_future(client.evaluate_async(np.array(7.0), retries=0))
            await asyncio.sleep(0.3)                # the request is being computed now
            proc.send_signal(signal.SIGTERM)
            return await task

        (out,) = loop.run_until_complete(call_and_terminate())
        assert out == 49.0
        assert proc.wait(timeout=15) == 0
    finally:
        if proc.poll() is None:
            proc.kill()
