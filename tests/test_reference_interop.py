"""Wire compatibility with the UNMODIFIED reference package (CPU).

``baseline/_ref`` holds ``pip install --no-deps --target`` of /root/reference; its missing
third-party imports (betterproto, grpclib) are satisfied by ``baseline/shims``.  These tests
start the reference's own ``ArraysToArraysService`` and query it with this package's client, and
the other way round.  Skipped when the reference install is absent (it is git-ignored).
"""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from _helpers import ServerProcess, free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
SHIMS = os.path.join(ROOT, "baseline", "shims")

pytestmark = [
    pytest.mark.timeout(180),
    pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pytensor_federated")), reason="reference not installed"),
]


def _ref_env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([SHIMS, REF, ROOT])
    return env


REF_SERVER = textwrap.dedent(
    """
    import asyncio, sys
    import numpy as np
    import grpclib.server
    from pytensor_federated import ArraysToArraysService, wrap_logp_grad_func

    def f(a, b):
        return np.asarray(-(a - 1.0) ** 2 - np.sum((b + 2.0) ** 2)), [np.asarray(-2 * (a - 1.0)), -2 * (b + 2.0)]

    async def main(port):
        server = grpclib.server.Server([ArraysToArraysService(wrap_logp_grad_func(f))])
        await server.start("127.0.0.1", port)
        print("READY", flush=True)
        await server.wait_closed()

    asyncio.new_event_loop().run_until_complete(main(int(sys.argv[1])))
    """
)

REF_CLIENT = textwrap.dedent(
    """
    import sys
    import numpy as np
    from pytensor_federated import ArraysToArraysServiceClient
    client = ArraysToArraysServiceClient("127.0.0.1", int(sys.argv[1]))
    out = client.evaluate(np.array(6), np.array(7))
    assert out[0] == 42, out
    out = client.evaluate(np.array(1.5), np.array(4.0), use_stream=False)
    assert float(out[0]) == 6.0, out
    print("OK", flush=True)
    """
)


def test_product_client_talks_to_reference_server():
    from pytensor_federated_b200 import LogpGradServiceClient, service

    port = free_port()
    proc = subprocess.Popen([sys.executable, "-c", REF_SERVER, str(port)], env=_ref_env(), stdout=subprocess.PIPE, text=True)
    try:
        assert proc.stdout.readline().strip() == "READY"
        client = LogpGradServiceClient("127.0.0.1", port)
        logp, (da, db) = client.evaluate(np.array(3.0), np.array([0.0, 1.0]))
        assert logp == -(2.0**2) - (4.0 + 9.0)
        assert da == -4.0
        np.testing.assert_array_equal(db, [-4.0, -6.0])
        load = service.get_useful_event_loop().run_until_complete(service.get_load_async("127.0.0.1", port))
        assert load.n_clients == 1 and 0 < load.percent_ram < 100
        del client
    finally:
        proc.terminate()
        proc.wait(10)


def test_reference_client_talks_to_product_server():
    with ServerProcess() as server:
        res = subprocess.run(
            [sys.executable, "-c", REF_CLIENT, str(server.port)], env=_ref_env(), capture_output=True, text=True, timeout=120
        )
        assert res.returncode == 0, res.stderr[-2000:]
        assert "OK" in res.stdout
