"""The two demo CLIs end to end over real gRPC (CPU): node pool + hierarchical model."""
import os
import subprocess
import sys
import time

import numpy as np
import pytest

from _helpers import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(300)


@pytest.fixture(scope="module")
def node_pool():
    ports = [free_port() for _ in range(2)]
    proc = subprocess.Popen(
        [sys.executable, os.path.join(ROOT, "demo_node.py"), "--bind", "127.0.0.1", "--ports", ",".join(map(str, ports)),
         "--device", "cpu"],
        cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
    )
    from pytensor_federated_b200 import service
    from pytensor_federated_b200.utils import get_useful_event_loop

    deadline = time.time() + 120
    loop = get_useful_event_loop()
    while time.time() < deadline:
        loads = loop.run_until_complete(service.get_loads_async([("127.0.0.1", p) for p in ports], timeout=1))
        if all(l is not None for l in loads):
            break
        time.sleep(0.5)
    else:
        proc.terminate()
        pytest.fail("demo_node workers did not come up")
    yield ports
    proc.terminate()
    proc.wait(20)


@pytest.mark.parametrize("use_async", [False, True])
def test_hierarchical_demo_over_grpc(node_pool, use_async):
    sys.path.insert(0, ROOT)
    import demo_model
    import scipy.stats

    from pytensor_federated_b200.models import make_demo_data

    ops, client = demo_model.remote_ops_grpc("127.0.0.1", node_pool, 3, use_async)
    res = demo_model.run_model(ops, 3, tune=150, draws=100)
    x, y, _ = make_demo_data()
    slope = scipy.stats.linregress(x, y).slope
    names = ["intercept_mu", "intercept[0]", "intercept[1]", "intercept[2]", "slope"]
    assert abs(np.median(res.samples[:, names.index("slope")]) - slope) < 0.1
    assert res.divergences == 0
