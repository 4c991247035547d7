"""The runnable examples stay runnable (CPU dry runs of the scripts under examples/)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(600)


def _run(script, *args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", B200FED_CONNECT_SLEEP="0,0")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), *args], capture_output=True, text=True,
                         env=env, timeout=500)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return res.stdout


def test_custom_ode_example_recovers_the_parameters():
    out = _run("custom_ode.py", "--series", "40", "--nodes", "2")
    assert "MAP theta = [1.8" in out and "fused launches" in out


def test_batched_serving_example_batches_requests():
    out = _run("batched_serving.py", "--chains", "3", "--evals", "10", "--rows", "2000")
    assert "dynamic batching (K=3)" in out


def test_hierarchical_linreg_example_samples_the_same_posterior_both_ways():
    out = _run("hierarchical_linreg.py", "--nodes", "3", "--gpus", "1", "--draws", "40")
    lines = [l for l in out.splitlines() if "graph nodes" in l]
    assert len(lines) == 2 and "one Op per node" in lines[0] and "one Op for the federation" in lines[1]
    assert all("slope = 0.4" in l or "slope = 0.5" in l for l in lines)
