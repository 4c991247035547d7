"""bench.py's output contract, exercised on CPU through the reference arm (tiny shapes)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "impl"}

pytestmark = pytest.mark.timeout(300)


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "pytensor_federated")),
                    reason="reference not installed (baseline/_ref is git-ignored)")
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, B200FED_REF_ALLOW_CPU="1")
    res = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "2000", "--features", "64",
         "--shards", "2", "--steps", "4", "--warmup", "3"],
        cwd=ROOT, env=env, capture_output=True, text=True, timeout=240,
    )
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert REQUIRED <= set(line)
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["scaling"] == "strong"
    assert line["steps"] == 4 and line["warmup"] == 3 and line["value"] > 0
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert "unmodified reference" in line["config"]["note"]


def test_reference_arm_reports_unavailable_instead_of_failing(tmp_path):
    """Without the install the arm must print {"impl": "reference", "unavailable": ...} and exit 0."""
    code = (
        "import sys, json; sys.path.insert(0, %r); import reference_arm as r; "
        "r._paths = lambda: None; "
        "import types; "
        "r.main(types.SimpleNamespace(shards=2, rows=10, features=64, steps=1, warmup=1, out=None))"
    ) % os.path.join(ROOT, "baseline")
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    res = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr[-1500:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][0])
    assert line["impl"] == "reference" and "unavailable" in line
