"""bench.py's reference arm runs on the host cores alone, so its half of the JSON contract can be checked on a CPU-only
box: one line, the keys the driver parses, the same metric / unit / workload string as the product arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(*flags):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", *flags], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


@pytest.mark.parametrize("workload,metric,unit", [("c4", "solved trajectories/s", "trajectories/s"),
                                                  ("c5", "NUTS grad-evals/s", "grad-evals/s")])
def test_reference_arm_prints_the_contract_line(workload, metric, unit):
    sys.path.insert(0, ROOT)
    import bench
    line = _run("--workload", workload)
    assert REQUIRED <= set(line), sorted(REQUIRED - set(line))
    assert line["impl"] == "reference" and line["metric"] == metric and line["unit"] == unit
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["config"]["workload"] == bench.WORKLOADS[workload][2]  # the product arm names the same workload
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gradient_work_formula_reduces_to_the_primal_count():
    sys.path.insert(0, ROOT)
    import bench
    n, m, T, f, att, acc, B = 26, 6, 366, 130, 10000, 9000, 100
    primal = bench.gradient_work(n, m, T, f, att, acc, B, P=0)
    assert primal == (6 * att + 3 * B) * f + att * (70 * n + 50) + B * T * (14 * m + 45) + B * T * m * 25
    assert bench.gradient_work(n, m, T, f, att, acc, B, P=2) > 2.5 * primal * 0.5
    assert bench.gradient_work(n, m, T, f, att, acc, B, adjoint=True) == primal + acc * (27 * f + 98 * n)
