"""bench.py contract pieces that run without a GPU: the reference arm (CPU port of the oracle) prints one
JSON line with the keys the driver reads; non-zero ranks of a torchrun launch exit silently."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra, *args):
    env = dict(os.environ, BENCH_CPU_BUDGET_S="0.3", **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                          capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_emits_one_json_line():
    r = _run({}, "--gpus", "1", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e5          # a compiled port on any host does far more than the Python reference's ~5e3/s
    # the reference's own Python loop beside it, when the reference (or its shipped copy baseline/_ref) is there
    pr = cb["python_reference"]
    if "unavailable" not in pr:
        assert pr["kind"] == "reference" and pr["cores"] == (os.cpu_count() or 1)
        assert 1e2 < pr["single_process_value"] < 1e5 and pr["value"] >= 0.5 * pr["single_process_value"]
        assert d["value"] > 20 * pr["value"]


def test_both_arms_name_the_same_workload():
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"workload": WORKLOAD.format(') == 2      # GPU arm and reference arm share one string


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""
