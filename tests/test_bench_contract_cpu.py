"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port timed on the host cores) prints ONE JSON line with
the keys of the benchmark contract, and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          env=env)


def test_reference_arm_json_line():
    res = _run("--impl", "reference", "--workload", "cfg2", "--steps", "2", "--warmup", "1",
               env=dict(os.environ, MAC_BENCH_REFERENCE_BUDGET_S="2"))
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "coverage_gain_evals_per_sec" and d["unit"] == "evals/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # the config object is built by ONE function for both arms (the driver compares them key for key)
    import argparse
    import bench
    args = argparse.Namespace(gpus=1, workload="cfg2")
    assert d["config"] == bench.make_config(args, bench.WORKLOADS["cfg2"])


def test_reference_arm_uses_all_host_threads_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1 into every rank; the CPU arm must still use every core it may run on."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", MAC_BENCH_REFERENCE_BUDGET_S="2")
    res = _run("--impl", "reference", "--workload", "cfg2", "--steps", "1", "--warmup", "1", "--gpus", "2", env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads(res.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = _run("--impl", "reference", "--workload", "cfg2", "--steps", "1", "--warmup", "1", "--gpus", "2", env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        return
    res = _run("--workload", "cfg2", "--steps", "1", "--warmup", "1")
    assert res.returncode != 0 and "CUDA" in (res.stderr + res.stdout)
