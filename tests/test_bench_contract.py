"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (the CPU port of the path timed
on the host cores) prints exactly ONE line on stdout, a JSON object carrying the keys the driver reads; under torchrun only
rank 0 prints."""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONTRACT_KEYS = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                 "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches")


def _run(cmd, env=None):
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.strip()]


def _check_line(line: dict, n_gpus: int):
    for key in CONTRACT_KEYS:
        assert key in line, key
    assert line["impl"] == "reference" and line["n_gpus"] == n_gpus and line["gpu_launches"] == 0
    assert line["metric"].startswith("align-eval entity pairs/sec") and line["unit"] == "pairs/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["config"]["workload"] == "c4_1m" and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sub-problem" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and line["ms_per_step"] > 0


def test_reference_arm_prints_one_json_line():
    out = _run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert len(out) == 1, out
    _check_line(json.loads(out[0]), 1)


def test_reference_arm_under_torchrun_only_rank0_prints():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                "--master-port", "29533", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env)
    lines = [ln for ln in out if ln.startswith("{")]
    assert len(lines) == 1, out
    _check_line(json.loads(lines[0]), 2)
