"""CPU: the bench.py contract that does not need a GPU -- the reference arm prints ONE JSON line with the agreed keys,
and the GPU arm refuses to run without a device (the product path has no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

from conftest import ROOT


def _run(args, timeout=240):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=env)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("queries/sec") and "workload" in d["config"] and d["data"] == "synthetic"
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle/st_util.semantic_search" in cb["sample"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "c1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1", "--warmup", "0"], timeout=120)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
