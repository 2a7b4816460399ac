"""CPU checks of bench.py's driver contract: the reference arm runs without a GPU (it is the oracle's CPU path), prints ONE
JSON line with the agreed keys, and the N > 1 form lets rank 0 alone work.  The GPU arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=e)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "int8_qlinear_tops" and d["unit"] == "TOPS" and d["higher_is_better"] is True
    assert d["config"]["workload"] == "llama7b_linears_2048tok" and d["config"]["tokens_per_gpu"] == 2048
    assert d["dtype"] == "int8" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "TOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_gpu_arm_refuses_to_run_without_cuda():
    p = _run(["--steps", "1", "--warmup", "1"])
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
