"""Sharded state over several GPUs of one node (needs >= 2 visible GPUs, skipped otherwise; the one-GPU
coverage of the same code is tests/test_sv_gpu.py): tools/sharded_check.py -- the q24 depth-20 fused trace on a
state sharded by global qubits (qb200_sv_*), exchanges over NVLink peer memory, must equal the single-GPU
result, one process per GPU under torchrun and one process driving all GPUs."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def visible_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_equals_single_gpu(world):
    if visible_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "sharded_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["ok"] and res["max_abs_err_vs_single_gpu"] < 1e-6 and res["swaps"] >= 1


@pytest.mark.parametrize("world", [2, 4])
def test_single_process_multi_device_equals_single_gpu(world):
    if visible_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sharded_check.py"), "--single", str(world)],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["ok"] and res["max_abs_err_vs_single_gpu"] < 1e-6 and res["swaps"] >= 1
