"""Multi-GPU paths on real devices (skipped on a one-GPU box; the CPU suite covers the same host logic
with gloo, tests/test_parallel_cpu.py): the NCCL scatter/gather edge of BASELINE configs[3] and one
DistributedDataParallel training step (configs[4]; reference base_model.py:111-114)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
two_gpus = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                              reason="needs two GPUs")


def _torchrun(script, *args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900)


@two_gpus
def test_sharded_enhance_matches_single_gpu():
    run = _torchrun("sharded_check.py", "--batch", "3", "--size", "128", port=29531)
    assert run.returncode == 0, run.stdout[-1500:] + run.stderr[-1500:]
    assert "identical to the single-GPU result" in run.stdout
    assert "identical to the single-GPU results (pipelined)" in run.stdout


@two_gpus
def test_ddp_training_step_synchronises_gradients():
    run = _torchrun("ddp_train_step.py", "--batch", "2", "--size", "64", "--steps", "2", port=29532)
    assert run.returncode == 0, run.stderr[-2000:]
    line = json.loads([l for l in run.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["ranks_hold_identical_parameters"] is True
    assert line["loss"] == line["loss"]          # finite (not NaN)
