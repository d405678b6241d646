"""BASELINE.json config 4 (correctness side): a training step of the reference's models on the new
operators under DistributedDataParallel + nn.SyncBatchNorm, one process per GPU over NCCL, against
the same step on one GPU with the whole batch (baseline/ddp_check.py does the work; this file
launches it the way the driver launches bench.py).  Needs >= 2 GPUs: `gpurun --gpus 2`."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from baseline import refmodels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs"),
              pytest.mark.skipif(not refmodels.available(), reason="reference models not copied")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("model,H,W", [("GANet11", 96, 192), ("GANet_deep", 240, 624)])
def test_ddp_training_step_matches_single_gpu_batch(model, H, W):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "baseline", "ddp_check.py"), "--model", model,
           "--height", str(H), "--width", str(W)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(lines[-1])
    assert r.returncode == 0 and res["ok"], json.dumps(res)
    assert res["frozen_bn"]["grads_identical_on_all_ranks"] and res["sync_bn"]["sync_bn_layers"] > 50
