"""2-GPU NCCL parity of the data-parallel step (skipped with fewer than two GPUs): see tests/_dist_worker.py."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "_dist_worker.py"), mode]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
    assert p.returncode == 0 and lines, (p.returncode, p.stdout[-2000:], p.stderr[-4000:])
    return json.loads(lines[-1][len("RESULT "):])


def test_ddp_step_global_confusion_and_gradients():
    """Default semantics: the all-reduced confusion matrix is bit-equal to the sum of the per-shard matrices of a
    single process, DDP's gradients are the mean of the per-shard gradients."""
    r = _run("ddp")
    assert r["confusion_equal"], r
    assert r["loss_rel"] <= 1e-6, r
    assert r["grad_total_rel_l2"] <= 1e-5 and r["grad_max_rel_l2"] <= 1e-4, r


def test_sync_stats_step_equals_single_gpu_on_concatenated_batch():
    """Global-batch-exact mode: the 2-GPU step is the single-GPU step on the concatenated batch."""
    r = _run("sync")
    assert r["loss_rel"] <= 1e-4, r
    assert r["grad_total_rel_l2"] <= 1e-4, r
    assert r["grad_median_rel_l2"] <= 1e-4, r
    assert r["running_max_rel"] <= 1e-5, r
    # argmax ties aside, the predictions are those of the whole-batch run
    assert r["confusion_mismatch"] <= max(2, r["pixels"] // 10000), r
