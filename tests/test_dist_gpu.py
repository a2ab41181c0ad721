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
    print(lines[-1])  # shown by pytest when an assertion below fails
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
    assert r["loss_rel"] <= 1e-6, r
    assert r["running_max_rel"] <= 1e-5, r
    # Gradients: this random-init fixture is ill-conditioned at the 1e-3 level for ANY fp32 arithmetic (the CPU
    # reference in fp32 differs from fp64 by a median 7e-4, see _dist_worker.py), so the 2-GPU step is measured
    # against the fp64 truth next to the single-GPU step on the concatenated batch: it must be as close.
    for q in ("median", "p90", "total"):
        assert r[f"yard_sync_{q}"] <= 2.0 * r[f"yard_whole_{q}"] + 1e-5, (q, r)
    assert r["grad_total_rel_l2"] <= 4.0 * r["yard_whole_total"] + 1e-5, r  # and the two fp32 runs agree to that level
    # argmax ties aside, the predictions are those of the whole-batch run
    assert r["confusion_mismatch"] <= max(2, r["pixels"] // 10000), r
