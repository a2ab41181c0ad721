"""bench.py's reference arm runs on CPU: check the JSON contract of its line here (no GPU needed)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    proc = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", "--cpu-sample-batch", "2"],
                          capture_output=True, text=True, timeout=600, cwd=REPO)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, proc.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["metric"] == "mtl_train_step_images_per_sec" and d["unit"] == "images/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # headline = the MTAN step (BASELINE.json configs[2]); the other GPU configs ride along as sub-records
    assert "workload" in d["config"] and d["config"]["workload"].startswith("mtan:")
    assert set(d["workloads"]) == {"csnet", "mtan_nyu"}
    for sub in d["workloads"].values():
        assert sub["impl"] == "reference" and sub["value"] > 0 and sub["unit"] == "images/s"


def test_reference_arm_runs_the_reference_loop():
    """With the reference tree present (/root/reference here, baseline/_ref on the GPU box) the CPU arm is
    the reference's own run_pipe + MTLModule, not the port."""
    import pytest

    from oracle import ref_runtime

    try:
        ref_runtime.reference_root()
    except RuntimeError:
        pytest.skip("no reference tree on this machine")
    import torch

    from vision_mtl_b200.synthetic import make_batch

    batch = make_batch(1, 32, 64, 19, "cityscapes", seed=11)
    ips, sec, threads = ref_runtime.time_reference_loop("mtan", batch, 19, 5e-4, 1, 0)
    assert ips > 0 and sec > 0 and threads >= 1
    ref = ref_runtime.load()
    assert ref["MTLModule"].__module__ == "vision_mtl.lit_module" and ref["run_pipe"].__module__ == "vision_mtl.training_lit"
    assert torch.optim.lr_scheduler.ReduceLROnPlateau is not None
