"""GPU parity of the drop-in models / step module against (a) the committed reference outputs in
tests/golden and (b) the CPU oracle run on the same inputs.  Tolerance: 1e-4 relative
(tensor-wise) on losses, logits, gradients and depth metrics; confusion matrices bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import metrics_np as MN
from oracle import torch_port as TP
from oracle.make_golden import MTAN_CASES

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-4


def dev():
    return torch.device("cuda:0")


def rel_err(got, ref):
    got = torch.as_tensor(np.asarray(got) if not isinstance(got, torch.Tensor) else got).detach().double().cpu()
    ref = torch.as_tensor(np.asarray(ref) if not isinstance(ref, torch.Tensor) else ref).detach().double().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def to_dev(batch):
    return {k: v.to(dev()) for k, v in batch.items()}


def near_tie_ok(pred, pred_ref, logits):
    """Mismatching argmax pixels must be fp32 near-ties of the two top logits."""
    mism = pred != pred_ref
    if not mism.any():
        return True
    top2 = logits.permute(0, 2, 3, 1)[mism].topk(2, dim=-1).values
    return bool(((top2[:, 0] - top2[:, 1]).abs() < 1e-4).all())


@pytest.mark.parametrize("name", list(MTAN_CASES))
@pytest.mark.parametrize("precision", ["tc_3xtf32", "fp32_ffma"])
def test_mtan_step_vs_reference_golden(name, precision):
    from vision_mtl_b200 import ops
    from vision_mtl_b200.lit_module import MTLModule
    from vision_mtl_b200.models.mtan_model import MTANMiniUnet

    hid, first, levels, B, H, W, C = MTAN_CASES[name]
    g = np.load(os.path.join(GOLDEN, "mtan.npz"))
    old = ops.default_gate_precision
    ops.default_gate_precision = precision
    try:
        net = MTANMiniUnet(3, {"depth": 1, "segm": C}, hid, first, levels)
        sd = FX.fill_state_dict(net.state_dict())
        net.load_state_dict(sd)
        net.to(dev()).to(memory_format=torch.channels_last)
        module = MTLModule(net, num_classes=C, device=dev())
        batch = to_dev(FX.image_batch(B, H, W, C, name))
        net.train()
        loss = module.training_step(batch, 0)
        loss.backward()
        scal = module.last_step_scalars.cpu().double().numpy()  # loss, acc, jaccard, fbeta, mae
        assert abs(scal[0] - g[f"{name}/losses"][0]) <= TOL * abs(g[f"{name}/losses"][0])
        assert abs(scal[4] - g[f"{name}/mae"][0]) <= TOL * abs(g[f"{name}/mae"][0])
        # confusion matrix from the fused head+argmax vs the reference's predictions
        pred_ref = torch.from_numpy(g[f"{name}/preds"].astype(np.int64))
        cm_ref = MN.confusion_matrix(pred_ref.numpy(), batch["mask"].cpu().numpy(), C)
        cm = module.last_confusion.cpu().numpy()
        if not np.array_equal(cm, cm_ref):
            logits_ref = torch.from_numpy(g[f"{name}/segm_logits"])
            assert np.abs(cm - cm_ref).sum() <= 4, "confusion differs by more than near-tie pixels"
            assert near_tie_ok(torch.from_numpy(g[f"{name}/preds"].astype(np.int64)), pred_ref, logits_ref)
        ref_m = MN.all_seg_metrics(cm)
        np.testing.assert_allclose(scal[1:4], [ref_m["accuracy"], ref_m["jaccard_index"], ref_m["fbeta_score"]], rtol=1e-6)
        # gradients and BN buffers: fingerprints of every parameter vs the reference's
        worst = 0.0
        for k, p in net.named_parameters():
            ref = g[f"{name}/grad/{k}"]
            got = FX.summarize(p.grad)
            worst = max(worst, float(np.abs(got - ref).max() / max(ref[1], 1e-12)))
        assert worst <= 3e-4, f"worst gradient fingerprint deviation {worst:.3e} (relative to the grad norm)"
        for k, b in net.named_buffers():
            if "num_batches_tracked" in k:
                assert int(b) == 1
            else:
                np.testing.assert_allclose(FX.summarize(b.float()), g[f"{name}/buf/{k}"], rtol=1e-4, atol=1e-6, err_msg=k)
        # full logits through the API-compatible forward (eval mode, running stats just updated)
        net.eval()
        with torch.no_grad():
            raw_e = net(batch["img"])
        np.testing.assert_allclose(FX.summarize(raw_e["segm"], 16), g[f"{name}/eval_segm"], rtol=3e-4, atol=2e-4)
        np.testing.assert_allclose(FX.summarize(raw_e["depth"], 16), g[f"{name}/eval_depth"], rtol=3e-4, atol=2e-4)
    finally:
        ops.default_gate_precision = old


def test_mtan_full_gradients_vs_oracle():
    """Every gradient tensor (not just fingerprints) against the CPU oracle, Cityscapes classes."""
    from vision_mtl_b200.lit_module import MTLModule
    from vision_mtl_b200.models.mtan_model import MTANMiniUnet

    C, B, H, W = 19, 2, 32, 64
    net = MTANMiniUnet(3, {"depth": 1, "segm": C}, 128, 32, 3)
    sd = FX.fill_state_dict(net.state_dict(), salt=3)
    net.load_state_dict(sd)
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
         for k, v in sd.items()}
    batch = FX.image_batch(B, H, W, C, "full-grad")
    raw = TP.mtan_forward(p, batch["img"], training=True)
    res = TP.step_losses_and_metrics(raw, batch["mask"], batch["depth"], C)
    res["loss"].backward()

    net.to(dev()).to(memory_format=torch.channels_last).train()
    module = MTLModule(net, num_classes=C, device=dev())
    loss = module.training_step(to_dev(batch), 0)
    loss.backward()
    assert rel_err(loss, res["loss"]) <= TOL
    assert np.abs(module.last_confusion.cpu().numpy() - res["confusion"]).sum() <= 4
    bad = {}
    for k, q in net.named_parameters():
        e = rel_err(q.grad, p[k].grad)
        # conv biases in front of a training-mode BN have analytically zero gradient: compare on the
        # scale of the matching weight gradient instead of their own (noise-level) magnitude
        if k.endswith("bias") and p[k].grad.abs().max() < 1e-5:
            continue
        if e > 2e-4:
            bad[k] = e
    assert not bad, f"gradient mismatches: {bad}"


@pytest.mark.parametrize("name,cw", [("csnet_cw", True), ("csnet_lw", False)])
def test_csnet_step_vs_reference_golden(name, cw):
    from vision_mtl_b200.lit_module import MTLModule
    from vision_mtl_b200.models import CSNet
    from vision_mtl_b200.utils.model_utils import get_model_with_dense_preds

    g = np.load(os.path.join(GOLDEN, "csnet.npz"))
    models = {"depth": get_model_with_dense_preds(1, None, dict(encoder_weights=None)),
              "segm": get_model_with_dense_preds(19, None, dict(encoder_weights=None))}
    net = CSNet(models, channel_wise_stitching=cw)
    net.load_state_dict(FX.fill_state_dict(net.state_dict()))
    net.to(dev()).to(memory_format=torch.channels_last).train()
    batch = to_dev(FX.image_batch(2, 64, 64, 19, name))
    batch["img"] = batch["img"].contiguous(memory_format=torch.channels_last)
    raw = net(batch["img"])
    assert rel_err(raw["segm"], g[f"{name}/segm_logits"]) <= TOL
    assert rel_err(raw["depth"], g[f"{name}/depth_logits"]) <= TOL
    module = MTLModule(net, num_classes=19, device=dev())
    net.load_state_dict({k: v.to(dev()) for k, v in FX.fill_state_dict(net.state_dict()).items()})  # reset BN buffers
    loss = module.training_step(batch, 0)
    loss.backward()
    assert abs(loss.item() - g[f"{name}/losses"][0]) <= TOL * abs(g[f"{name}/losses"][0])
    worst, worst_k = 0.0, None
    for k, p in net.named_parameters():
        key = f"{name}/grad/{k}"
        if key not in g.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        ref = g[key]
        dev_ = float(np.abs(FX.summarize(p.grad) - ref).max() / max(ref[1], 1e-12))
        if dev_ > worst:
            worst, worst_k = dev_, k
    assert worst <= 1e-3, f"worst gradient fingerprint deviation {worst:.3e} at {worst_k}"
    # off-diagonal alphas get exactly-zero gradients in reference mode
    for layer in net.cross_stitch_layers.values():
        gr = layer.weights.grad.cpu()
        assert float(gr[0, 1].abs().max()) == 0.0 and float(gr[1, 0].abs().max()) == 0.0


def test_module_api_compat_paths():
    """postprocess_raw_out / calc_losses / calc_metrics (the reference's unfused call sequence)
    agree with the fused step."""
    from vision_mtl_b200.lit_module import MTLModule
    from vision_mtl_b200.models.mtan_model import MTANMiniUnet

    C = 14
    net = MTANMiniUnet(3, {"depth": 1, "segm": C}, 64, 32, 2)
    net.load_state_dict(FX.fill_state_dict(net.state_dict(), salt=5))
    net.to(dev()).to(memory_format=torch.channels_last).eval()
    module = MTLModule(net, num_classes=C, device=dev())
    batch = to_dev(FX.image_batch(2, 32, 32, C, "api"))
    with torch.no_grad():
        fused = module.fused_losses_and_metrics(batch["img"], batch["mask"], batch["depth"], want_preds=True)
        out = module.postprocess_raw_out(module(batch["img"]))
        losses = module.calc_losses(batch["mask"], batch["depth"], out)
        metrics = module.calc_metrics(batch["mask"], batch["depth"], out)
    assert rel_err(losses["loss"], fused["loss"]) <= TOL
    assert rel_err(metrics["mae"], fused["mae"]) <= TOL
    assert rel_err(out["depth_predictions"], fused["depth_predictions"]) <= TOL
    for k in ("accuracy", "jaccard_index", "fbeta_score"):
        assert abs(float(metrics[k]) - float(fused[k])) <= 1e-3  # near-tie pixels only
    preds = module.predict_step(batch)
    assert preds["segm"].shape == (2, 32, 32) and preds["depth"].shape == (2, 32, 32, 1)
    ep = module.on_predict_epoch_end()
    assert set(ep) == {"predict/loss", "predict/accuracy", "predict/jaccard_index", "predict/fbeta_score", "predict/mae"}
