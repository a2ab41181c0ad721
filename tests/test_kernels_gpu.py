"""GPU parity tests: every C-ABI kernel against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): confusion matrices bit-exact; losses, gradients and depth
metrics within 1e-4 relative in fp32.  The relative error is taken tensor-wise:
max|got - ref| <= TOL * max|ref|.
"""
import copy
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import kernels_ref as K
from oracle import metrics_np as MN

pytestmark = pytest.mark.gpu
TOL = 1e-4


def dev():
    return torch.device("cuda:0")


def assert_rel(got, ref, tol=TOL, what=""):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    scale = max(ref.abs().max().item(), 1e-30)
    err = (got - ref).abs().max().item() if got.numel() else 0.0
    assert err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e})"


def to_cl(x):
    return x.to(dev()).contiguous(memory_format=torch.channels_last)


# ----------------------------------------------------------------------------- cross-stitch
XS_SHAPES = [(2, 2, 16, 8, 12), (2, 3, 24, 5, 7), (3, 2, 40, 4, 4), (2, 1, 1072, 2, 3), (2, 2, 32, 16, 16),
             (4, 1, 8, 3, 5)]


@pytest.mark.parametrize("T,B,C,H,W", XS_SHAPES)
@pytest.mark.parametrize("channel_wise", [False, True])
@pytest.mark.parametrize("mode", ["reference_diag", "full_mix"])
def test_xstitch_fwd_bwd(T, B, C, H, W, channel_wise, mode):
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(11)
    x = torch.randn(T, B, C, H, W, generator=g)
    alpha = torch.rand(T, T, C, generator=g) if channel_wise else torch.rand(T, T, generator=g)
    dy = torch.randn(T, B, C, H, W, generator=g)
    # oracle (fp32 on CPU, reference layout)
    xr = x.clone().requires_grad_(True)
    ar = alpha.clone().requires_grad_(True)
    fn = K.xstitch_reference_diag if mode == "reference_diag" else K.xstitch_full_mix
    yr = fn(ar, xr)
    yr.backward(dy)
    # product
    xs = [to_cl(x[t]).requires_grad_(True) for t in range(T)]
    ad = alpha.to(dev()).requires_grad_(True)
    ys = ops.cross_stitch(xs, ad, mode)
    torch.autograd.backward(ys, [to_cl(dy[t]) for t in range(T)])
    for t in range(T):
        if mode == "reference_diag":  # a single fp32 multiply: bit-exact
            assert torch.equal(ys[t].cpu(), yr[t].detach()), "diag forward must be bit-exact"
        else:
            assert_rel(ys[t], yr[t], what=f"y[{t}]")
        assert_rel(xs[t].grad, xr.grad[t], what=f"dx[{t}]")
    assert_rel(ad.grad, ar.grad, what="dalpha")
    if mode == "reference_diag":
        off = ~torch.eye(T, dtype=torch.bool)
        assert (ad.grad.cpu()[off] == 0).all(), "off-diagonal alpha grads must be exact zeros"


# ----------------------------------------------------------------------------- confusion / metrics
@pytest.mark.parametrize("P", [0, 1, 15, 16, 17, 1000, 128 * 256 * 2 + 3])
@pytest.mark.parametrize("C", [13, 14, 19])
@pytest.mark.parametrize("u8", [False, True])
def test_confusion_bit_exact(P, C, u8):
    from vision_mtl_b200 import ops

    rng = np.random.default_rng(11 + P + C)
    pred = rng.integers(0, C, size=P)
    target = rng.integers(0, C, size=P)
    if P > 20:
        target[::7] = -100  # ignored
        target[3] = C + 5  # out of range -> dropped
    ref = MN.confusion_matrix(pred, target, C, ignore_index=-100)
    pt = torch.from_numpy(pred).to(dev())
    if u8:
        pt = pt.to(torch.uint8)
    conf = ops.confusion_accumulate(pt, torch.from_numpy(target).to(dev()), C)
    assert np.array_equal(conf.cpu().numpy(), ref)
    # accumulation (+=) semantic
    conf = ops.confusion_accumulate(pt, torch.from_numpy(target).to(dev()), C, conf=conf)
    assert np.array_equal(conf.cpu().numpy(), 2 * ref)


def test_confusion_unaligned_views():
    from vision_mtl_b200 import ops

    C, P = 19, 5000
    rng = np.random.default_rng(3)
    pred = torch.from_numpy(rng.integers(0, C, size=P + 1)).to(dev())
    target = torch.from_numpy(rng.integers(0, C, size=P + 1)).to(dev())
    got = ops.confusion_accumulate(pred[1:], target[1:], C)  # 8-byte aligned only
    ref = MN.confusion_matrix(pred[1:].cpu().numpy(), target[1:].cpu().numpy(), C)
    assert np.array_equal(got.cpu().numpy(), ref)
    got8 = ops.confusion_accumulate(pred[1:].to(torch.uint8)[3:], target[4:], C)
    ref8 = MN.confusion_matrix(pred[4:].cpu().numpy(), target[4:].cpu().numpy(), C)
    assert np.array_equal(got8.cpu().numpy(), ref8)


@pytest.mark.parametrize("C", [14, 19])
def test_seg_metrics_from_confusion(C):
    from vision_mtl_b200 import ops

    rng = np.random.default_rng(5)
    cm = rng.integers(0, 5000, size=(C, C))
    cm[:, 3] = 0
    cm[3, :] = 0  # an absent class
    m = ops.seg_metrics(torch.from_numpy(cm).to(dev())).cpu().numpy()
    ref = MN.all_seg_metrics(cm)
    np.testing.assert_allclose(m, [ref["accuracy"], ref["jaccard_index"], ref["fbeta_score"]], rtol=1e-6)


@pytest.mark.parametrize("P", [1, 7, 4096, 128 * 256 + 5])
def test_depth_err_sums(P):
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(P)
    pred = torch.rand(P, generator=g)
    target = torch.rand(P, generator=g) * 0.5
    target[torch.rand(P, generator=g) < 0.2] = 0.0
    out = ops.depth_error_sums(pred.to(dev()), target.to(dev())).cpu()
    pd, td = pred.double(), target.double()
    m = td > 1e-3
    ref = torch.tensor([P, (pd - td).abs().sum(), m.sum(), ((pd - td).abs()[m] / td[m]).sum()], dtype=torch.float64)
    assert out[0].item() == P and out[2].item() == m.sum().item()
    assert_rel(out, ref, tol=1e-6, what="depth sums")


# ----------------------------------------------------------------------------- CE on logits
@pytest.mark.parametrize("B,C,H,W", [(2, 19, 16, 32), (1, 13, 7, 9), (3, 14, 8, 8), (2, 19, 128, 256)])
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("ignore", [False, True])
def test_ce_logits(B, C, H, W, layout, ignore):
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(11)
    logits = torch.randn(B, C, H, W, generator=g) * 3
    target = torch.randint(0, C, (B, H, W), generator=g)
    if ignore:
        target[torch.rand(B, H, W, generator=g) < 0.3] = -100
    lr = logits.clone().requires_grad_(True)
    loss_r = K.cross_entropy(lr, target)
    (loss_r * 1.7).backward()
    ld = logits.to(dev())
    if layout == "nhwc":
        ld = ld.contiguous(memory_format=torch.channels_last)
    ld.requires_grad_(True)
    conf = torch.zeros(C, C, dtype=torch.int64, device=dev())
    loss, pred = ops.cross_entropy_logits(ld, target.to(dev()), -100, conf, True)
    (loss * 1.7).backward()
    assert_rel(loss, loss_r, what="loss")
    assert_rel(ld.grad, lr.grad, what="dlogits")
    pred_r = K.segm_predictions(logits)
    # random fp32 logits: argmax(logits) == argmax(softmax(logits)) unless a near tie
    mism = pred.cpu().long() != pred_r
    if mism.any():
        top2 = logits.permute(0, 2, 3, 1)[mism].topk(2, dim=-1).values
        assert ((top2[:, 0] - top2[:, 1]).abs() < 1e-5).all()
    ref_cm = MN.confusion_matrix(pred.cpu().numpy(), target.numpy(), C, ignore_index=-100)
    assert np.array_equal(conf.cpu().numpy(), ref_cm)


# ----------------------------------------------------------------------------- fused heads
@pytest.mark.parametrize("B,C,H,W", [(2, 19, 16, 32), (1, 13, 5, 9), (2, 14, 32, 32), (2, 19, 128, 256),
                                     (1, 27, 20, 20), (1, 13, 15, 9), (3, 16, 33, 47), (1, 32, 24, 24),
                                     # 5-6 and 7 tiles per CTA: every barrier ring (mod 2, 3, 4, 6, 8) wraps
                                     (1, 19, 300, 321), (1, 19, 331, 400)])
@pytest.mark.parametrize("ignore", [False, True])
def test_head_ce(B, C, H, W, ignore):
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(11)
    feat = torch.randn(B, 32, H, W, generator=g)
    head = torch.nn.Conv2d(32, C, 1)
    target = torch.randint(0, C, (B, H, W), generator=g)
    if ignore:
        target[torch.rand(B, H, W, generator=g) < 0.25] = -100
    fr = feat.clone().requires_grad_(True)
    logits_r = K.head_project(fr, head.weight, head.bias)
    loss_r = K.cross_entropy(logits_r, target)
    loss_r.backward()
    fd = to_cl(feat).requires_grad_(True)
    wd = head.weight.detach().to(dev()).requires_grad_(True)
    bd = head.bias.detach().to(dev()).requires_grad_(True)
    conf = torch.zeros(C, C, dtype=torch.int64, device=dev())
    loss, pred = ops.head_cross_entropy(fd, wd, bd, target.to(dev()), -100, conf, True)
    loss.backward()
    assert_rel(loss, loss_r, what="loss")
    assert_rel(fd.grad, fr.grad, what="dfeat")
    assert_rel(wd.grad, head.weight.grad, what="dW")
    assert_rel(bd.grad, head.bias.grad, what="db")
    pred_r = K.segm_predictions(logits_r.detach())
    mism = pred.cpu().long() != pred_r
    if mism.any():
        top2 = logits_r.detach().permute(0, 2, 3, 1)[mism].topk(2, dim=-1).values
        assert ((top2[:, 0] - top2[:, 1]).abs() < 1e-5).all()
    assert np.array_equal(conf.cpu().numpy(),
                          MN.confusion_matrix(pred.cpu().numpy(), target.numpy(), C, ignore_index=-100))


def test_head_ce_exact_arithmetic_confusion():
    """Dyadic features/weights -> every fp32 dot product is exact in any order, so the fused
    argmax must reproduce the reference confusion matrix bit for bit."""
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(7)
    B, C, H, W = 2, 19, 24, 40
    feat = torch.randint(-8, 9, (B, 32, H, W), generator=g).float() / 4
    weight = torch.randint(-8, 9, (C, 32, 1, 1), generator=g).float() / 8
    bias = torch.randint(-8, 9, (C,), generator=g).float() / 16
    target = torch.randint(0, C, (B, H, W), generator=g)
    logits = K.head_project(feat, weight, bias)
    pred_r = torch.argmax(logits, dim=1)  # exact ties resolve to the lowest index in both
    ref = MN.confusion_matrix(pred_r.numpy(), target.numpy(), C)
    conf = torch.zeros(C, C, dtype=torch.int64, device=dev())
    _, pred = ops.head_cross_entropy(to_cl(feat), weight.to(dev()), bias.to(dev()), target.to(dev()), -100, conf, True)
    assert torch.equal(pred.cpu().long(), pred_r)
    assert np.array_equal(conf.cpu().numpy(), ref)


@pytest.mark.parametrize("B,H,W", [(2, 16, 32), (1, 5, 9), (2, 128, 256)])
@pytest.mark.parametrize("cin", [32, 1])
def test_head_silog(B, H, W, cin):
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(11)
    target = torch.rand(B, H, W, 1, generator=g) * 0.5
    target[torch.rand(B, H, W, 1, generator=g) < 0.2] = 0.0
    if cin == 32:
        feat = torch.randn(B, 32, H, W, generator=g)
        head = torch.nn.Conv2d(32, 1, 1)
        fr = feat.clone().requires_grad_(True)
        zr = K.head_project(fr, head.weight, head.bias)
    else:
        feat = torch.randn(B, 1, H, W, generator=g)
        fr = feat.clone().requires_grad_(True)
        zr = fr
    pr = K.depth_predictions(zr)
    loss_r = K.silog(pr, target)
    (loss_r * 0.7).backward()
    fd = (to_cl(feat) if cin == 32 else feat.to(dev())).requires_grad_(True)
    if cin == 32:
        wd = head.weight.detach().to(dev()).requires_grad_(True)
        bd = head.bias.detach().to(dev()).requires_grad_(True)
    else:
        wd = bd = None
    silog, mae, absrel, pred = ops.head_silog(fd, wd, bd, target.to(dev()), 1e-3, True)
    (silog * 0.7).backward()
    assert_rel(silog, loss_r, what="silog")
    assert_rel(pred, pr, what="pred")
    assert_rel(mae, K.depth_mae(pr.detach(), target), what="mae")
    assert_rel(absrel, K.depth_abs_rel(pr.detach(), target), what="abs_rel")
    assert_rel(fd.grad, fr.grad, what="dfeat")
    if cin == 32:
        assert_rel(wd.grad, head.weight.grad, what="dw")
        assert_rel(bd.grad, head.bias.grad, what="db")


def test_silog_loss_module_mask_and_saturation():
    """``SILogLoss.forward`` on sigmoid predictions (the API path of losses.py:14-36): explicit ``mask`` argument
    (losses.py:29-33), and saturated logits -- fp32 sigmoid(-100) is a denormal the reference still takes the log of."""
    from vision_mtl_b200 import ops
    from vision_mtl_b200.losses import SILogLoss

    g = torch.Generator().manual_seed(5)
    B, H, W = 2, 12, 20
    target = torch.rand(B, H, W, 1, generator=g) * 0.5 + 0.01
    mask = torch.rand(B, H, W, 1, generator=g) < 0.6
    z = torch.randn(B, H, W, 1, generator=g)
    pr = torch.sigmoid(z).requires_grad_(True)
    gref = torch.log(pr[mask]) - torch.log(target[mask])  # the reference's arithmetic
    loss_r = 10 * torch.sqrt(torch.var(gref) + 0.15 * torch.mean(gref) ** 2)
    loss_r.backward()
    pd = torch.sigmoid(z).to(dev()).requires_grad_(True)
    loss = SILogLoss()(pd, target.to(dev()), mask=mask.to(dev()))
    loss.backward()
    assert_rel(loss, loss_r, what="masked silog")
    assert_rel(pd.grad, pr.grad, what="masked silog d/dpred")
    # ... and against the output of the UNMODIFIED reference SILogLoss(mask=...) (tests/golden/kernels.npz)
    from oracle import fixtures as FX
    from oracle.make_golden import silog_explicit_mask

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "kernels.npz"))
    bg = FX.image_batch(2, 8, 16, 19, "silog")
    lg = FX.tensor((2, 8, 16, 1), "silog/logit", 2.0).to(dev()).requires_grad_(True)
    loss_g = SILogLoss()(torch.sigmoid(lg), bg["depth"].to(dev()), mask=silog_explicit_mask(bg["depth"]).to(dev()))
    loss_g.backward()
    assert_rel(loss_g, torch.tensor(float(gold["silog_masked/loss"])), what="masked silog vs reference golden")
    assert_rel(lg.grad.cpu(), torch.from_numpy(gold["silog_masked/dlogit"]), what="masked silog dlogit vs reference golden")
    # saturation: logits down to -100 (sigmoid flushes to 0 in the fast form) and up to +40 (sigmoid == 1.0f)
    zs = torch.linspace(-100.0, 40.0, B * H * W).reshape(B, 1, H, W)
    ts = torch.full((B, H, W, 1), 0.25)
    gs = torch.nn.functional.logsigmoid(zs.double()).reshape(-1) - torch.log(ts.double()).reshape(-1)
    ref = 10 * torch.sqrt(torch.var(gs) + 0.15 * torch.mean(gs) ** 2)
    zd = zs.to(dev()).requires_grad_(True)
    silog, _, _, _ = ops.head_silog(zd, None, None, ts.to(dev()), 1e-3, False)
    silog.backward()
    assert torch.isfinite(silog) and torch.isfinite(zd.grad).all()
    assert_rel(silog, ref.float(), what="saturated silog")


# ----------------------------------------------------------------------------- MTAN gate
GATE_SHAPES = [(2, 32, 16, 24), (1, 64, 9, 13), (2, 128, 8, 8), (1, 256, 4, 8), (4, 32, 64, 64),
               (3, 128, 20, 24), (2, 256, 12, 10), (1, 192, 16, 20),
               # several 128-row tiles per CTA (148 CTAs): the persistent pipelines wrap their stage / TMEM
               # buffers and mbarrier phases (3-4, 3-4 and 2-3 tiles per CTA)
               (2, 32, 128, 256), (2, 64, 128, 256), (1, 128, 160, 256)]


def _gate_case(B, N, H, W, seed=11):
    g = torch.Generator().manual_seed(seed)
    h = torch.relu(torch.randn(B, 128, H, W, generator=g))
    s = torch.relu(torch.randn(B, N, H, W, generator=g))
    conv = torch.nn.Conv2d(128, N, 1)
    bn = torch.nn.BatchNorm2d(N)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g)
        bn.bias.uniform_(-0.5, 0.5, generator=g)
        bn.running_mean.uniform_(-0.2, 0.2, generator=g)
        bn.running_var.uniform_(0.5, 1.5, generator=g)
    dy = torch.randn(B, N, H, W, generator=g)
    return h, s, conv, bn, dy


@pytest.mark.parametrize("B,N,H,W", GATE_SHAPES)
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("precision", ["fp32_ffma", "tc_3xtf32"])
def test_gate_fwd_bwd(B, N, H, W, training, precision):
    from vision_mtl_b200 import ops

    h, s, conv, bn, dy = _gate_case(B, N, H, W)
    hr, sr = h.clone().requires_grad_(True), s.clone().requires_grad_(True)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    yr = K.gate_forward(hr, sr, conv.weight, conv.bias, bn.weight, bn.bias, rm, rv, training)
    yr.backward(dy)

    hd, sd = to_cl(h).requires_grad_(True), to_cl(s).requires_grad_(True)
    p = [t.detach().clone().to(dev()).requires_grad_(True) for t in (conv.weight, conv.bias, bn.weight, bn.bias)]
    rmd, rvd = bn.running_mean.clone().to(dev()), bn.running_var.clone().to(dev())
    y = ops.attention_gate(hd, sd, p[0], p[1], p[2], p[3], rmd, rvd, training, 0.1, 1e-5, precision)
    y.backward(to_cl(dy))
    assert_rel(y, yr, what="y")
    assert_rel(sd.grad, sr.grad, what="ds")
    assert_rel(hd.grad, hr.grad, what="dh")
    assert_rel(p[0].grad, conv.weight.grad, what="dW")
    assert_rel(p[2].grad, bn.weight.grad, what="dgamma")
    assert_rel(p[3].grad, bn.bias.grad, what="dbeta")
    if training:
        assert_rel(rmd, rm, what="running_mean")
        assert_rel(rvd, rv, what="running_var")
        # d bias is analytically zero under batch statistics; compare on the scale of dbeta
        assert p[1].grad.abs().max().item() <= 1e-4 * max(bn.bias.grad.abs().max().item(), 1e-30) + 1e-6
    else:
        assert_rel(p[1].grad, conv.bias.grad, what="dbias")
        with torch.no_grad():  # inference: the tensor-core path fuses the whole gate into one pass
            y2 = ops.attention_gate(hd, sd, p[0], p[1], p[2], p[3], rmd, rvd, False, 0.1, 1e-5, precision)
        assert_rel(y2, yr, what="y (no_grad eval)")


# ----------------------------------------------------------------------------- BatchNorm (+ReLU, +max-pool)
BN_SHAPES = [(2, 32, 16, 24), (1, 64, 9, 13), (3, 128, 20, 24), (2, 256, 12, 10), (1, 512, 4, 8), (2, 4, 8, 8),
             (4, 32, 64, 64), (2, 128, 128, 256)]


@pytest.mark.parametrize("B,C,H,W", BN_SHAPES)
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("relu,pool", [(True, False), (False, False), (True, True)])
def test_bn_relu_pool_fwd_bwd(B, C, H, W, training, relu, pool):
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(17)
    x = torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g)
        bn.bias.uniform_(-0.5, 0.5, generator=g)
        bn.running_mean.uniform_(-0.2, 0.6, generator=g)
        bn.running_var.uniform_(1.5, 3.0, generator=g)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    dy = torch.randn(B, C, Ho, Wo, generator=g)
    xr = x.clone().requires_grad_(True)
    gr, br = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    yr = K.bn_relu(xr, gr, br, rm, rv, training, relu, pool)
    yr.backward(dy)

    bnd = torch.nn.BatchNorm2d(C).to(dev())
    bnd.load_state_dict(bn.state_dict())
    bnd.train(training)
    xd = to_cl(x).requires_grad_(True)
    assert ops.bn_supported(bnd, xd)
    y = ops.batch_norm_relu(xd, bnd, relu=relu, pool=pool)
    y.backward(to_cl(dy))
    assert_rel(y, yr, what="y")
    assert_rel(xd.grad, xr.grad, what="dx")
    assert_rel(bnd.weight.grad, gr.grad, what="dgamma")
    assert_rel(bnd.bias.grad, br.grad, what="dbeta")
    if training:
        assert_rel(bnd.running_mean, rm, what="running_mean")
        assert_rel(bnd.running_var, rv, what="running_var")
        assert int(bnd.num_batches_tracked) == 1
    else:
        assert torch.equal(bnd.running_mean.cpu(), bn.running_mean) and int(bnd.num_batches_tracked) == 0


# ----------------------------------------------------------------------------- bilinear x2 up-sampling + cat
@pytest.mark.parametrize("B,C1,Cx,Hi,Wi", [(2, 32, 128, 8, 16), (1, 64, 128, 1, 1), (3, 8, 4, 5, 7), (2, 128, 128, 16, 32),
                                           (1, 4, 12, 2, 3), (2, 32, 128, 64, 128)])
def test_upsample2_bilinear_cat(B, C1, Cx, Hi, Wi):
    """``torch.cat((first, nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)(x)), 1)`` as the
    decoder attention modules run it (mtan_model.py:125,143-145) against the fused op: output and both gradients."""
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(17)
    first = to_cl(torch.randn(B, C1, 2 * Hi, 2 * Wi, generator=g))
    x = to_cl(torch.randn(B, Cx, Hi, Wi, generator=g))
    dy = to_cl(torch.randn(B, C1 + Cx, 2 * Hi, 2 * Wi, generator=g))
    fr, xr = first.clone().requires_grad_(True), x.clone().requires_grad_(True)
    up = torch.nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
    yr = torch.cat((fr, up(xr)), dim=1)
    yr.backward(dy)
    fd, xd = first.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y = ops.upsample2_cat(fd, xd)
    y.backward(dy)
    assert y.is_contiguous(memory_format=torch.channels_last) or min(y.shape[1:]) == 1
    assert_rel(y, yr, tol=2e-6, what="y")
    assert torch.equal(fd.grad, fr.grad)
    assert_rel(xd.grad, xr.grad, tol=2e-6, what="dx")


# ----------------------------------------------------------------------------- gate with the folded hidden layer
@pytest.mark.parametrize("pool", [False, True])
@pytest.mark.parametrize("cin,cout,k,H,W", [(8, 32, 3, 16, 24), (12, 128, 1, 10, 14), (16, 64, 3, 9, 13)])
def test_conv_bias_folded_into_batchnorm(cin, cout, k, H, W, pool):
    """conv (with bias) -> BatchNorm2d (batch statistics) -> ReLU [-> MaxPool2d(2)] as the reference runs it
    (mtan_model.py:77-81, :141-142, :165-167) against the convolution WITHOUT its bias + the BatchNorm kernels
    that take the bias: same output, same running statistics, same gradients; the bias gradient is the
    round-off-sized number it analytically is."""
    from vision_mtl_b200 import ops

    if pool and (H % 2 or W % 2):
        pytest.skip("pooled case uses even sizes")
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(cin, cout, k, padding=k // 2).to(dev())
    with torch.no_grad():
        conv.bias.mul_(20.0).add_(1.0)  # a bias far from zero: a kernel that forgot it would show
    bn = torch.nn.BatchNorm2d(cout).to(dev())
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    conv_r, bn_r = copy.deepcopy(conv), copy.deepcopy(bn)
    x = to_cl(torch.randn(3, cin, H, W))
    xr = x.clone().requires_grad_(True)
    xd = x.clone().requires_grad_(True)
    act, mp = torch.nn.ReLU(), torch.nn.MaxPool2d(2)
    yr = act(bn_r(conv_r(xr)))
    yr = mp(yr) if pool else yr
    dy = torch.randn_like(yr) + 0.3
    yr.backward(dy)
    y0, cb = ops.conv_without_bias(conv, xd, bn)
    assert cb is conv.bias
    y = ops.batch_norm_relu(y0, bn, relu=True, pool=pool, conv_bias=cb)
    y.backward(dy)
    assert_rel(y, yr, what="y")
    assert_rel(bn.running_mean, bn_r.running_mean, what="running_mean")
    assert_rel(bn.running_var, bn_r.running_var, what="running_var")
    assert int(bn.num_batches_tracked) == 1
    assert_rel(xd.grad, xr.grad, what="dx")
    assert_rel(conv.weight.grad, conv_r.weight.grad, what="dW")
    assert_rel(bn.weight.grad, bn_r.weight.grad, what="dgamma")
    assert_rel(bn.bias.grad, bn_r.bias.grad, what="dbeta")
    # analytically zero: both sides hold round-off on the scale of dbeta
    scale = float(bn_r.bias.grad.abs().max())
    assert float(conv.bias.grad.abs().max()) <= 1e-4 * scale and float(conv_r.bias.grad.abs().max()) <= 1e-4 * scale


@pytest.mark.parametrize("B,N,H,W", [(2, 32, 16, 24), (1, 64, 9, 13), (2, 128, 8, 8), (1, 256, 4, 8), (2, 32, 128, 256),
                                     (1, 128, 160, 256)])
@pytest.mark.parametrize("training", [True, False])
def test_gate_folded_hidden_layer(B, N, H, W, training):
    """relu(bn1(c)) is folded into the gate kernels: against conv1-output -> bn1 -> relu -> gate
    sequenced with ATen ops, forward, every gradient and both BatchNorms' running statistics."""
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(23)
    c = torch.randn(B, 128, H, W, generator=g) * 1.3 + 0.2
    s = torch.relu(torch.randn(B, N, H, W, generator=g))
    dy = torch.randn(B, N, H, W, generator=g)
    bn1, conv2, bn2 = torch.nn.BatchNorm2d(128), torch.nn.Conv2d(128, N, 1), torch.nn.BatchNorm2d(N)
    with torch.no_grad():
        for bn in (bn1, bn2):
            bn.weight.uniform_(0.5, 1.5, generator=g)
            bn.bias.uniform_(-0.5, 0.5, generator=g)
            bn.running_mean.uniform_(-0.2, 0.2, generator=g)
            bn.running_var.uniform_(0.5, 1.5, generator=g)
    mods = torch.nn.ModuleList([bn1, conv2, bn2])
    ref = type(mods)([type(m)(*a) for m, a in ((bn1, (128,)), (conv2, (128, N, 1)), (bn2, (N,)))])
    ref.load_state_dict(mods.state_dict())
    ref.train(training)
    cr, sr = c.clone().requires_grad_(True), s.clone().requires_grad_(True)
    yr = sr * torch.sigmoid(ref[2](ref[1](torch.relu(ref[0](cr)))))
    yr.backward(dy)

    mods.to(dev()).train(training)
    cd, sd = to_cl(c).requires_grad_(True), to_cl(s).requires_grad_(True)
    assert ops.folded_gate_supported(cd, sd, mods[0], mods[2])
    y = ops.attention_gate_folded(cd, mods[0], sd, mods[1], mods[2])
    y.backward(to_cl(dy))
    assert_rel(y, yr, what="y")
    assert_rel(sd.grad, sr.grad, what="ds")
    assert_rel(cd.grad, cr.grad, what="dc")
    for (k, p), (_, q) in zip(mods.named_parameters(), ref.named_parameters()):
        if k == "1.bias" and training:  # analytically zero under batch statistics
            assert p.grad.abs().max().item() <= 1e-4 * max(ref[2].bias.grad.abs().max().item(), 1e-30) + 1e-6
        else:
            assert_rel(p.grad, q.grad, what=k)
    for (k, b), (_, q) in zip(mods.named_buffers(), ref.named_buffers()):
        if "num_batches" in k:
            assert int(b) == int(q), k
        else:
            assert_rel(b, q, what=k)


# ----------------------------------------------------------------------------- pad + cat + stitch / upsample + stitch
CAT_CASES = [  # (T, B, Cs, Ho, Wo, Cx, Hi, Wi, up2)
    (2, 2, 112, 8, 16, 960, 4, 8, False),   # decoder block 0 of the csnet config (1072 channels, x zero-padded)
    (2, 2, 40, 16, 32, 256, 8, 16, False),
    (2, 1, 24, 12, 20, 128, 6, 10, False),
    (2, 2, 16, 9, 13, 64, 5, 6, False),     # odd sizes: asymmetric centred padding
    (3, 1, 8, 6, 6, 12, 6, 6, False),       # same size: pure cat
    (2, 2, 0, 16, 24, 32, 8, 12, True),     # last decoder block: nearest x2
    (3, 1, 0, 8, 8, 16, 4, 4, True),
]


@pytest.mark.parametrize("T,B,Cs,Ho,Wo,Cx,Hi,Wi,up2", CAT_CASES)
@pytest.mark.parametrize("cw", [True, False])
@pytest.mark.parametrize("mode", ["reference_diag", "full_mix"])
def test_xstitch_cat_fwd_bwd(T, B, Cs, Ho, Wo, Cx, Hi, Wi, up2, cw, mode):
    """The gather-stitch kernels against the reference's op sequence (F.pad + torch.cat / nearest interpolate, then
    the stitch oracle): outputs, input gradients and alpha gradients."""
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(29)
    C = Cs + Cx
    xs = [torch.randn(B, Cx, Hi, Wi, generator=g) for _ in range(T)]
    skips = [torch.randn(B, Cs, Ho, Wo, generator=g) for _ in range(T)] if Cs else []
    alpha = torch.rand(T, T, C, generator=g) if cw else torch.rand(T, T, generator=g)
    dys = [torch.randn(B, C, Ho, Wo, generator=g) for _ in range(T)]
    xr = [x.clone().requires_grad_(True) for x in xs]
    sr = [s.clone().requires_grad_(True) for s in skips]
    ar = alpha.clone().requires_grad_(True)
    if up2:
        feats = [torch.nn.functional.interpolate(x, scale_factor=2, mode="nearest") for x in xr]
    else:
        dh, dw = Ho - Hi, Wo - Wi
        feats = [torch.cat([s, torch.nn.functional.pad(x, [dw // 2, dw - dw // 2, dh // 2, dh - dh // 2])], dim=1)
                 for x, s in zip(xr, sr)]
    stacked = torch.stack(feats)
    yr = K.xstitch_reference_diag(ar, stacked) if mode == "reference_diag" else K.xstitch_full_mix(ar, stacked)
    yr.backward(torch.stack(dys))

    xd = [to_cl(x).requires_grad_(True) for x in xs]
    sd = [to_cl(s).requires_grad_(True) for s in skips]
    ad = alpha.to(dev()).requires_grad_(True)
    ys = ops.cross_stitch_cat(sd, xd, ad, mode, up2=up2)
    torch.autograd.backward(ys, [to_cl(d) for d in dys])
    for t in range(T):
        if mode == "reference_diag":
            assert torch.equal(ys[t].cpu(), yr[t].detach()), "one multiply per element: bit-exact"
        else:
            assert_rel(ys[t], yr[t], what=f"y[{t}]")
        assert_rel(xd[t].grad, xr[t].grad, what=f"dx[{t}]")
        if Cs:
            assert_rel(sd[t].grad, sr[t].grad, what=f"dskip[{t}]")
    assert_rel(ad.grad, ar.grad, what="dalpha")
    if mode == "reference_diag":
        off = ~torch.eye(T, dtype=torch.bool)
        assert float(ad.grad.cpu()[off].abs().max()) == 0.0
