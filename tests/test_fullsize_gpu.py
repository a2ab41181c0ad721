"""BASELINE.json full sizes (Cityscapes-shaped batch 32, 128x256: P = M = 1 048 576 pixels).

Two kinds of checks at the size the bench runs:
* size-independent properties that need no oracle (conservation of the confusion matrix, linearity of the
  cross-stitch and of the gate backward in dy, determinism of repeated launches);
* the oracle itself on the largest site of every op (it finishes in seconds on the host cores).
"""
import numpy as np
import pytest
import torch

from oracle import kernels_ref as K
from oracle import metrics_np as MN

pytestmark = pytest.mark.gpu
TOL = 1e-4
B, H, W, C = 32, 128, 256, 19


def dev():
    return torch.device("cuda:0")


def rel(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)


def cl(x):
    return x.to(dev()).contiguous(memory_format=torch.channels_last)


def test_confusion_conservation_full_size():
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(3)
    target = torch.randint(0, C, (B, H, W), generator=g)
    target[torch.rand(B, H, W, generator=g) < 0.1] = -100
    pred = torch.randint(0, C, (B, H, W), generator=g)
    for p in (pred.to(dev()), pred.to(dev()).to(torch.uint8)):
        conf = ops.confusion_accumulate(p, target.to(dev()), C)
        valid = target != -100
        assert int(conf.sum()) == int(valid.sum())                                   # every valid pixel once
        assert torch.equal(conf.sum(1).cpu(), torch.bincount(target[valid], minlength=C))  # rows = targets
        assert torch.equal(conf.sum(0).cpu(), torch.bincount(pred[valid], minlength=C))    # cols = predictions
        assert np.array_equal(conf.cpu().numpy(), MN.confusion_matrix(pred.numpy(), target.numpy(), C, ignore_index=-100))


def test_xstitch_full_size_oracle_and_linearity():
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(5)
    Cc = 32
    xs = [torch.randn(B, Cc, H, W, generator=g) for _ in range(2)]
    alpha = torch.rand(2, 2, Cc, generator=g)
    xd = [cl(x).requires_grad_(True) for x in xs]
    ad = alpha.to(dev()).requires_grad_(True)
    ys = ops.cross_stitch(xd, ad, "reference_diag")
    ref = K.xstitch_reference_diag(alpha, torch.stack(xs))
    for t in range(2):
        assert torch.equal(ys[t].cpu(), ref[t])  # one multiply per element: bit-exact
    # linearity: stitch(2x) == 2 stitch(x) exactly (power-of-two scaling commutes with fp32 rounding)
    y2 = ops.cross_stitch([2 * x.detach() for x in xd], ad.detach(), "reference_diag")
    assert torch.equal(y2[0], 2 * ys[0].detach()) and torch.equal(y2[1], 2 * ys[1].detach())
    dys = [cl(torch.randn(B, Cc, H, W, generator=g)) for _ in range(2)]
    torch.autograd.backward(ys, dys)
    # dalpha[a,a,c] = sum dy_a * x_a over B,H,W (fp64 reference), off-diagonal exactly zero
    for a in range(2):
        want = (dys[a].double() * xd[a].detach().double()).sum((0, 2, 3))
        assert rel(ad.grad[a, a], want) <= TOL
        assert float(ad.grad[a, 1 - a].abs().max()) == 0.0
        assert torch.equal(xd[a].grad, dys[a] * ad.detach()[a, a].view(1, Cc, 1, 1))


# the four MTAN gate sites of BASELINE.json: N = 32 @ full resolution (M = 1 048 576), 64 @ /2
# (262 144), 128 @ /4 (65 536), 256 @ /8 (16 384).  The wide sites take the column-chunk paths of the
# forward and of both backward kernels; at M = 16 384 fewer tiles than SMs exist.
@pytest.mark.parametrize("N,down", [(32, 1), (64, 2), (128, 4), (256, 8)])
def test_gate_full_size_oracle_and_properties(N, down):
    from vision_mtl_b200 import ops

    H, W = globals()["H"] // down, globals()["W"] // down
    g = torch.Generator().manual_seed(11)
    h = torch.relu(torch.randn(B, 128, H, W, generator=g))
    s = torch.relu(torch.randn(B, N, H, W, generator=g))
    conv, bn = torch.nn.Conv2d(128, N, 1), torch.nn.BatchNorm2d(N)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g)
        bn.bias.uniform_(-0.5, 0.5, generator=g)
    dy = torch.randn(B, N, H, W, generator=g)

    def run(dy_dev):
        hd, sd = cl(h).requires_grad_(True), cl(s).requires_grad_(True)
        p = [t.detach().clone().to(dev()).requires_grad_(True) for t in (conv.weight, conv.bias, bn.weight, bn.bias)]
        rm, rv = bn.running_mean.clone().to(dev()), bn.running_var.clone().to(dev())
        y = ops.attention_gate(hd, sd, p[0], p[1], p[2], p[3], rm, rv, True, 0.1, 1e-5, "tc_3xtf32")
        y.backward(dy_dev)
        return y.detach(), hd.grad, sd.grad, p[0].grad, p[2].grad, p[3].grad, rm, rv

    out = run(cl(dy))
    again = run(cl(dy))
    for a, b in zip(out, again):  # fixed-order reductions, no float atomics: bitwise repeatable
        assert torch.equal(a, b)
    twice = run(cl(2 * dy))
    for a, b in zip(out[1:6], twice[1:6]):  # the backward is linear in dy; x2 is exact in fp32
        assert rel(b, 2 * a) <= 1e-6

    hr, sr = h.clone().requires_grad_(True), s.clone().requires_grad_(True)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    yr = K.gate_forward(hr, sr, conv.weight, conv.bias, bn.weight, bn.bias, rm, rv, True)
    yr.backward(dy)
    y, dh, ds, dW, dgamma, dbeta, rmd, rvd = out
    assert rel(y, yr) <= TOL
    assert rel(ds, sr.grad) <= TOL
    assert rel(dh, hr.grad) <= TOL
    assert rel(dW, conv.weight.grad) <= TOL
    assert rel(dgamma, bn.weight.grad) <= TOL
    assert rel(dbeta, bn.bias.grad) <= TOL
    assert rel(rmd, rm) <= TOL and rel(rvd, rv) <= TOL


def test_head_ce_full_size_oracle_and_conservation():
    from vision_mtl_b200 import ops

    g = torch.Generator().manual_seed(13)
    feat = torch.randn(B, 32, H, W, generator=g)
    head = torch.nn.Conv2d(32, C, 1)
    target = torch.randint(0, C, (B, H, W), generator=g)
    target[torch.rand(B, H, W, generator=g) < 0.25] = -100
    fd = cl(feat).requires_grad_(True)
    wd = head.weight.detach().to(dev()).requires_grad_(True)
    bd = head.bias.detach().to(dev()).requires_grad_(True)
    conf = torch.zeros(C, C, dtype=torch.int64, device=dev())
    loss, pred = ops.head_cross_entropy(fd, wd, bd, target.to(dev()), -100, conf, True)
    loss.backward()
    valid = target != -100
    assert int(conf.sum()) == int(valid.sum())
    assert torch.equal(conf.sum(1).cpu(), torch.bincount(target[valid], minlength=C))
    assert torch.equal(conf.sum(0).cpu(), torch.bincount(pred.cpu().long()[valid], minlength=C))
    # ignored pixels carry no gradient; every other row of dl sums to zero, so sum_c db_c == 0
    assert float(fd.grad.permute(0, 2, 3, 1)[~valid].abs().max()) == 0.0
    assert abs(float(bd.grad.sum())) <= 1e-6

    fr = feat.clone().requires_grad_(True)
    logits = K.head_project(fr, head.weight, head.bias)
    loss_r = K.cross_entropy(logits, target)
    loss_r.backward()
    assert rel(loss, loss_r) <= TOL
    assert rel(fd.grad, fr.grad) <= TOL
    assert rel(wd.grad, head.weight.grad) <= TOL
    assert rel(bd.grad, head.bias.grad) <= TOL
    pred_r = K.segm_predictions(logits.detach())
    mism = pred.cpu().long() != pred_r
    if mism.any():  # only fp32 near-ties of the top two logits may differ
        top2 = logits.detach().permute(0, 2, 3, 1)[mism].topk(2, dim=-1).values
        assert ((top2[:, 0] - top2[:, 1]).abs() < 1e-5).all()
