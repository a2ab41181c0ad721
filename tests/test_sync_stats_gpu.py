"""Global-batch statistics on ONE GPU: two replicas are emulated by running each op on the two
halves of a batch with a fake all-reduce, and must reproduce the plain op on the whole batch -- outputs, input
gradients, and parameter gradients (whose replica sum is the whole-batch gradient).

The fake collective works in passes: pass k knows the replica sums of collectives 0..k-1 (recorded by earlier
passes) and substitutes them; a collective whose inputs depend only on earlier collectives is therefore exact
from pass k on.  An op with n collectives needs n + 1 passes.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-5


class FakeWorld:
    """Replaces ops._allreduce_moments / ops.stat_sync_world: 2 replicas, run one after the other."""

    def __init__(self, ops):
        self.ops = ops
        self.sums = []  # sums[i] = replica sum of collective i, once known

    def run(self, fn_per_replica, n_collectives):
        ops = self.ops
        old = (ops._allreduce_moments, ops.stat_sync_world)
        ops.stat_sync_world = lambda: 2
        results = None
        try:
            for _ in range(n_collectives + 1):
                recorded = [[], []]
                results = []
                for r in range(2):
                    idx = [0]

                    def fake(t, r=r, idx=idx):
                        recorded[r].append(t.clone())
                        if idx[0] < len(self.sums):
                            t.copy_(self.sums[idx[0]])
                        idx[0] += 1

                    ops._allreduce_moments = fake
                    results.append(fn_per_replica(r))
                n_known = len(self.sums)
                if len(recorded[0]) > n_known:  # collective n_known had exact inputs in this pass
                    self.sums.append(recorded[0][n_known] + recorded[1][n_known])
        finally:
            ops._allreduce_moments, ops.stat_sync_world = old
        return results


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _cl(x):
    return x.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("C,H,W", [(32, 16, 24), (128, 8, 8), (64, 9, 13)])
@pytest.mark.parametrize("relu,pool", [(True, False), (False, False), (True, True)])
def test_bn_global_statistics_equal_whole_batch(C, H, W, relu, pool):
    from vision_mtl_b200 import ops

    if pool and (H % 2 or W % 2):
        pytest.skip("pooled case uses even sizes")
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    x = _cl((torch.randn(4, C, H, W, generator=g) * 1.5 + 0.3).to(dev))
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev)
    beta = torch.randn(C, generator=g).to(dev)
    dy_shape = (4, C, H // 2, W // 2) if pool else (4, C, H, W)
    # a non-zero mean (and a correlation with x) keeps the batch terms c1 = sum(g)/M, c2 = sum(g xhat)/M away from
    # zero: with centred random gradients a wrong normalisation of those terms would go unnoticed
    dy = _cl((torch.randn(dy_shape, generator=g) + 0.7).to(dev))
    if not pool:
        dy = _cl(dy + 0.2 * x)

    def run(xs, dys, rm, rv):
        xs = xs.clone().requires_grad_(True)
        ga, be = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        y = ops.BNReLUFunction.apply(xs, ga, be, rm, rv, True, 0.1, 1e-5, relu, pool)
        y.backward(dys)
        return y.detach(), xs.grad, ga.grad, be.grad

    rm_ref, rv_ref = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    y_ref, dx_ref, dg_ref, db_ref = run(x, dy, rm_ref, rv_ref)

    rms = [torch.zeros(C, device=dev) for _ in range(2)]
    rvs = [torch.ones(C, device=dev) for _ in range(2)]

    def replica(r):
        rms[r].zero_()
        rvs[r].fill_(1.0)
        return run(_cl(x[2 * r:2 * r + 2]), _cl(dy[2 * r:2 * r + 2]), rms[r], rvs[r])

    out = FakeWorld(ops).run(replica, 2)
    y = torch.cat([out[0][0], out[1][0]])
    dx = torch.cat([out[0][1], out[1][1]])
    assert _rel(y, y_ref) <= TOL
    assert _rel(dx, dx_ref) <= TOL
    assert _rel(out[0][2] + out[1][2], dg_ref) <= TOL
    assert _rel(out[0][3] + out[1][3], db_ref) <= TOL
    for r in range(2):  # running statistics of the GLOBAL batch on every replica
        assert _rel(rms[r], rm_ref) <= TOL and _rel(rvs[r], rv_ref) <= TOL


@pytest.mark.parametrize("N,H,W", [(32, 16, 24), (64, 9, 13), (128, 8, 8), (256, 4, 8), (48, 8, 8)])
@pytest.mark.parametrize("folded", [False, True])
def test_gate_global_statistics_equal_whole_batch(N, H, W, folded):
    from vision_mtl_b200 import ops

    dev = torch.device("cuda:0")
    K = 128
    tc = bool(ops._lib.load().vmtl_gate_tc_supported(K, N))
    if folded and not tc:
        pytest.skip("the folded hidden layer runs on tensor-core shapes only")
    prec = ops.GATE_TC_3XTF32 if tc else ops.GATE_FP32_FFMA
    g = torch.Generator().manual_seed(9)
    h = _cl(torch.randn(4, K, H, W, generator=g).to(dev))
    s = _cl(torch.randn(4, N, H, W, generator=g).to(dev))
    # non-zero-mean s and dy: the batch terms c1 / c2 of the gate's BatchNorm backward stay away from zero
    s = _cl(s + 0.8)
    dy = _cl((torch.randn(4, N, H, W, generator=g) + 0.6).to(dev))
    Wt = (torch.randn(N, K, 1, 1, generator=g) / K ** 0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    g2, b2 = (torch.rand(N, generator=g) + 0.5).to(dev), torch.randn(N, generator=g).to(dev)
    g1, b1 = (torch.rand(K, generator=g) + 0.5).to(dev), torch.randn(K, generator=g).to(dev)

    def run(hs, ss, dys):
        leaves = [t.clone().requires_grad_(True) for t in (hs, ss, Wt, bias, g2, b2, g1, b1)]
        hh, sv, w, bb, ga, be, ga1, be1 = leaves
        buf = lambda n, v: torch.full((n,), v, device=dev)  # noqa: E731
        if folded:
            y = ops.FoldedGateFunction.apply(hh, ga1, be1, buf(K, 0.0), buf(K, 1.0), True, 0.1, 1e-5, sv, w, bb, ga, be,
                                             buf(N, 0.0), buf(N, 1.0), True, 0.1, 1e-5, prec)
        else:
            y = ops.GateFunction.apply(hh, sv, w, bb, ga, be, buf(N, 0.0), buf(N, 1.0), True, 0.1, 1e-5, prec)
        y.backward(dys)
        grads = [t.grad for t in leaves[:6]] + ([leaves[6].grad, leaves[7].grad] if folded else [])
        return [y.detach()] + grads

    ref = run(h, s, dy)
    out = FakeWorld(ops).run(lambda r: run(_cl(h[2 * r:2 * r + 2]), _cl(s[2 * r:2 * r + 2]), _cl(dy[2 * r:2 * r + 2])),
                             4 if folded else 2)
    names = ["y", "dh", "ds", "dW", "dbias", "dgamma", "dbeta", "dgamma1", "dbeta1"]
    for i, name in enumerate(names[:len(ref)]):
        got = torch.cat([out[0][i], out[1][i]]) if i < 3 else out[0][i] + out[1][i]
        if name == "dbias":  # analytically zero under batch statistics: compare against the scale of dbeta
            assert float((got - ref[i]).abs().max()) <= 1e-4 * float(ref[6].abs().max()), name
            continue
        assert _rel(got, ref[i]) <= (1e-4 if name in ("dh", "dW", "dgamma1", "dbeta1") else TOL), (name, _rel(got, ref[i]))


def test_silog_global_moments_equal_whole_batch():
    from vision_mtl_b200 import ops

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    feat = _cl(torch.randn(4, 32, 16, 24, generator=g).to(dev))
    w = (torch.randn(1, 32, 1, 1, generator=g) * 0.2).to(dev)
    b = torch.zeros(1, device=dev)
    t = torch.rand(4, 16, 24, 1, generator=g)
    t[t < 0.2] = 0.0
    t = t.to(dev)

    def run(f, tt, scale):
        f = f.clone().requires_grad_(True)
        ww, bb = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        silog, mae, absrel, _ = ops.head_silog(f, ww, bb, tt, 1e-3, want_pred=False)
        (silog * scale).backward()
        return silog.detach(), mae, absrel, f.grad, ww.grad, bb.grad

    ref = run(feat, t, 1.0)
    # a data-parallel wrapper averages gradients over the 2 replicas: emulate with the 1/2 factor
    out = FakeWorld(ops).run(lambda r: run(_cl(feat[2 * r:2 * r + 2]), t[2 * r:2 * r + 2].contiguous(), 0.5), 1)
    for r in range(2):
        for i in range(3):
            assert _rel(out[r][i], ref[i]) <= TOL
    assert _rel(torch.cat([out[0][3], out[1][3]]), ref[3]) <= TOL
    assert _rel(out[0][4] + out[1][4], ref[4]) <= TOL
    assert _rel(out[0][5] + out[1][5], ref[5]) <= TOL
