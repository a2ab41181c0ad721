#!/usr/bin/env python
"""Per-site micro-benchmark of the hand-written kernels at the BASELINE.json shapes (SURVEY A.1-A.3:
Cityscapes-shaped batch 32, 128x256).  Each op is captured into a CUDA graph and timed alone with
CUDA events around the replay (device time of all launches of the call, no host launch gaps), an L2
flush (write of a 256 MiB buffer) before every replay and between forward and backward, and reported
as algorithmic GB/s against MEASURED_PEAKS.json.  Also the target command for the ncu captures kept under profiles/.

    python tools/site_bench.py [--iters 5] [--only gate|bn|up|xstitch|heads|metrics] [--once] [--json out.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

from vision_mtl_b200 import ops  # noqa: E402

B, H, W, C = 32, 128, 256, 19
GATE_SITES = [("enc0/dec3", 32, 1), ("enc1/dec2", 64, 2), ("enc2/dec1", 128, 4), ("enc3/dec0", 256, 8)]
XS_SITES = [(16, 64, 128), (24, 32, 64), (40, 16, 32), (80, 8, 16), (112, 8, 16), (160, 4, 8), (1072, 8, 16),
            (296, 16, 32), (152, 32, 64), (80, 64, 128), (32, 128, 256)]


def peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--once", action="store_true", help="one un-timed pass (ncu target)")
    ap.add_argument("--precision", default="tc_3xtf32")
    ap.add_argument("--json", default="")
    ap.add_argument("--profile-set", action="store_true",
                    help="only the largest gate site, the largest cross-stitch site and the heads (ncu --set full target)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pk = peak()
    rows = []

    def cl(*shape):
        return torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last)

    def replay_ms(graph, iters):
        best = 1e9
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    # the L2 flush that separates forward and backward inside the fwd+bwd graph, timed alone
    g_flush = torch.cuda.CUDAGraph()
    flush.zero_()
    torch.cuda.synchronize()
    with torch.cuda.graph(g_flush, stream=torch.cuda.current_stream()):
        flush.zero_()
    t_flush = replay_ms(g_flush, 5)

    def timed(name, site, nbytes_fwd, nbytes_bwd, fwd, bwd=None):
        """Each call is captured into a CUDA graph and replayed: the number is device time of ALL launches
        of the call with no host launch gaps (eager timing of 20-50 us calls mostly measures the host)."""
        for _ in range(1 if args.once else 2):  # eager warm-up (also the ncu target with --once)
            out = fwd()
            if bwd is not None:
                bwd(out)
        torch.cuda.synchronize()
        if args.once:
            return
        g_f = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_f, stream=torch.cuda.current_stream()):
            out = fwd()
        best = [replay_ms(g_f, args.iters), 0.0]
        if bwd is not None:
            g_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fb, stream=torch.cuda.current_stream()):
                out = fwd()
                flush.zero_()
                bwd(out)
            best[1] = replay_ms(g_fb, args.iters) - best[0] - t_flush
        for tag, ms, nb in (("fwd", best[0], nbytes_fwd), ("bwd", best[1], nbytes_bwd)):
            if nb:
                gbps = nb / (ms * 1e-3) / 1e9
                rows.append({"op": f"{name}_{tag}", "site": site, "ms": ms, "MB": nb / 1e6, "gbps": gbps, "frac": gbps / pk})
                print(f"{name + '_' + tag:18s} {site:22s} {ms * 1e3:9.1f} us {nb / 1e6:9.1f} MB {gbps:8.1f} GB/s  {gbps / pk:5.3f}")

    if args.only in ("", "gate"):
        for site, N, down in (GATE_SITES[:1] if args.profile_set else GATE_SITES):
            h, w = H // down, W // down
            M = B * h * w
            hh = torch.relu(cl(B, 128, h, w)).requires_grad_(True)
            ss = torch.relu(cl(B, N, h, w)).requires_grad_(True)
            conv, bn = torch.nn.Conv2d(128, N, 1).to(dev), torch.nn.BatchNorm2d(N).to(dev)
            dy = cl(B, N, h, w)

            def fwd():
                return ops.attention_gate(hh, ss, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean,
                                          bn.running_var, True, 0.1, 1e-5, args.precision)

            def bwd(y):
                hh.grad = ss.grad = None
                y.backward(dy)

            # SURVEY 8(d) algorithmic bytes (what the kernels issue is ops.gate_bytes(...)[1])
            nb_f, nb_b = ops.gate_bytes(M, 128, N, True, False)[0], ops.gate_bytes(M, 128, N, True, True)[0]
            timed("gate", f"{site} N={N} M={M}", nb_f, nb_b, fwd, bwd)
            if args.profile_set:
                continue
            # the same gate with the hidden layer folded in: c = conv1 output, h = relu(bn1(c)) never materialised
            # (statistics pass over c + gate forward; gate backward + BatchNorm/ReLU backward over (dh, c))
            cc = cl(B, 128, h, w).requires_grad_(True)
            bn1 = torch.nn.BatchNorm2d(128).to(dev)

            def fwd_f():
                return ops.attention_gate_folded(cc, bn1, ss, conv, bn, args.precision)

            def bwd_f(y):
                cc.grad = ss.grad = None
                y.backward(dy)

            timed("gate_folded", f"{site} N={N} M={M}", nb_f + 4 * M * 128, nb_b + 4 * M * 128 * 5, fwd_f, bwd_f)

    if args.only in ("", "bn"):
        # DoubleConv / attention conv3 BatchNorm + ReLU (+ 2x2 max-pool) sites of the MTAN network, batch statistics
        bn_sites = ((32, 1, False), (64, 2, False), (128, 4, False), (256, 8, False), (32, 1, True), (64, 2, True))
        for Cc, down, pool in (bn_sites[:1] if args.profile_set else bn_sites):
            h, w = H // down, W // down
            M = B * h * w
            x = cl(B, Cc, h, w).requires_grad_(True)
            bnm = torch.nn.BatchNorm2d(Cc).to(dev)
            dyb = cl(B, Cc, h // 2, w // 2) if pool else cl(B, Cc, h, w)
            out_rows = M // 4 if pool else M

            def fwd_b():
                return ops.batch_norm_relu(x, bnm, True, pool)

            def bwd_b(y):
                x.grad = None
                y.backward(dyb)

            timed("bnrelu_pool" if pool else "bnrelu", f"C={Cc} M={M}", 4 * Cc * (2 * M + out_rows),
                  4 * Cc * (2 * (M + out_rows) + M), fwd_b, bwd_b)

    if args.only in ("", "up"):
        # decoder attention modules: cat((conv1_shared, bilinear_x2(prev)), 1) -- dec3 / dec2 / dec1 of the MTAN network
        up_sites = ((64, 128, 1), (128, 128, 2), (256, 128, 4))
        for C1, Cx, down in (up_sites[:1] if args.profile_set else up_sites):
            h, w = H // down, W // down
            first = cl(B, C1, h, w).requires_grad_(True)
            xin = cl(B, Cx, h // 2, w // 2).requires_grad_(True)
            dyu = cl(B, C1 + Cx, h, w)

            def fwd_u():
                return ops.upsample2_cat(first, xin)

            def bwd_u(y):
                first.grad = xin.grad = None
                y.backward(dyu)

            # algorithmic: the up-sampling itself (R x, W 4x) + the copy of `first` into the concatenation (R + W)
            nb = 20 * xin.numel() + 8 * first.numel()
            timed("up2_cat", f"C1={C1} Cx={Cx} {h}x{w}", nb, 20 * xin.numel(), fwd_u, bwd_u)

    if args.only in ("", "xstitch"):
        for Cc, h, w in (XS_SITES[-1:] if args.profile_set else XS_SITES):
            xs = [cl(B, Cc, h, w).requires_grad_(True) for _ in range(2)]
            alpha = torch.rand(2, 2, Cc, device=dev, requires_grad=True)
            dys = [cl(B, Cc, h, w) for _ in range(2)]
            E = 2 * B * Cc * h * w

            def fwd():
                return ops.cross_stitch(xs, alpha, "reference_diag")

            def bwd(ys):
                for t_ in xs:
                    t_.grad = None  # otherwise AccumulateGrad adds into the old .grad (an extra 3x pass)
                alpha.grad = None
                torch.autograd.backward(ys, dys)

            timed("xstitch", f"C={Cc} {h}x{w}", 8 * E, 12 * E, fwd, bwd)

    if args.only in ("", "heads"):
        P = B * H * W
        feat = cl(B, 32, H, W).requires_grad_(True)
        head = torch.nn.Conv2d(32, C, 1).to(dev)
        dhead = torch.nn.Conv2d(32, 1, 1).to(dev)
        tgt = torch.randint(0, C, (B, H, W), device=dev)
        dt = torch.rand(B, H, W, 1, device=dev) * 0.5
        conf = torch.zeros(C, C, dtype=torch.int64, device=dev)
        def bw(leaves):
            def f(l):
                for t_ in leaves:
                    t_.grad = None
                l.backward()
            return f

        timed("head_ce", f"P={P} C={C}", P * (4 * 32 + 8 + 1), P * (8 * 32 + 8),
              lambda: ops.head_cross_entropy(feat, head.weight, head.bias, tgt, -100, conf, True)[0],
              bw([feat, head.weight, head.bias]))
        timed("head_silog", f"P={P}", P * (4 * 32 + 4), P * (8 * 32 + 4),
              lambda: ops.head_silog(feat, dhead.weight, dhead.bias, dt, 1e-3, False)[0],
              bw([feat, dhead.weight, dhead.bias]))
        for layout in ("nhwc", "nchw"):
            lg = torch.randn(B, C, H, W, device=dev)
            if layout == "nhwc":
                lg = lg.contiguous(memory_format=torch.channels_last)
            lg.requires_grad_(True)
            timed(f"ce_logits_{layout}", f"P={P} C={C}", P * (4 * C + 8 + 1), P * (8 * C + 8),
                  lambda: ops.cross_entropy_logits(lg, tgt, -100, conf, True)[0], bw([lg]))
        dl = torch.randn(B, 1, H, W, device=dev, requires_grad=True)
        timed("silog_logits", f"P={P}", P * 8, P * 12, lambda: ops.head_silog(dl, None, None, dt, 1e-3, False)[0],
              bw([dl]))

    if args.only in ("", "metrics"):
        P = B * H * W
        tgt = torch.randint(0, C, (B, H, W), device=dev)
        p64 = torch.randint(0, C, (B, H, W), device=dev)
        p8 = p64.to(torch.uint8)
        pr, dt = torch.rand(P, device=dev), torch.rand(P, device=dev)
        timed("confusion_i64", f"P={P}", P * 16, 0, lambda: ops.confusion_accumulate(p64, tgt, C))
        timed("confusion_u8", f"P={P}", P * 9, 0, lambda: ops.confusion_accumulate(p8, tgt, C))
        timed("depth_err_sums", f"P={P}", P * 8, 0, lambda: ops.depth_error_sums(pr, dt))

    if args.json:
        json.dump({"peak_gbs": pk, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    # everything (leaf creation, warm-up, capture, replay) on ONE non-default stream: autograd binds a
    # leaf's gradient accumulator to the stream it first ran on, and capture cannot touch the legacy stream
    _side = torch.cuda.Stream()
    with torch.cuda.stream(_side):
        main()
