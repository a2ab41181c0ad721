#!/usr/bin/env python
"""Per-site micro-benchmark of the hand-written kernels at the BASELINE.json shapes (SURVEY A.1-A.3:
Cityscapes-shaped batch 32, 128x256).  Each op is timed alone with CUDA events, an L2 flush
(write of a 256 MiB buffer) between iterations, and reported as algorithmic GB/s against
MEASURED_PEAKS.json.  Also the target command for the ncu captures kept under profiles/.

    python tools/site_bench.py [--iters 5] [--only gate|xstitch|heads|metrics] [--once] [--json out.json]
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

    def timed(name, site, nbytes_fwd, nbytes_bwd, fwd, bwd=None):
        iters = 1 if args.once else args.iters
        best = [1e9, 1e9]
        for it in range(iters + (0 if args.once else 2)):
            flush.zero_()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
            out = fwd()
            e[1].record()
            if bwd is not None:
                flush.zero_()
                e[2].record()
                bwd(out)
                e[3].record()
            torch.cuda.synchronize()
            if it >= (0 if args.once else 2):
                best[0] = min(best[0], e[0].elapsed_time(e[1]))
                if bwd is not None:
                    best[1] = min(best[1], e[2].elapsed_time(e[3]))
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

            timed("gate", f"{site} N={N} M={M}", 4 * M * (128 + 4 * N), 4 * M * (2 * 128 + 7 * N), fwd, bwd)

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
    main()
