"""Observed relative errors of the tensor-core head (fwd + bwd) against the CPU oracle, several seeds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import kernels_ref as K
from vision_mtl_b200 import ops
dev = torch.device("cuda:0")
def rel(g, r):
    g, r = g.detach().double().cpu(), r.detach().double().cpu()
    return (g - r).abs().max().item() / max(r.abs().max().item(), 1e-30)
for (B, C, H, W) in [(1, 19, 16, 32), (2, 19, 128, 256), (1, 13, 64, 64), (8, 19, 128, 256), (1, 32, 32, 32)]:
    for seed in range(3):
        g = torch.Generator().manual_seed(100 + seed)
        feat = torch.randn(B, 32, H, W, generator=g) * (1 + seed)
        head = torch.nn.Conv2d(32, C, 1)
        tgt = torch.randint(0, C, (B, H, W), generator=g)
        tgt[torch.rand(B, H, W, generator=g) < 0.2] = -100
        fr = feat.clone().requires_grad_(True)
        lr_ = K.cross_entropy(K.head_project(fr, head.weight, head.bias), tgt); lr_.backward()
        fd = feat.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        wd = head.weight.detach().to(dev).requires_grad_(True); bd = head.bias.detach().to(dev).requires_grad_(True)
        l, _ = ops.head_cross_entropy(fd, wd, bd, tgt.to(dev), -100, None, True); l.backward()
        print(f"P={B*H*W:8d} C={C:2d} seed={seed} loss {rel(l, lr_):.1e} dfeat {rel(fd.grad, fr.grad):.1e} "
              f"dW {rel(wd.grad, head.weight.grad):.1e} db {rel(bd.grad, head.bias.grad):.1e}")
