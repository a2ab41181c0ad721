"""MMA-issuer cycle accounting of head_ce_tc_bwd_kernel (library built with VMTL_NVCC_EXTRA=-DVMTL_HT_PROF)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vision_mtl_b200 import ops, _lib
dev = torch.device("cuda:0")
B, H, W, C = 32, 128, 256, 19
feat = torch.randn(B, 32, H, W, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
head = torch.nn.Conv2d(32, C, 1).to(dev)
tgt = torch.randint(0, C, (B, H, W), device=dev)
conf = torch.zeros(C, C, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    loss = ops.head_cross_entropy(feat, head.weight, head.bias, tgt, -100, conf, True)[0]
    loss.backward()
torch.cuda.synchronize()
lib = _lib.load()
out = np.zeros((148, 8), dtype=np.int64)
f = lib.vmtl_debug_head_bwd_prof
f.argtypes = [ctypes.c_void_p]; f.restype = ctypes.c_int
assert f(out.ctypes.data) == 0
for i, n in enumerate(["wait_conv", "mma1+commit", "wait_dl", "mma2+commit", "mma3+commit", "total"]):
    print(f"{n:14s} mean {out[:, i].mean():10.0f}  min {out[:, i].min():8d} max {out[:, i].max():8d}")
