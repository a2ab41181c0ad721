"""Role-level cycle accounting of head_ce_tc_fwd_kernel (library built with VMTL_NVCC_EXTRA=-DVMTL_HT_PROF)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vision_mtl_b200 import ops, _lib
dev = torch.device("cuda:0")
B, H, W, C = 32, 128, 256, 19
feat = torch.randn(B, 32, H, W, device=dev).contiguous(memory_format=torch.channels_last)
head = torch.nn.Conv2d(32, C, 1).to(dev)
tgt = torch.randint(0, C, (B, H, W), device=dev)
conf = torch.zeros(C, C, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    ops.head_cross_entropy(feat, head.weight, head.bias, tgt, -100, conf, True)
torch.cuda.synchronize()
lib = _lib.load() if hasattr(_lib, "load") else _lib.lib
out = np.zeros((148, 16), dtype=np.int64)
f = lib.vmtl_debug_head_prof
f.argtypes = [ctypes.c_void_p]; f.restype = ctypes.c_int
assert f(out.ctypes.data) == 0
names = ["conv_total", "conv_w_full", "conv_w_amma", "conv_w_st", "epi_total", "epi_w_dfull", "prod_total", "prod_w_empty",
         "mma_total", "mma_w_aready", "mma_w_dfree", "mma_issue", "mma_commit"]
for i, n in enumerate(names):
    print(f"{n:14s} mean {out[:, i].mean():10.0f}  min {out[:, i].min():8d} max {out[:, i].max():8d}")

t_entry, t_loop0, t_loop1, t_exit = out[:, 13], out[:, 14], out[:, 15], out[:, 3]
t0 = t_entry.min()
print("entry  spread ns", t_entry.max() - t0)
print("prologue mean ns", (t_loop0 - t_entry).mean(), "max", (t_loop0 - t_entry).max())
print("loop     mean ns", (t_loop1 - t_loop0).mean(), "max", (t_loop1 - t_loop0).max())
print("tail     mean ns", (t_exit - t_loop1).mean(), "max", (t_exit - t_loop1).max())
print("span first entry -> last exit ns", t_exit.max() - t0)
