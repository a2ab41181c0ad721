"""Accuracy of the channel-stacked 3xTF32 convolution (vision_mtl_b200/conv3x.py) against fp64, next to cuDNN's strict
fp32 and plain TF32 paths; then the golden MTAN step tests with the scheme enabled."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from vision_mtl_b200 import conv3x
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
for (B, Ci, Co, H, W, k) in [(8, 32, 32, 64, 128, 3), (8, 192, 128, 32, 64, 1), (4, 128, 128, 32, 32, 3)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Ci, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).to(dev)
    dy = torch.randn(B, Co, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    def run(kind):
        xx, ww = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        if kind == "fp64":
            y = F.conv2d(xx.double(), ww.double(), None, 1, k // 2); y.backward(dy.double())
        elif kind == "3x":
            y = conv3x.Conv3xTF32Function.apply(xx, ww, None, (k // 2, k // 2), (1, 1)); y.backward(dy)
        else:
            torch.backends.cudnn.allow_tf32 = kind == "tf32"
            y = F.conv2d(xx, ww, None, 1, k // 2); y.backward(dy)
            torch.backends.cudnn.allow_tf32 = False
        return y.detach(), xx.grad, ww.grad
    ref = run("fp64")
    for kind in ("fp32", "3x", "tf32"):
        out = run(kind)
        print(f"Ci={Ci} Co={Co} k={k}: {kind:5s} y {rel(out[0], ref[0]):.2e} dx {rel(out[1], ref[1]):.2e} dW {rel(out[2], ref[2]):.2e}")
if len(sys.argv) > 1 and sys.argv[1] == "golden":
    import pytest
    conv3x.enable(True)
    sys.exit(pytest.main(["-q", "-x", "-m", "gpu", os.path.join(os.path.dirname(__file__), "..", "tests", "test_models_gpu.py"), "-k", "golden or default_width"]))
