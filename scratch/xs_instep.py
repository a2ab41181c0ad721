"""Per-call cross-stitch timings inside a real csnet step (device kept ahead of the host)."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_mtl_b200 import ops
from vision_mtl_b200.synthetic import make_batch
from vision_mtl_b200.utils.pipeline_utils import DataShape, init_model
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
torch.manual_seed(11)
margs = argparse.Namespace(model_name="csnet", backbone_weights=None, channel_wise_stitching=True,
                           stitch_mode="reference_diag", lr=1e-3, device=dev, ckpt_dir=None)
module = init_model(margs, DataShape(num_classes=19, height=128, width=256, name="cityscapes"))
module.to(dev); module.model.to(memory_format=torch.channels_last); module.model.train()
opt = torch.optim.Adam(module.parameters(), lr=1e-3, fused=True)
batch = {k: v.to(dev) for k, v in make_batch(32, 128, 256, 19, "cityscapes", seed=11).items()}
batch["img"] = batch["img"].contiguous(memory_format=torch.channels_last)
def step():
    opt.zero_grad(set_to_none=True)
    loss = module.training_step(batch, 0); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with ops.kernel_timing() as rec:
    torch.cuda._sleep(100_000_000)
    step()
    torch.cuda.synchronize()
tot = {}
for name, nb, e0, e1 in rec:
    if "xstitch" in name:
        ms = e0.elapsed_time(e1)
        print(f"{name:14s} {nb/1e6:8.1f} MB {ms*1e3:8.1f} us {nb/ms/1e6:8.1f} GB/s")
        t = tot.setdefault(name, [0, 0.0]); t[0] += nb; t[1] += ms
for k, (nb, ms) in tot.items(): print(k, nb/1e6, "MB", ms*1e3, "us", nb/ms/1e6, "GB/s")
