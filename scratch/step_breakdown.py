"""Where the device time of one eager training step goes (torch.profiler / CUPTI), grouped by kernel family."""
import argparse, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from vision_mtl_b200.synthetic import make_batch
from vision_mtl_b200.utils.pipeline_utils import DataShape, init_model

ap = argparse.ArgumentParser(); ap.add_argument("--model", default="mtan"); a = ap.parse_args()
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.benchmark = True
torch.manual_seed(11)
margs = argparse.Namespace(model_name=a.model, backbone_weights=None, channel_wise_stitching=True,
                           stitch_mode="reference_diag", lr=1e-3, device=dev, ckpt_dir=None)
module = init_model(margs, DataShape(num_classes=19, height=128, width=256, name="cityscapes"))
module.to(dev); module.model.to(memory_format=torch.channels_last); module.model.train()
from vision_mtl_b200.optim import Adam
opt = Adam(module.parameters(), lr=1e-3)
batch = {k: v.to(dev) for k, v in make_batch(32, 128, 256, 19, "cityscapes", seed=11).items()}
batch["img"] = batch["img"].contiguous(memory_format=torch.channels_last)
def step():
    opt.zero_grad(set_to_none=True)
    loss = module.training_step(batch, 0); loss.backward(); opt.step()
for _ in range(4): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
fam = collections.OrderedDict([
    ("hand-written (vmtl::*)", lambda n: "vmtl::" in n),
    ("batch norm (cuDNN / ATen)", lambda n: any(t in n.lower() for t in ("bn_fw", "bn_bw", "batch_norm", "batchnorm"))),
    ("cuDNN / cuBLAS convolution + GEMM", lambda n: any(t in n for t in ("cudnn", "cutlass", "gemm", "conv", "xmma", "sm80_", "sm90_", "sm100_", "implicit", "wgrad", "dgrad", "nchwToNhwc", "nhwcToNchw"))),
    ("multi-tensor ATen (counters)", lambda n: "multi_tensor" in n),
    ("pooling / upsample", lambda n: any(t in n for t in ("pool", "upsample", "interp"))),
    ("other elementwise / copies / reductions", lambda n: True)])
tot = collections.Counter(); cnt = collections.Counter(); byname = collections.Counter(); full = collections.Counter(); fcnt = collections.Counter()
for e in prof.events():
    if e.device_type is None or "cuda" not in str(e.device_type).lower(): continue
    t = getattr(e, "device_time", None) or getattr(e, "cuda_time", 0)
    for name, f in fam.items():
        if f(e.name): tot[name] += t; cnt[name] += 1; byname[e.name[:90]] += t; full[(name, e.name[:260])] += t; fcnt[(name, e.name[:260])] += 1; break
s = sum(tot.values())
print(f"{a.model}: one eager training step, batch 32 x 128x256, fp32 convolutions; device time {s/1e3:.1f} ms in {sum(cnt.values())} kernels")
for name in fam: print(f"  {name:42s} {tot[name]/1e3:8.2f} ms  {100*tot[name]/max(s,1):5.1f} %  {cnt[name]:5d} kernels")
print("  top kernels:")
for n, t in byname.most_common(12): print(f"    {t/1e3:8.2f} ms  {n}")

print("  non-convolution ATen kernels (family, ms, launches, name):")
for (fam_name, n), t in full.most_common(400):
    if fam_name in ("pooling / upsample", "other elementwise / copies / reductions", "multi-tensor ATen (counters)"):
        print(f"    {t/1e3:7.2f} ms {fcnt[(fam_name, n)]:4d}x  {n}")
