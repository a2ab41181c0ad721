"""My single-GPU whole-batch gradients vs the CPU oracle's autograd, on the dist-test model."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import fixtures as FX, torch_port as TP
from vision_mtl_b200.lit_module import MTLModule
from vision_mtl_b200.models.mtan_model import MTANMiniUnet
C, B, H, W = 19, int(sys.argv[1]) if len(sys.argv) > 1 else 4, 32, 64
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(5)
net = MTANMiniUnet(3, {"depth": 1, "segm": C}, 128, 32, 3)
sd = {k: v.clone() for k, v in net.state_dict().items()}
full = FX.image_batch(4, H, W, C, "dist/batch")
full = {k: v[:B] for k, v in full.items()}
p = {k: v.clone().requires_grad_(v.dtype == torch.float32 and "running" not in k) for k, v in sd.items()}
raw = TP.mtan_forward(p, full["img"], training=True)
res = TP.step_losses_and_metrics(raw, full["mask"], full["depth"], C)
res["loss"].backward()
net.to(dev).to(memory_format=torch.channels_last).train()
module = MTLModule(net, num_classes=C, device=dev)
batch = {k: v.to(dev) for k, v in full.items()}
batch["img"] = batch["img"].contiguous(memory_format=torch.channels_last)
loss = module.training_step(batch, 0)
loss.backward()
print("loss", float(loss), float(res["loss"]))
errs = []
for k, q in net.named_parameters():
    r = p[k].grad
    nr = float(r.norm())
    errs.append((float((q.grad.cpu().double() - r.double()).norm()) / max(nr, 1e-30), k, nr))
typ = sorted(e[2] for e in errs)[len(errs) // 2]
live = sorted([e for e in errs if e[2] > 1e-5 * typ], reverse=True)
print("median", live[len(live) // 2][0])
for e in live[:10]:
    print(e)
