import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import torch.nn.functional as F
from oracle import fixtures as FX, torch_port as TP
from vision_mtl_b200 import ops
from vision_mtl_b200.lit_module import MTLModule
from vision_mtl_b200.models.mtan_model import MTANMiniUnet
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
dev='cuda:0'
C,B,H,W=19,2,32,64
def build():
    net = MTANMiniUnet(3, {"depth":1,"segm":C},128,32,3)
    sd = FX.fill_state_dict(net.state_dict(), salt=3); net.load_state_dict(sd); return net, sd
net, sd = build()
p = {k:(v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k,v in sd.items()}
batch = FX.image_batch(B,H,W,C,"full-grad")
raw = TP.mtan_forward(p, batch["img"], True)
res = TP.step_losses_and_metrics(raw, batch["mask"], batch["depth"], C); res["loss"].backward()
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu(); return ((a-b).abs().max()/b.abs().max().clamp_min(1e-30)).item()
calls=[]
orig = ops.attention_gate
def torch_gate(h,s,w,b,g,be,rm,rv,training,momentum=0.1,eps=1e-5,precision=None):
    z=F.conv2d(h,w,b); u=F.batch_norm(z,rm,rv,g,be,training,momentum,eps); return s*torch.sigmoid(u)
def run(gate, tag):
    import vision_mtl_b200.models.mtan_model as mm
    mm.ops.attention_gate = gate
    net,_ = build(); net.to(dev).to(memory_format=torch.channels_last).train()
    module = MTLModule(net, num_classes=C, device=dev)
    b = {k:v.to(dev) for k,v in batch.items()}
    loss = module.training_step(b,0); loss.backward()
    worst = sorted(((rel(q.grad,p[k].grad),k) for k,q in net.named_parameters() if p[k].grad.abs().max()>1e-5), reverse=True)[:4]
    print(tag, 'loss rel', rel(loss,res['loss']), 'worst', worst)
run(torch_gate, 'torch-gate')
run(orig, 'vmtl-gate ')
# capture inputs of each gate call and compare bwd against torch autograd on identical inputs
def checking_gate(h,s,w,b,g,be,rm,rv,training,momentum=0.1,eps=1e-5,precision=None):
    y = orig(h,s,w,b,g,be,rm.clone(),rv.clone(),training,momentum,eps,precision)
    hh=h.detach().clone().requires_grad_(True); ss=s.detach().clone().requires_grad_(True)
    ww=w.detach().clone().requires_grad_(True)
    yr = torch_gate(hh,ss,ww,b.detach(),g.detach(),be.detach(),rm.clone(),rv.clone(),training)
    h2=h.detach().clone().requires_grad_(True); s2=s.detach().clone().requires_grad_(True)
    y2 = orig(h2,s2,w.detach().clone().requires_grad_(True),b.detach(),g.detach(),be.detach(),rm.clone(),rv.clone(),training,momentum,eps,precision)
    dy = torch.randn_like(y)
    yr.backward(dy); y2.backward(dy)
    print('  gate call N=%d M=%d h.cl=%s s.cl=%s  y %.1e dh %.1e ds %.1e | h strides %s' % (s.shape[1], h.shape[0]*h.shape[2]*h.shape[3], h.is_contiguous(memory_format=torch.channels_last), s.is_contiguous(memory_format=torch.channels_last), rel(y2,yr), rel(h2.grad,hh.grad), rel(s2.grad,ss.grad), tuple(h.stride())))
    return orig(h,s,w,b,g,be,rm,rv,training,momentum,eps,precision)
run(checking_gate, 'checked   ')
