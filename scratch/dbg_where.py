import sys, torch
sys.path.insert(0, '/root/repo')
import torch.nn.functional as F
from oracle import fixtures as FX, torch_port as TP
from vision_mtl_b200.models.mtan_model import MTANMiniUnet
C,B,H,W=19,2,32,64
net = MTANMiniUnet(3, {"depth":1,"segm":C},128,32,3)
sd = FX.fill_state_dict(net.state_dict(), salt=3)
batch = FX.image_batch(B,H,W,C,"full-grad")
store={}
orig_conv, orig_bn = TP._conv, TP._bn
def run(dtype):
    acts={}
    def conv(p,pre,x,padding=0):
        y=orig_conv(p,pre,x,padding); y.retain_grad(); acts['conv:'+pre]=y; return y
    def bn(p,pre,x,training,momentum=0.1,eps=1e-5):
        y=orig_bn(p,pre,x,training,momentum,eps); y.retain_grad(); acts['bn:'+pre]=y; return y
    TP._conv, TP._bn = conv, bn
    p = {k:(v.clone().to(dtype).requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.clone().to(dtype) if v.is_floating_point() else v.clone())) for k,v in sd.items()}
    raw = TP.mtan_forward(p, batch["img"].to(dtype), True)
    res = TP.step_losses_and_metrics(raw, batch["mask"], batch["depth"].to(dtype), C); res["loss"].backward()
    return acts
a32=run(torch.float32); a64=run(torch.float64)
def rel(a,b): a=a.double(); b=b.double(); return ((a-b).abs().max()/b.abs().max().clamp_min(1e-30)).item()
for k in a32:
    if 'task_attn_modules.0' in k and ('dec_layers.2' in k or 'dec_layers.1' in k):
        print('%-60s fwd %.1e  grad %.1e' % (k, rel(a32[k],a64[k]), rel(a32[k].grad,a64[k].grad)))
