import sys, torch
sys.path.insert(0, '/root/repo')
from oracle import fixtures as FX, torch_port as TP
from vision_mtl_b200.models.mtan_model import MTANMiniUnet
C,B,H,W=19,2,32,64
net = MTANMiniUnet(3, {"depth":1,"segm":C},128,32,3)
sd = FX.fill_state_dict(net.state_dict(), salt=3)
batch = FX.image_batch(B,H,W,C,"full-grad")
orig=TP._bn
stats=[]
def bn(p,pre,x,training,momentum=0.1,eps=1e-5):
    v=x.var(dim=(0,2,3),unbiased=False); m=x.mean(dim=(0,2,3))
    stats.append((float(v.min()), float((v/(m*m+1e-30)).min()), pre, tuple(x.shape)))
    return orig(p,pre,x,training,momentum,eps)
TP._bn=bn
p={k:v.clone() for k,v in sd.items()}
raw=TP.mtan_forward(p,batch["img"],True)
for s in sorted(stats)[:8]: print('minvar %.2e  min var/mean2 %.2e  %s %s'%s)
print('depth logits range', raw['depth'].min().item(), raw['depth'].max().item(), 'segm', raw['segm'].abs().max().item())
