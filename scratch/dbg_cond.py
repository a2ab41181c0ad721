import sys, torch
sys.path.insert(0, '/root/repo')
from oracle import fixtures as FX, torch_port as TP
from vision_mtl_b200.models.mtan_model import MTANMiniUnet
C,B,H,W=19,2,32,64
net = MTANMiniUnet(3, {"depth":1,"segm":C},128,32,3)
sd = FX.fill_state_dict(net.state_dict(), salt=3)
batch = FX.image_batch(B,H,W,C,"full-grad")
def run(dtype):
    p = {k:(v.clone().to(dtype).requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.clone().to(dtype) if v.is_floating_point() else v.clone())) for k,v in sd.items()}
    raw = TP.mtan_forward(p, batch["img"].to(dtype), True)
    res = TP.step_losses_and_metrics(raw, batch["mask"], batch["depth"].to(dtype), C); res["loss"].backward()
    return p, res
p32,r32 = run(torch.float32); p64,r64 = run(torch.float64)
def rel(a,b): a=a.double(); b=b.double(); return ((a-b).abs().max()/b.abs().max().clamp_min(1e-30)).item()
print('loss', r32['loss'].item(), r64['loss'].item())
w = sorted(((rel(p32[k].grad,p64[k].grad),k) for k in p32 if p32[k].requires_grad and p64[k].grad.abs().max()>1e-5), reverse=True)[:8]
for e,k in w: print('%.3e %s'%(e,k))
print('---- abs view')
for k in ['dec_layers.1.task_attn_modules.0.conv_out.weight','dec_layers.1.task_attn_modules.1.conv_out.weight','dec_layers.2.task_attn_modules.0.conv_out.weight','map_tasks_to_heads.depth.weight','map_tasks_to_heads.segm.weight','enc_layers.0.dconv.double_conv.0.weight']:
    a=p32[k].grad.double(); b=p64[k].grad
    print('%-55s max|g64| %.3e  max|d| %.3e  norm64 %.3e  normd %.3e'%(k,b.abs().max(),(a-b).abs().max(),b.norm(),(a-b).norm()))
