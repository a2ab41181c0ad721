"""Model-level emulation of the global-statistics mode on ONE GPU (FakeWorld of tests/test_sync_stats_gpu.py):
2 replicas run one after the other, collectives replaced by recorded sums; the averaged gradients must equal the
whole-batch gradients.  Diagnostic for tests/test_dist_gpu.py (sync mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import fixtures as FX
from vision_mtl_b200 import ops
from vision_mtl_b200.lit_module import MTLModule
from vision_mtl_b200.models.mtan_model import MTANMiniUnet
from test_sync_stats_gpu import FakeWorld

C, B, H, W = 19, 4, 32, 64
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.deterministic = True
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 3


def build():
    torch.manual_seed(5)
    net = MTANMiniUnet(3, {"depth": 1, "segm": C}, 128, 32, levels)
    net.to(dev).to(memory_format=torch.channels_last).train()
    return MTLModule(net, num_classes=C, device=dev)


def to_dev(batch, sl=slice(None)):
    out = {k: v[sl].to(dev) for k, v in batch.items()}
    out["img"] = out["img"].contiguous(memory_format=torch.channels_last)
    return out


full = FX.image_batch(B, H, W, C, "dist/batch")
module = build()
state = {k: v.clone() for k, v in module.model.state_dict().items()}


def run(batch):
    module.model.load_state_dict(state)
    for p in module.parameters():
        p.grad = None
    loss = module.training_step(batch, 0)
    loss.backward()
    return loss.detach().double().item(), {k: p.grad.detach().clone() for k, p in module.model.named_parameters()}


ref_loss, ref = run(to_dev(full))
count = [0]
orig = ops._allreduce_moments
def counting(t):
    count[0] += 1
ops._allreduce_moments = counting
old_world = ops.stat_sync_world
ops.stat_sync_world = lambda: 2
run(to_dev(full, slice(0, 2)))
ops._allreduce_moments, ops.stat_sync_world = orig, old_world
n = count[0]
print("collectives per step:", n, flush=True)
out = FakeWorld(ops).run(lambda r: run(to_dev(full, slice(2 * r, 2 * r + 2))), n)
loss = (out[0][0] + out[1][0]) / 2
print("loss rel", abs(loss - ref_loss) / abs(ref_loss))
errs = []
for k in ref:
    g = (out[0][1][k] + out[1][1][k]) / 2
    nr = float(ref[k].double().norm())
    errs.append((float((g.double() - ref[k].double()).norm()) / max(nr, 1e-30), k, nr))
typ = sorted(e[2] for e in errs)[len(errs) // 2]
live = sorted([e for e in errs if e[2] > 1e-5 * typ], reverse=True)
print("median", live[len(live) // 2][0], "max", live[0])
for e in live[:8]:
    print(e)
for e in errs:
    if e[1].startswith("dec_layers.%d." % (levels - 1)) or e[1].startswith("map_"):
        print("LAST", e)
# which layers are fine?  smallest errors
for e in live[-8:]:
    print("ok", e)

# ---- where does the backward first diverge?  record every gate / BatchNorm backward call -----------------------------
rec = []
_gb, _bb = ops._gate_backward, ops._bn_backward
def gate_bwd(dy, h, h_coef, s, z, *a, **k):
    out = _gb(dy, h, h_coef, s, z, *a, **k)
    rec.append(("gate", dy.detach().clone(), None if out[0] is None else out[0].detach().clone(), out[1].detach().clone() if out[1] is not None else None))
    return out
def bn_bwd(dy, x, stats, training, relu, pool, need_dx, world, *extra):
    out = _bb(dy, x, stats, training, relu, pool, need_dx, world, *extra)
    rec.append(("bn", dy.detach().clone(), None if out[0] is None else out[0].detach().clone(), out[2].detach().clone()))
    return out
ops._gate_backward, ops._bn_backward = gate_bwd, bn_bwd
rec.clear(); run(to_dev(full)); whole = list(rec)
fw = FakeWorld(ops)
reps = [None, None]
def rep(r):
    rec.clear()
    o = run(to_dev(full, slice(2 * r, 2 * r + 2)))
    reps[r] = list(rec)
    return o
fw.run(rep, n)
def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
for i, (kind, dy, dx, extra) in enumerate(whole):
    dy2 = torch.cat([reps[0][i][1], reps[1][i][1]]) / 2
    line = f"{i:3d} {kind:4s} C={dy.shape[1]:4d} HW={dy.shape[2]}x{dy.shape[3]} dy {rel(dy2, dy):.2e}"
    if dx is not None:
        dx2 = torch.cat([reps[0][i][2], reps[1][i][2]]) / 2
        line += f" dx {rel(dx2, dx):.2e}"
    if extra is not None:
        e2 = (torch.cat([reps[0][i][3], reps[1][i][3]]) / 2) if kind == "gate" else (reps[0][i][3] + reps[1][i][3]) / 2
        line += f" {'ds' if kind == 'gate' else 'dbeta'} {rel(e2, extra):.2e}"
    print(line)
    if i > 14:
        break

# ---- are the saved tensors of a BatchNorm intact when its backward runs? ---------------------------------------------
ops._gate_backward, ops._bn_backward = _gb, _bb
saved = {}
_bf = ops._bn_forward
def bn_fwd(x, gamma, beta, rm, rv, training, momentum, eps, relu, pool, y, stats, world, *extra):
    _bf(x, gamma, beta, rm, rv, training, momentum, eps, relu, pool, y, stats, world, *extra)
    saved[(x.data_ptr(), stats.data_ptr())] = (x.detach().clone(), stats.detach().clone())
def bn_bwd2(dy, x, stats, training, relu, pool, need_dx, world, *extra):
    key = (x.data_ptr(), stats.data_ptr())
    if key in saved:
        x0, s0 = saved[key]
        dx_ = float((x0 - x).abs().max()); ds_ = float((s0 - stats).abs().max() / s0.abs().max())
        if dx_ > 0 or ds_ > 0:
            print(f"  CHANGED since forward: C={x.shape[1]} HW={x.shape[2]}x{x.shape[3]} max|dx|={dx_:.3e} stats rel={ds_:.3e} rows changed:",
                  [int(i) for i in range(stats.shape[0]) if float((s0[i] - stats[i]).abs().max()) > 0])
    else:
        print("  (no forward record)", x.shape)
    return _bb(dy, x, stats, training, relu, pool, need_dx, world, *extra)
ops._bn_forward, ops._bn_backward = bn_fwd, bn_bwd2
old_world = ops.stat_sync_world
ops.stat_sync_world = lambda: 2
ops._allreduce_moments = lambda t: None
print("replica run with checks:")
run(to_dev(full, slice(0, 2)))
ops.stat_sync_world = old_world
print("whole-batch run with checks:")
run(to_dev(full))
