"""Worker of tests/test_dist_gpu.py: launched by torch.distributed.run with 2 ranks (NCCL), one GPU each.

mode "ddp"  : default data-parallel semantics (per-replica BatchNorm / SILog).  The 2-rank step -- DDP around the
              model, whole step captured by GraphedTrainStep, packed metric all-reduce inside the graph -- must give
                * a global confusion matrix bit-equal to confusion(shard 0) + confusion(shard 1) of a single process,
                * gradients equal to the mean of the two per-shard gradients.
mode "sync" : global-batch-exact mode (ops.set_stat_sync).  The 2-rank step must equal the
              single-process step on the CONCATENATED batch: loss, every gradient, running statistics.

Rank 0 also computes the single-process references (no collectives) and prints one line `RESULT {json}`.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import fixtures as FX  # noqa: E402  (deterministic inputs only)
from vision_mtl_b200 import dist as vdist  # noqa: E402
from vision_mtl_b200 import ops  # noqa: E402
from vision_mtl_b200.graph_step import GraphedTrainStep, make_optimizer  # noqa: E402
from vision_mtl_b200.lit_module import MTLModule  # noqa: E402
from vision_mtl_b200.models.mtan_model import MTANMiniUnet  # noqa: E402

C, B, H, W = 19, 4, 32, 64


def build(dev):
    torch.manual_seed(5)
    net = MTANMiniUnet(3, {"depth": 1, "segm": C}, 128, 32, 3)
    net.to(dev).to(memory_format=torch.channels_last).train()
    return MTLModule(net, num_classes=C, device=dev)


def to_dev(batch, dev, sl=slice(None)):
    out = {k: v[sl].to(dev) for k, v in batch.items()}
    out["img"] = out["img"].contiguous(memory_format=torch.channels_last)
    return out


def eager_grads(module, batch):
    for p in module.parameters():
        p.grad = None
    loss = module.training_step(batch, 0)
    loss.backward()
    return (loss.detach().double().item(), module.last_confusion.clone(),
            {k: p.grad.detach().clone() for k, p in module.model.named_parameters()})


def fp64_oracle_grads(init, full):
    """Every parameter gradient of the step on the whole batch from the CPU oracle in float64 (tests only)."""
    from oracle import torch_port as TP

    p = {k: (v.double() if v.dtype == torch.float32 else v.clone()) for k, v in init.items()}
    for k, v in p.items():
        if v.dtype == torch.float64 and "running" not in k:
            v.requires_grad_(True)
    raw = TP.mtan_forward(p, full["img"].double(), training=True)
    res = TP.step_losses_and_metrics(raw, full["mask"], full["depth"].double(), C)
    res["loss"].backward()
    return {k: v.grad for k, v in p.items() if v.requires_grad}


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    mode = sys.argv[1]
    rank, local_rank, world = vdist.init_distributed("nccl")
    dev = torch.device("cuda", local_rank)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    full = FX.image_batch(B, H, W, C, "dist/batch" + (sys.argv[2] if len(sys.argv) > 2 else ""))
    per = B // world
    shard = to_dev(full, dev, slice(rank * per, (rank + 1) * per))

    # ---- the product path: DDP + (optional) global statistics + whole-step graph -------------------------------
    module = build(dev)
    if mode == "sync":
        vdist.enable_stat_sync(module)
    vdist.wrap_data_parallel(module, local_rank)
    opt = make_optimizer(module.parameters(), 1e-3, dev)

    def exchange():
        module.last_global_stats = vdist.allreduce_step_stats(module.last_confusion, module.last_step_scalars[0],
                                                              module.last_depth_sums)

    graphed = GraphedTrainStep(module, opt, shard, warmup=11, after_backward=exchange, preserve_state=True)
    graphed(shard)
    torch.cuda.synchronize()
    stats = module.last_global_stats
    g_conf = stats["confusion"].clone()
    g_loss = float(stats["loss"])
    inner = module.model.module
    g_grads = {k: p.grad.detach().clone() for k, p in inner.named_parameters()}
    g_running = {k: v.detach().clone() for k, v in inner.named_buffers() if "running" in k}
    dist.barrier()

    # ---- single-process references on rank 0 (no collectives) ---------------------------------------------------
    result = {"mode": mode, "world": world}
    ops.set_stat_sync(False)
    if rank == 0:
        ref = build(dev)
        if mode == "ddp":
            # training-mode BatchNorm: the running buffers moved by run 0 do not enter run 1's arithmetic
            runs = [eager_grads(ref, to_dev(full, dev, slice(r * per, (r + 1) * per))) for r in range(world)]
            conf = runs[0][1] + runs[1][1]
            loss = (runs[0][0] + runs[1][0]) / world
            grads = {k: (runs[0][2][k] + runs[1][2][k]) / world for k in runs[0][2]}
        else:
            init = {k: v.detach().cpu().clone() for k, v in ref.model.state_dict().items()}
            loss, conf, grads = eager_grads(ref, to_dev(full, dev))
            running = {k: v.detach().clone() for k, v in ref.model.named_buffers() if "running" in k}
            result["running_max_rel"] = max(rel_l2(g_running[k], running[k]) for k in running)
            # fp64 yardstick.  A random-init MTAN has ill-conditioned gradients (BatchNorm backward removes the mean and
            # the xhat-correlated part of nearly uniform SILog gradients: what is left is a small residual of large
            # terms) -- the reference's own fp32 arithmetic differs from fp64 by ~1e-3 here.  So the 2-GPU step is
            # held to being as close to the fp64 truth as the single-GPU step is.
            truth = fp64_oracle_grads(init, full)
            live64 = [k for k in grads if float(truth[k].norm()) > 1e-5 * sorted(float(v.norm()) for v in truth.values())[len(truth) // 2]]
            e_sync = sorted(rel_l2(g_grads[k].cpu(), truth[k]) for k in live64)
            e_whole = sorted(rel_l2(grads[k].cpu(), truth[k]) for k in live64)
            for name, e in (("sync", e_sync), ("whole", e_whole)):
                result[f"yard_{name}_median"] = e[len(e) // 2]
                result[f"yard_{name}_p90"] = e[(9 * len(e)) // 10]
            cat = lambda d: torch.cat([d[k].flatten().double().cpu() for k in live64])  # noqa: E731
            t64 = cat(truth)
            result["yard_sync_total"] = float((cat(g_grads) - t64).norm() / t64.norm())
            result["yard_whole_total"] = float((cat(grads) - t64).norm() / t64.norm())
        result["confusion_equal"] = bool(torch.equal(g_conf, conf))
        result["confusion_mismatch"] = int((g_conf - conf).abs().sum().item()) // 2
        result["pixels"] = int(conf.sum().item())
        result["loss_rel"] = abs(g_loss - loss) / abs(loss)
        # biases in front of a training-mode BatchNorm have an analytically zero gradient: what is left is round-off
        # of either run, and a relative error against it means nothing
        norms = sorted(float(grads[k].double().norm()) for k in grads)
        typical = norms[len(norms) // 2]
        live = [k for k in grads if float(grads[k].double().norm()) > 1e-5 * typical]
        result["grad_zero_params"] = len(grads) - len(live)
        result["grad_zero_max_abs"] = max([float(g_grads[k].abs().max()) for k in grads if k not in live] + [0.0])
        result["grad_typical_norm"] = typical
        errs = sorted(((rel_l2(g_grads[k], grads[k]), k) for k in live), reverse=True)
        result["grad_worst"] = errs[:6]
        result["grad_p90_rel_l2"] = errs[len(errs) // 10][0]
        result["grad_max_rel_l2"] = errs[0][0]
        result["grad_median_rel_l2"] = errs[len(errs) // 2][0]
        flat_g = torch.cat([g_grads[k].flatten().double() for k in grads])
        flat_r = torch.cat([grads[k].flatten().double() for k in grads])
        result["grad_total_rel_l2"] = float((flat_g - flat_r).norm() / flat_r.norm())
        print("RESULT " + json.dumps(result), flush=True)
    # tearing NCCL down after captured graphs referenced the communicator can block for minutes: line up and leave
    sys.stdout.flush()
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
