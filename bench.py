#!/usr/bin/env python
"""Benchmark of the multi-task training step (BASELINE.json metric: MTL train-step images/sec;
fused-kernel HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W [--workload all|mtan|csnet|mtan_nyu] [--impl reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N > 1).  Prints ONE JSON line.
With no ``--workload`` the three GPU configurations of BASELINE.json run back to back in one process:
the headline record is ``mtan`` (configs[2]: attention-gated step, Cityscapes-shaped 128x256, per-GPU batch
32 -- the configuration whose step is built on the tcgen05 gate and fused-head kernels); ``csnet``
(configs[1]) and ``mtan_nyu`` (configs[3]) are complete sub-records under ``workloads``.

* ``value``  : images/s, whole job, batch resident in HBM, CUDA-event timed, max over ranks.
* ``e2e``    : same step through the public API with the batch copied from pinned host memory and
               the five step scalars read back every step.
* ``roofline``: the hand-written kernel with the largest share of the step: ALGORITHMIC bytes of SURVEY 8(d)
               / CUDA-event time (external events around every library call inside an instrumented copy of
               the step graph, same method at every N), against MEASURED_PEAKS.json; ``issued_bytes`` (what
               the kernels request) and ``traffic`` (ncu dram bytes of a captured launch) reported beside it.
* ``cpu_baseline`` (rank 0, N = 1) and ``--impl reference``: the reference's OWN loop
               (``training_lit.run_pipe`` -> ``lit_module.MTLModule``, unmodified, from /root/reference or
               the shipped baseline/_ref copy; import stubs in oracle/ref_runtime.py) timed on the host cores
               on a bounded sample of the same workload; the oracle port is timed as a cross-check.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import torch  # noqa: E402

WORKLOADS = {
    # name: (model, dataset, classes, H, W, per-GPU batch)   -- BASELINE.json configs[1..3]
    "csnet": ("csnet", "cityscapes", 19, 128, 256, 32),
    "mtan": ("mtan", "cityscapes", 19, 128, 256, 32),
    "mtan_nyu": ("mtan", "nyuv2", 14, 256, 256, 16),
}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-graph-events", action="store_true",
                    help="skip the instrumented graph copy; per-op times then come from the eager pass")
    ap.add_argument("--workload", choices=["all", *WORKLOADS], default="all",
                    help="all = mtan (headline) + csnet + mtan_nyu sub-records")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's)")
    ap.add_argument("--conv-tf32", action="store_true", help="allow TF32 in the cuDNN 3x3 convs (default: strict fp32)")
    ap.add_argument("--conv-3xtf32", action="store_true",
                    help="EXPERIMENTAL: fp32-grade convolutions as channel-stacked 3xTF32 cuDNN calls (vision_mtl_b200/conv3x.py)")
    ap.add_argument("--gate-precision", default="tc_3xtf32", choices=["tc_3xtf32", "tc_tf32", "fp32_ffma"])
    ap.add_argument("--stitch-mode", default="reference_diag", choices=["reference_diag", "full_mix"])
    ap.add_argument("--cpu-sample-batch", type=int, default=8)
    ap.add_argument("--cpu-steps", type=int, default=5, help="timed steps of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lr", type=float, default=5e-4)
    ap.add_argument("--no-graph", action="store_true", help="eager step instead of the whole-step CUDA graph")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# reference CPU path (oracle port), used by cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------
def cpu_reference_step_fn(workload: str, batch_size: int, lr: float):
    """Build the reference's CPU training step (fwd + CE/SILog + metrics + bwd + Adam) for a
    ``batch_size`` sample of the workload; returns (step_fn, n_threads)."""
    from oracle import torch_port as TP
    from vision_mtl_b200.synthetic import make_batch
    from vision_mtl_b200.utils.model_utils import get_model_with_dense_preds

    model, dataset, C, H, W, _ = WORKLOADS[workload]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(11)
    batch = make_batch(batch_size, H, W, C, dataset, seed=11)
    if model == "csnet":
        nets = {"depth": get_model_with_dense_preds(1, None, dict(encoder_weights=None)),
                "segm": get_model_with_dense_preds(C, None, dict(encoder_weights=None))}
        net = TP.CSNetOracle(nets, channel_wise_stitching=True).train()
        params = list(net.parameters())
        fwd = lambda img: net(img)  # noqa: E731
    else:
        from vision_mtl_b200.models.mtan_model import MTANMiniUnet  # structure / init only

        skel = MTANMiniUnet(3, {"depth": 1, "segm": C}, 128, 32, 4)
        p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
             for k, v in skel.state_dict().items()}
        params = [v for v in p.values() if v.requires_grad]
        fwd = lambda img: TP.mtan_forward(p, img, training=True)  # noqa: E731
    opt = torch.optim.Adam(params, lr=lr)

    def step():
        opt.zero_grad()
        raw = fwd(batch["img"])
        res = TP.step_losses_and_metrics(raw, batch["mask"], batch["depth"], C)
        res["loss"].backward()
        opt.step()
        return float(res["loss"].detach())

    return step, threads


def time_cpu_port(workload: str, batch_size: int, lr: float, steps: int, warmup: int):
    """Oracle port of the reference's CPU path (cross-check of the reference-loop timing)."""
    step, threads = cpu_reference_step_fn(workload, batch_size, lr)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch_size * steps / dt, dt / steps, threads


def time_cpu_reference(workload: str, batch_size: int, lr: float, steps: int, warmup: int):
    """The reference's own ``run_pipe`` loop on a batch-``batch_size`` sample: (images/s, s/step, threads, kind)."""
    from vision_mtl_b200.synthetic import make_batch

    model, dataset, C, H, W, _ = WORKLOADS[workload]
    batch = make_batch(batch_size, H, W, C, dataset, seed=11)
    try:
        from oracle import ref_runtime

        ips, sec, threads = ref_runtime.time_reference_loop(model, batch, C, lr, steps, warmup)
        return ips, sec, threads, "reference"
    except Exception as exc:  # no reference tree on this box: time the port and say so
        sys.stderr.write(f"[bench] reference loop unavailable ({exc!r}); timing the oracle port\n")
        ips, sec, threads = time_cpu_port(workload, batch_size, lr, steps, warmup)
        return ips, sec, threads, "port"


def workload_text(name: str, batch: int) -> str:
    model, dataset, C, H, W, _ = WORKLOADS[name]
    return (f"{name}: {model} training step (fwd + fused CE/SILog/metrics + bwd + Adam), "
            f"{dataset}-shaped {H}x{W}, {C} classes, per-GPU batch {batch}")


def cpu_baseline_record(name: str, args, steps: int, with_port: bool) -> dict:
    model, dataset, C, H, W, _ = WORKLOADS[name]
    bs = args.cpu_sample_batch
    ips, sec, threads, kind = time_cpu_reference(name, bs, args.lr, steps, 1)
    what = ("the reference's own training_lit.run_pipe + lit_module.MTLModule (unmodified; torchmetrics "
            "restated, see oracle/ref_runtime.py)") if kind == "reference" else "oracle port of the reference's PyTorch CPU path"
    rec = {"value": ips, "unit": "images/s", "cores": threads, "kind": kind,
           "sample": f"1 warm-up + {steps} timed {model} train steps on a batch-{bs} sample of the same {H}x{W} workload "
                     f"({sec:.2f} s/step), {what}, fp32, torch CPU"}
    if with_port and kind == "reference":
        pips, psec, _ = time_cpu_port(name, bs, args.lr, 2, 1)
        rec["port_cross_check"] = {"value": pips, "unit": "images/s", "sample": f"oracle port, 2 timed steps ({psec:.2f} s/step)"}
    return rec


# ------------------------------------------------------------------------------------------
# clocks sampling
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
HEADLINE = "mtan"
ALL_ORDER = ("mtan", "csnet", "mtan_nyu")


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    names = list(ALL_ORDER) if args.workload == "all" else [args.workload]
    recs = {}
    for i, name in enumerate(names):
        model, dataset, C, H, W, B = WORKLOADS[name]
        bs = args.cpu_sample_batch
        steps = max(args.steps, 1) if i == 0 else max(min(args.steps, args.cpu_steps), 1)  # sub-records: bounded
        ips, sec, threads, kind = time_cpu_reference(name, bs, args.lr, steps, min(args.warmup, 1))
        what = ("reference training_lit.run_pipe + lit_module.MTLModule, unmodified (torchmetrics restated)"
                if kind == "reference" else "oracle port of the reference's PyTorch CPU path")
        sample = (f"{model} train step (fwd+CE/SILog+metrics+bwd+Adam) on a batch-{bs} sample of the {H}x{W} workload, "
                  f"{steps} timed steps, {what}, fp32, torch CPU")
        recs[name] = {
            "impl": "reference", "metric": "mtl_train_step_images_per_sec", "value": ips, "unit": "images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(name, B if not args.batch else args.batch),
                       "global_batch": (args.batch or B) * args.gpus, "parallelism": f"dp{args.gpus}",
                       "conv_math": "fp32 (torch CPU)", "cpu_sample_batch": bs,
                       "note": "the reference is single-process CPU code: rank 0 times it on all host cores, "
                               "throughput does not depend on n_gpus"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
    line = recs[names[0]]
    if len(names) > 1:
        line["workloads"] = {n: recs[n] for n in names[1:]}
    print(json.dumps(line))


_T0 = time.perf_counter()


def _phase(msg: str) -> None:
    """Progress line on stderr (rank 0): where the wall time of a bench run goes."""
    if int(os.environ.get("RANK", 0)) == 0:
        sys.stderr.write(f"[bench +{time.perf_counter() - _T0:7.1f}s] {msg}\n")
        sys.stderr.flush()


def run_workload(args, name: str, rank: int, local_rank: int, world: int, with_port: bool) -> dict:
    """One workload end to end on this rank; returns the record (complete on rank 0)."""
    import gc

    import torch.distributed as dist

    from vision_mtl_b200 import dist as vdist
    from vision_mtl_b200 import ops
    from vision_mtl_b200.graph_step import GraphedTrainStep
    from vision_mtl_b200.synthetic import make_batch
    from vision_mtl_b200.utils.pipeline_utils import DataShape, init_model

    dev = torch.device("cuda", local_rank)
    model, dataset, C, H, W, B = WORKLOADS[name]
    if args.batch:
        B = args.batch
    torch.manual_seed(11)  # identical weights on every rank
    margs = argparse.Namespace(model_name=model, backbone_weights=None, channel_wise_stitching=True,
                               stitch_mode=args.stitch_mode, lr=args.lr, device=dev, ckpt_dir=None)
    module = init_model(margs, DataShape(num_classes=C, height=H, width=W, name=dataset))
    module.to(dev)
    module.model.to(memory_format=torch.channels_last)
    module.model.train()
    use_graph = not args.no_graph
    if world > 1:
        side = torch.cuda.Stream()  # DDP built on a side stream so its buckets are capture-friendly
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            vdist.wrap_data_parallel(module, local_rank)
        torch.cuda.current_stream().wait_stream(side)
    from vision_mtl_b200.optim import Adam  # the library's one-launch multi-tensor Adam (capture-safe)

    opt = Adam(module.parameters(), lr=args.lr)
    _phase(f"[{name}] model + optimizer built")

    host = make_batch(B, H, W, C, dataset, seed=11 + rank, pin=True)          # pinned host copy
    resident = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    resident["img"] = resident["img"].contiguous(memory_format=torch.channels_last)
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def metric_exchange():
        if world > 1:  # the one metric exchange of the step (confusion matrix + loss)
            module.last_global_stats = vdist.allreduce_step_stats(module.last_confusion, module.last_step_scalars[0])

    def eager_step(batch):
        opt.zero_grad(set_to_none=True)
        loss = module.training_step(batch, 0)
        loss.backward()
        metric_exchange()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        eager_step(resident)
    barrier()
    _phase(f"[{name}] eager warm-up done")

    # launches per step (the `gpu_launches` claim): counted on one eager pass of the identical step
    ops.reset_launch_count()
    eager_step(resident)
    barrier()
    launches_per_step = ops.launch_count()

    sampler = ClockSampler(local_rank)
    graphed = None
    if use_graph:
        graphed = GraphedTrainStep(module, opt, resident, warmup=11 if world > 1 else 3, after_backward=metric_exchange)
        for _ in range(3):
            graphed()
        barrier()
        _phase(f"[{name}] step graph captured and replayed")

    def train_step(batch):
        if graphed is not None:
            return graphed(None if batch is resident else batch)
        return eager_step(batch)

    # e2e input pipeline (graph mode): every step's batch is copied from pinned host memory on a copy stream
    # into one of two device staging buffers WHILE the previous step's graph runs, then moved into the
    # graph's static inputs by a device copy -- the usual prefetching loader, one H2D of the full batch per
    # step inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    staging = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    staged = [torch.cuda.Event(), torch.cuda.Event()]     # H2D into staging[b] complete
    consumed = [torch.cuda.Event(), torch.cuda.Event()]   # staging[b] copied into the static inputs
    e2e_state = {"i": 0, "primed": False}

    def prefetch(b):
        copy_stream.wait_event(consumed[b])
        with torch.cuda.stream(copy_stream):
            for k, v in host.items():
                staging[b][k].copy_(v, non_blocking=True)
            staged[b].record(copy_stream)

    def e2e_step():
        if graphed is not None:
            if not e2e_state["primed"]:
                for ev in consumed:
                    ev.record()
                prefetch(0)
                e2e_state["primed"] = True
            b = e2e_state["i"] & 1
            e2e_state["i"] += 1
            torch.cuda.current_stream().wait_event(staged[b])
            graphed.load(staging[b])          # device copy into the static inputs
            consumed[b].record()
            prefetch(b ^ 1)                   # next step's H2D overlaps this step's graph
            graphed()
        else:
            eager_step({k: v.to(dev, non_blocking=True) for k, v in host.items()})
        return module.last_step_scalars.cpu()  # one D2H copy: loss + 4 metrics

    # ---- timed region: batch resident in HBM ---------------------------------------------------
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        train_step(resident)
    ev1.record()
    barrier()
    launches = launches_per_step * args.steps
    ms = ev0.elapsed_time(ev1)
    _phase(f"[{name}] timed region done")
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host batch in, step scalars out, every step ------------------------------------
    for _ in range(2):
        e2e_step()
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for _ in range(args.steps):
        scal = e2e_step()
    ev3.record()
    barrier()
    ms_e2e = ev2.elapsed_time(ev3)
    _phase(f"[{name}] e2e region done")

    # ---- per-op device time: external CUDA events around every library call inside an instrumented copy
    # of the step graph, replayed right after the timed region.  Same method at every N (the copy contains
    # the same NCCL nodes as the timed graph, so all ranks capture and replay it together).
    kstats, ksteps, events_from = {}, 1, None
    if graphed is not None and not args.no_graph_events:
        try:
            prof = GraphedTrainStep(module, opt, resident, warmup=1, after_backward=metric_exchange, profile=True)
            for _ in range(3):
                prof()
            barrier()
            gstats = prof.kernel_stats()
            if gstats and all(v["ms"] > 0 for v in gstats.values()):
                kstats = gstats
                events_from = ("external CUDA events around every library call inside an instrumented copy of the "
                               "step graph (third replay, right after the timed region)")
            del prof
        except Exception as exc:
            sys.stderr.write(f"[bench] in-graph event pass unavailable: {exc!r}\n")
    if not kstats:  # eager mode, or the instrumented capture failed: events around the calls of eager steps
        n = min(args.steps, 5)
        with ops.kernel_timing() as records:
            for _ in range(n):
                torch.cuda._sleep(100_000_000)  # keep the device behind the host: events bracket kernel execution
                eager_step(resident)
            barrier()
        kstats, ksteps = ops.summarize_timing(records), n
        events_from = "eager pass of the identical step, device kept busy ahead of the host so events bracket kernel execution"
    _phase(f"[{name}] per-kernel event pass done")

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    rec = None
    if rank == 0:
        ips = B * world * args.steps / (ms * 1e-3)
        ips_e2e = B * world * args.steps / (ms_e2e * 1e-3)
        peak, peak_src = measured_peak()
        top = max(kstats.items(), key=lambda kv: kv[1]["ms"]) if kstats else None
        roofline = None
        if top is not None:
            kname, d = top
            traffic, traffic_note = None, None
            tpath = os.path.join(REPO, "profiles", "ncu_traffic.json")
            if os.path.exists(tpath):  # dram__bytes_read+write of ONE captured launch (ncu --set full)
                trec = json.load(open(tpath)).get(name, {}).get(kname)
                if trec:
                    traffic = trec["bytes"]
                    traffic_note = (f"captured launch: {trec['site']}; algorithmic bytes of that launch "
                                    f"{trec['algorithmic_bytes_same_launch']}")
            roofline = {"bound": "hbm", "kernel": "vmtl_" + kname, "achieved": d["gbps"], "peak": peak, "unit": "GB/s",
                        "frac": d["gbps"] / peak, "traffic": traffic, "traffic_note": traffic_note,
                        "peak_source": peak_src, "events_from": events_from,
                        "bytes_model": "SURVEY 8(d) algorithmic bytes (ops.gate_bytes and the _call sites of ops.py)",
                        "launches_timed": d["calls"], "avg_launch_ms": d["ms"] / d["calls"],
                        "algorithmic_bytes_per_launch": d["bytes"] / d["calls"],
                        "issued_bytes_per_launch": d["issued_bytes"] / d["calls"],
                        "frac_on_issued_bytes": d["issued_bytes"] / (d["ms"] * 1e-3) / 1e9 / peak}
        rec = {
            "metric": "mtl_train_step_images_per_sec", "value": ips, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_text(name, B),
                "global_batch": B * world, "parallelism": f"dp{world}",
                "conv_math": ("3xTF32 channel-stacked cuDNN tensor-core convolutions (fp32-grade, experimental)" if args.conv_3xtf32
                              else "tf32 (cuDNN)" if args.conv_tf32 else "fp32 (cuDNN, TF32 off)"),
                "gate_precision": args.gate_precision, "stitch_mode": args.stitch_mode,
                "step_launch": "one CUDA graph per step (fwd + fused losses/metrics + bwd + Adam)" if graphed is not None else "eager",
                "l2": "per-step working set (activations of a batch-%d step) >> 126 MB L2; no explicit flush" % B,
                "backbone": "stand-in MobileNetV3-Large/Unet (smp/timm not installable offline)" if model != "mtan" else "reference MTAN mini-UNet",
            },
            "e2e": {"value": ips_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes * world,
                    "d2h_bytes_per_step": int(scal.numel() * scal.element_size()) * world,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "kernels": {k: {"calls": v["calls"], "ms_per_step": v["ms"] / ksteps, "gbps": v["gbps"],
                            "frac_of_peak": v["gbps"] / peak,
                            "issued_over_algorithmic_bytes": v["issued_bytes"] / max(v["bytes"], 1)}
                        for k, v in sorted(kstats.items())},
            "hand_written_ms_per_step": sum(v["ms"] for v in kstats.values()) / ksteps,
        }
    # release this workload's graphs / buffers before the next one is built
    graphed = None
    del module, opt, resident, staging, host
    gc.collect()
    barrier()
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rec["cpu_baseline"] = cpu_baseline_record(name, args, args.cpu_steps, with_port)
        _phase(f"[{name}] cpu baseline done")
    return rec


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    from vision_mtl_b200 import dist as vdist
    from vision_mtl_b200 import ops

    rank, local_rank, world = vdist.init_distributed()
    _phase(f"process group ready (world {world})")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    torch.backends.cudnn.allow_tf32 = bool(args.conv_tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(args.conv_tf32)
    torch.backends.cudnn.benchmark = True
    if args.conv_3xtf32:
        from vision_mtl_b200 import conv3x

        conv3x.enable(True)
    ops.default_gate_precision = args.gate_precision

    names = list(ALL_ORDER) if args.workload == "all" else [args.workload]
    recs = {}
    for i, name in enumerate(names):
        recs[name] = run_workload(args, name, rank, local_rank, world, with_port=(i == 0))
    if rank == 0:
        line = recs[names[0]]
        if len(names) > 1:
            line["workloads"] = {n: recs[n] for n in names[1:]}
        print(json.dumps(line))
    if world > 1:
        # Tearing NCCL down after captured graphs referenced the communicator can block for minutes
        # (destroy_process_group waits on work the graphs owned).  Everything is measured and printed:
        # line the ranks up and leave without the teardown.
        sys.stdout.flush()
        sys.stderr.flush()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
