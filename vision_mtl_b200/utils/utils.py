"""CLI arguments and small helpers (reference: vision_mtl/utils/utils.py)."""
from __future__ import annotations

import argparse
import functools
import typing as t

import torch


def build_arg_parser() -> argparse.ArgumentParser:
    """Same flags and defaults as the reference ``parse_args`` (utils.py:8-49) plus the knobs of
    this implementation (``--stitch_mode``, ``--gate_precision``, ``--conv_tf32``)."""
    p = argparse.ArgumentParser()
    g = p.add_argument_group("pipe")
    for flag in ("--do_overfit", "--do_optimize", "--do_plot_preds", "--do_show_preds", "--exp_disabled"):
        g.add_argument(flag, action="store_true")
    g.add_argument("--ckpt_dir")
    g.add_argument("--run_name")
    g.add_argument("--device", default="cuda:0")
    g.add_argument("--exp_tags", nargs="*", default=[])
    g = p.add_argument_group("model")
    g.add_argument("--model_name", choices=["basic", "mtan", "csnet"], default="basic")
    g.add_argument("--backbone_weights", choices=["imagenet"])
    g.add_argument("--channel_wise_stitching", action="store_true")
    g.add_argument("--stitch_mode", choices=["reference_diag", "full_mix"], default="reference_diag")
    g.add_argument("--gate_precision", choices=["tc_3xtf32", "tc_tf32", "fp32_ffma"], default="tc_3xtf32")
    g.add_argument("--conv_tf32", action="store_true", help="let cuDNN use TF32 for the 3x3 convs")
    g.add_argument("--no_graph_step", dest="graph_step", action="store_false",
                   help="run the training step eagerly instead of replaying one CUDA graph per step")
    g.add_argument("--sync_stats", action="store_true",
                   help="data parallel: all-reduce every batch statistic (BatchNorm, SILog) so that an N-GPU "
                        "step equals the single-GPU step on the concatenated batch")
    g = p.add_argument_group("data")
    g.add_argument("--dataset_name", choices=["cityscapes", "nyuv2", "synthetic"], default="cityscapes")
    g.add_argument("--batch_size", type=int, default=1)
    g.add_argument("--num_workers", type=int, default=0)
    g = p.add_argument_group("opt")
    g.add_argument("--n_trials", type=int, default=7)
    g.add_argument("--n_jobs", type=int, default=2)
    g = p.add_argument_group("trainer")
    g.add_argument("--lr", type=float, default=5e-3)
    g.add_argument("--loss_segm_weight", type=float, default=1)
    g.add_argument("--loss_depth_weight", type=float, default=1)
    g.add_argument("--num_epochs", type=int, default=10)
    g.add_argument("--val_epoch_freq", type=int, default=1)
    g.add_argument("--save_epoch_freq", type=int, default=10)
    return p


def parse_args(argv: t.Optional[t.Sequence[str]] = None) -> argparse.Namespace:
    args, _ = build_arg_parser().parse_known_args(argv)
    return args


def get_module_by_name(module: torch.nn.Module, access_string: str) -> torch.nn.Module:
    """Resolve a dotted ``named_modules`` path (works through ``Sequential`` indices)."""
    return functools.reduce(getattr, access_string.split("."), module)


def update_args(args: argparse.Namespace, kv_map: t.Dict[str, float]) -> argparse.Namespace:
    for k, v in kv_map.items():
        if not hasattr(args, k):
            raise AttributeError(f"unknown argument {k!r}")
        setattr(args, k, v)
    return args
