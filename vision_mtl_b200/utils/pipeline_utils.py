"""Model construction and checkpoint helpers (reference: vision_mtl/utils/pipeline_utils.py).

``build_model`` reproduces the three parameter-aligned configurations of the reference
(pipeline_utils.py:80-136); comet/TensorBoard plumbing is out of scope for the hot path.
"""
from __future__ import annotations

import argparse
import glob
import os
import re
import typing as t
from dataclasses import dataclass

import torch

from ..lit_module import MTLModule
from ..models.basic_model import BasicMTLModel
from ..models.cross_stitch_model import CSNet
from ..models.mtan_model import MTANMiniUnet
from .model_utils import get_model_with_dense_preds


@dataclass
class DataShape:
    """What the models need from the reference's ``DataConfig`` (cfg.py:96-146)."""

    num_classes: int
    height: int
    width: int
    name: str = "synthetic"


CITYSCAPES = DataShape(num_classes=19, height=128, width=256, name="cityscapes")
NYUV2 = DataShape(num_classes=14, height=256, width=256, name="nyuv2")  # 13 + background (cfg.py:124)


def build_model(args: argparse.Namespace, data_cfg) -> t.Union[BasicMTLModel, MTANMiniUnet, CSNet]:
    weights = getattr(args, "backbone_weights", "imagenet")
    name = args.model_name
    if name == "basic":
        return BasicMTLModel(segm_classes=data_cfg.num_classes, decoder_first_channel=540,
                             num_decoder_layers=5, encoder_weights=weights)
    if name == "mtan":
        return MTANMiniUnet(in_channels=3, map_tasks_to_num_channels={"depth": 1, "segm": data_cfg.num_classes},
                            task_subnets_hidden_channels=128, encoder_first_channel=32, encoder_num_channels=4)
    if name == "csnet":
        bp = dict(encoder_name="timm-mobilenetv3_large_100", encoder_weights=weights,
                  decoder_first_channel=256, num_decoder_layers=5)
        models = {
            "depth": get_model_with_dense_preds(segm_classes=1, activation=None, backbone_params=bp),
            "segm": get_model_with_dense_preds(segm_classes=data_cfg.num_classes, activation=None, backbone_params=bp),
        }
        return CSNet(models, channel_wise_stitching=getattr(args, "channel_wise_stitching", True),
                     stitch_mode=getattr(args, "stitch_mode", "reference_diag"))
    raise NotImplementedError(f"Unknown model name: {name}")


def init_model(args: argparse.Namespace, data_cfg) -> MTLModule:
    model = build_model(args, data_cfg)
    module = MTLModule(model=model, num_classes=data_cfg.num_classes, lr=args.lr, device=args.device)
    if getattr(args, "ckpt_dir", None):
        module.load_state_dict(load_ckpt_model(args.ckpt_dir)["model"])
    return module


def _unwrapped_state_dict(module: torch.nn.Module) -> dict:
    """state_dict with any DDP ``module.`` prefix removed, so files match the reference layout."""
    return {k.replace("model.module.", "model.", 1): v for k, v in module.state_dict().items()}


def save_ckpt(module, optimizer, scheduler, epoch: int, save_path_model: str, save_path_session: str, exp=None):
    """``model_{epoch}.pt = {"model": state_dict}``, ``session_{epoch}.pt`` = optimizer/scheduler/epoch
    (pipeline_utils.py:139-167)."""
    torch.save({"model": _unwrapped_state_dict(module)}, save_path_model)
    torch.save({"optimizer": optimizer.state_dict(), "scheduler": scheduler.state_dict(), "epoch": epoch},
               save_path_session)


def _latest(ckpt_dir: str, stem: str) -> str:
    paths = glob.glob(os.path.join(ckpt_dir, f"{stem}_*.pt"))
    if not paths:
        raise FileNotFoundError(f"no {stem}_*.pt under {ckpt_dir}")
    return max(paths, key=lambda p: int(re.search(rf"{stem}_(\d+)\.pt$", p).group(1)))


def load_ckpt_model(ckpt_dir: str) -> dict:
    return torch.load(_latest(ckpt_dir, "model"), map_location="cpu")


def load_ckpt_session(ckpt_dir: str) -> dict:
    return torch.load(_latest(ckpt_dir, "session"), map_location="cpu")
