"""Step module, drop-in for ``vision_mtl/lit_module.py`` (``MTLModule``).

Same public surface as the reference (constructor, ``training_step`` / ``validation_step`` /
``test_step`` / ``predict_step``, ``calc_losses``, ``calc_metrics``, ``postprocess_raw_out``,
``step_outputs``, epoch hooks, ``parameters()`` / ``to()``), without the pytorch-lightning and
torchmetrics dependencies (neither contributes arithmetic on this path: the reference drives the
module from a manual loop, ``self.log`` is a no-op without a Trainer).

``shared_step`` takes the fused route: the model's last features (MTAN) or logits (basic /
csnet) go straight into the fused head+loss+metric kernels, so logits are never re-read by a
separate softmax / argmax / log_softmax / one-hot pass and all five step scalars land in one
device buffer (``last_step_scalars``) -- one D2H copy per step instead of the reference's six
``.item()`` syncs.
"""
from __future__ import annotations

import typing as t

import torch
import torch.nn as nn

from . import ops
from .losses import SILogLoss
from .metrics import Accuracy, FBetaScore, JaccardIndex, MeanAbsoluteError
from .utils.loss_utils import summarize_epoch_metrics

STAGES = ("train", "val", "test", "predict")
STEP_KEYS = ("loss", "accuracy", "jaccard_index", "fbeta_score", "mae")


class MTLModule(nn.Module):
    def __init__(
        self,
        model: nn.Module,
        num_classes: int,
        optim_dict: t.Optional[dict] = None,
        lr: t.Optional[float] = None,
        device: t.Union[str, torch.device] = "cuda",
        loss_segm_weight: float = 1.0,
        loss_depth_weight: float = 1.0,
        ignore_index: int = -100,
        check_labels: bool = False,
    ):
        """``check_labels`` (debug): count, per step and on the device, the segmentation labels that are neither
        ``ignore_index`` nor in ``[0, num_classes)`` into ``last_bad_label_count``.  The fused kernels drop such
        pixels from the loss mean and the confusion matrix, where ``nn.CrossEntropyLoss`` / torchmetrics in the
        reference would raise: a ``num_classes`` / label-map mismatch is otherwise silent."""
        super().__init__()
        self.check_labels = bool(check_labels)
        self.last_bad_label_count: t.Optional[torch.Tensor] = None
        self.hparams = {"num_classes": num_classes, "lr": lr, "device": device,
                        "loss_segm_weight": loss_segm_weight, "loss_depth_weight": loss_depth_weight}
        self.num_classes = num_classes
        self.model = model
        self.ignore_index = ignore_index
        self.segm_criterion = nn.CrossEntropyLoss(ignore_index=ignore_index)  # config carrier; fused kernel computes it
        self.depth_criterion = SILogLoss()
        self.optim_dict = optim_dict
        self.loss_segm_weight = loss_segm_weight
        self.loss_depth_weight = loss_depth_weight
        self.step_outputs = {stage: {k: [] for k in STEP_KEYS} for stage in STAGES}
        self.metrics = {
            "accuracy": Accuracy(threshold=0.5, num_classes=num_classes, ignore_index=None, average="micro").to(device),
            "fbeta_score": FBetaScore(beta=1.0, threshold=0.5, num_classes=num_classes, average="weighted",
                                      ignore_index=None, mdmc_average="global").to(device),
            "jaccard_index": JaccardIndex(threshold=0.5, num_classes=num_classes, ignore_index=None).to(device),
            "mae": MeanAbsoluteError().to(device),
        }
        self.automatic_optimization = False
        self.last_step_scalars: t.Optional[torch.Tensor] = None  # [loss, acc, jaccard, fbeta, mae] on device
        self.last_confusion: t.Optional[torch.Tensor] = None     # int64 [C,C] of the last step
        self.last_depth_sums: t.Optional[torch.Tensor] = None    # float64 SILog / error moments of the last step

    # ------------------------------------------------------------------ forward / fused step
    def forward(self, x: torch.Tensor) -> dict:
        return self.model(x)

    def _inner_model(self) -> nn.Module:
        m = self.model
        return m.module if hasattr(m, "module") and isinstance(m.module, nn.Module) else m

    def fused_losses_and_metrics(self, img, gt_mask, gt_depth, want_preds: bool = False) -> dict:
        """Model forward + fused heads/losses/metrics.  Returns device scalars (no host sync)."""
        C = self.num_classes
        if self.check_labels:
            self.last_bad_label_count = ((gt_mask != self.ignore_index) & ((gt_mask < 0) | (gt_mask >= C))).sum()
        conf = torch.zeros((C, C), dtype=torch.int64, device=img.device)
        inner = self._inner_model()
        if hasattr(inner, "forward_features"):
            # MTAN: the 1x1 heads are fused with their losses, logits never reach HBM
            with ops.deferred_batch_counters():  # one multi-tensor add for every BatchNorm step counter
                feats = self.model(img, features_only=True)
            heads = inner.map_tasks_to_heads
            loss_segm, pred = ops.head_cross_entropy(feats["segm"], heads["segm"].weight, heads["segm"].bias,
                                                     gt_mask, self.ignore_index, conf, True)
            silog, mae, absrel, dpred, moments = ops.head_silog(
                feats["depth"], heads["depth"].weight, heads["depth"].bias, gt_depth,
                self.depth_criterion.min_depth, want_preds, return_moments=True)
        else:
            with ops.deferred_batch_counters():
                raw = self.model(img)
            loss_segm, pred = ops.cross_entropy_logits(raw["segm"], gt_mask, self.ignore_index, conf, True)
            silog, mae, absrel, dpred, moments = ops.head_silog(
                raw["depth"], None, None, gt_depth, self.depth_criterion.min_depth, want_preds, return_moments=True)
        seg = ops.seg_metrics(conf)
        loss = self.loss_segm_weight * loss_segm + self.loss_depth_weight * silog
        self.last_confusion = conf
        # {P, sum|p-t|, n(t > min_depth), sum|p-t|/t}: the depth part of the packed data-parallel statistics
        self.last_depth_sums = torch.stack([moments[7], moments[3], moments[0], moments[4]])
        return {
            "loss": loss, "loss_segm": loss_segm, "loss_depth": silog,
            "accuracy": seg[0], "jaccard_index": seg[1], "fbeta_score": seg[2], "mae": mae, "abs_rel": absrel,
            "segm_predictions": pred, "depth_predictions": dpred if want_preds else None,
        }

    def shared_step(self, batch: dict, stage: str) -> torch.Tensor:
        out = self.fused_losses_and_metrics(batch["img"], batch["mask"], batch["depth"])
        self.update_step_stats(stage, out, out)
        self.last_step_scalars = torch.stack([out[k].detach() for k in STEP_KEYS])
        return out["loss"]

    def update_step_stats(self, stage: str, all_losses: dict, all_metrics: dict) -> None:
        rec = self.step_outputs[stage]
        rec["loss"].append(all_losses["loss"].detach())
        for k in STEP_KEYS[1:]:
            rec[k].append(all_metrics[k])

    # ------------------------------------------------------------------ reference-shaped pieces
    def postprocess_raw_out(self, out: dict) -> dict:
        """lit_module.py:133-144 on full logits (API compatibility; the fused step never needs it)."""
        segm_logits, depth_logits = out["segm"], out["depth"]
        return {
            "segm_logits": segm_logits,
            "segm_predictions": torch.argmax(torch.softmax(segm_logits, dim=1), dim=1),
            "depth_predictions": torch.sigmoid(depth_logits).permute(0, 2, 3, 1),
        }

    def calc_losses(self, gt_mask: torch.Tensor, gt_depth: torch.Tensor, out: dict) -> dict:
        loss_segm, _ = ops.cross_entropy_logits(out["segm_logits"], gt_mask, self.ignore_index, None, False)
        loss_depth = self.depth_criterion(out["depth_predictions"], gt_depth)
        loss = self.loss_segm_weight * loss_segm + self.loss_depth_weight * loss_depth
        return {"loss": loss, "loss_segm": loss_segm, "loss_depth": loss_depth}

    def calc_metrics(self, gt_mask: torch.Tensor, gt_depth: torch.Tensor, out: dict) -> dict:
        preds = out["segm_predictions"]
        conf = ops.confusion_accumulate(preds, gt_mask, self.num_classes, None, self.ignore_index)
        seg = ops.seg_metrics(conf)
        for name in ("accuracy", "jaccard_index", "fbeta_score"):
            self.metrics[name].update_state(conf)
        mae = self.metrics["mae"](out["depth_predictions"], gt_depth)
        return {"accuracy": seg[0], "jaccard_index": seg[1], "fbeta_score": seg[2], "mae": mae}

    def training_step(self, batch: dict, batch_idx: t.Any = 0):
        return self.shared_step(batch, "train")

    def validation_step(self, batch: dict, batch_idx: t.Any = 0):
        return self.shared_step(batch, "val")

    def test_step(self, batch: dict, batch_idx: t.Any = 0):
        return self.shared_step(batch, "test")

    def predict_step(self, batch: dict, batch_idx: int = 0, dataloader_idx: int = 0):
        img = batch["img"]
        if "mask" in batch and "depth" in batch:
            out = self.fused_losses_and_metrics(img, batch["mask"], batch["depth"], want_preds=True)
            self.update_step_stats("predict", out, out)
            return {"segm": out["segm_predictions"].long(), "depth": out["depth_predictions"]}
        out = self.postprocess_raw_out(self(img))
        return {"segm": out["segm_predictions"], "depth": out["depth_predictions"]}

    # ------------------------------------------------------------------ epoch hooks / plumbing
    def log(self, *a, **k):  # Trainer-less, like the reference's manual loop
        return None

    def log_dict(self, *a, **k):
        return None

    def shared_epoch_end(self, stage: str) -> dict:
        return summarize_epoch_metrics(self.step_outputs[stage], metric_name_prefix=stage)

    def on_train_epoch_end(self):
        return self.shared_epoch_end("train")

    def on_validation_epoch_end(self):
        return self.shared_epoch_end("val")

    def on_test_epoch_end(self):
        return self.shared_epoch_end("test")

    def on_predict_epoch_end(self):
        return self.shared_epoch_end("predict")

    def configure_optimizers(self):
        if self.optim_dict:
            return self.optim_dict
        params = list(self.parameters())
        if params and params[0].is_cuda:  # the library's one-launch multi-tensor Adam (same arithmetic, same state_dict)
            from .optim import Adam

            optimizer = Adam(params, lr=self.hparams["lr"])
        else:
            optimizer = torch.optim.Adam(params=params, lr=self.hparams["lr"])
        scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer=optimizer, patience=5, factor=0.95)
        return {"optimizer": optimizer,
                "lr_scheduler": {"scheduler": scheduler, "interval": "epoch", "monitor": "train_loss"}}

    def transfer_batch_to_device(self, batch, device, dataloader_idx: int = 0):
        if isinstance(batch, dict):
            for key in batch:
                batch[key] = batch[key].to(device, non_blocking=True)
            return batch
        return batch.to(device, non_blocking=True)

    def parameters(self, recurse: bool = True):
        yield from self.model.parameters()

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        device = None
        for a in args:
            if isinstance(a, (str, torch.device)):
                device = a
        device = kwargs.get("device", device)
        if device is not None:
            for m in self.metrics.values():
                m.to(device)
        return self
