"""Training / prediction loop, drop-in for ``vision_mtl/training_lit.py``.

``run_pipe`` keeps the reference's signature and epoch structure (Adam + ReduceLROnPlateau,
train epoch, optional validation epoch under ``no_grad`` with BatchNorm left in training mode
exactly like the reference -- SURVEY F2 --, periodic checkpoints), with three changes:

* the five step scalars come back in ONE device->host copy per step
  (``module.last_step_scalars``) instead of six ``.item()`` syncs;
* ``ReduceLROnPlateau`` is built without the ``verbose`` argument torch 2.11 rejects (F11);
* under ``torchrun`` every rank runs the same loop on its batch shard: gradients are all-reduced
  by DDP, the step's confusion matrix, loss and depth sums by one packed all-reduce whose result (the
  GLOBAL loss / metrics) is what every rank records and rank 0 logs and saves;
* on CUDA the training step replays as one CUDA graph (``--no_graph_step`` keeps it eager);
  ``--sync_stats`` makes an N-GPU step equal the single-GPU step on the concatenated batch.

Experiment tracking (comet) and TensorBoard are optional duck-typed hooks (``exp.log_metric``,
``logger.log_metrics`` / ``logger.log_dir``); none is required.
"""
from __future__ import annotations

import argparse
import os
import typing as t
from collections import defaultdict

import torch

from . import dist as vdist
from .lit_module import STEP_KEYS, MTLModule
from .synthetic import make_batch
from .utils.loss_utils import print_metrics
from .utils.pipeline_utils import CITYSCAPES, NYUV2, DataShape, init_model, save_ckpt
from .utils.utils import parse_args


class SyntheticDataModule:
    """Minimal stand-in for ``MTLDataModule`` (lit_datamodule.py): fixed synthetic batches with the
    reference's tensor layouts; real datasets are host-side I/O outside the hot path."""

    def __init__(self, shape: DataShape, batch_size: int, steps_per_epoch: int = 4, val_steps: int = 2, seed: int = 11):
        self.shape, self.batch_size = shape, batch_size
        self.steps_per_epoch, self.val_steps, self.seed = steps_per_epoch, val_steps, seed
        self.benchmark_batch = None

    def _batches(self, n: int, offset: int):
        s = self.shape
        for i in range(n):
            yield make_batch(self.batch_size, s.height, s.width, s.num_classes, s.name, seed=self.seed + offset + i)

    def setup(self):
        return self

    def train_dataloader(self):
        return self._batches(self.steps_per_epoch, 0)

    def val_dataloader(self):
        return self._batches(self.val_steps, 10_000)

    def predict_dataloader(self):
        return self._batches(self.val_steps, 20_000)


def _log(logger, exp, values: dict, step: int) -> None:
    if logger is not None:
        logger.log_metrics(values, step=step)
    if exp:
        for k, v in values.items():
            exp.log_metric(k, v, step=step)


def _step_scalars_to_host(module: MTLModule, stage: str) -> dict:
    """One D2H copy for the step's loss and metrics; replaces the device scalars stored in
    ``step_outputs`` by host floats so epoch summaries need no further syncs.  Under data parallelism the
    GLOBAL values (all ranks: mean loss, metrics of the summed confusion matrix) travel in the same copy and
    are what gets recorded and logged."""
    glob = getattr(module, "last_global_scalars", None)
    if glob is not None:
        vals = torch.cat([module.last_step_scalars, glob]).tolist()[len(STEP_KEYS):]
    else:
        vals = module.last_step_scalars.tolist()
    rec = module.step_outputs[stage]
    for k, v in zip(STEP_KEYS, vals):
        rec[k][-1] = v
    return dict(zip(STEP_KEYS, vals))


def _make_metric_exchange(module: MTLModule, world: int) -> t.Optional[t.Callable[[], None]]:
    """The step's ONE metric collective (dist.allreduce_step_stats); ``None`` for a single process."""
    if world <= 1:
        module.last_global_scalars = None
        return None

    def exchange() -> None:
        stats = vdist.allreduce_step_stats(module.last_confusion, module.last_step_scalars[0], module.last_depth_sums)
        module.last_global_stats = stats
        module.last_global_scalars = torch.stack(
            [stats[k].to(torch.float32) for k in ("loss", "accuracy", "jaccard_index", "fbeta_score", "mae")])

    return exchange


def _same_shapes(batch: dict, static: dict) -> bool:
    return all(k in static and batch[k].shape == static[k].shape and batch[k].dtype == static[k].dtype for k in batch)


def run_pipe(
    args: argparse.Namespace,
    module: MTLModule,
    datamodule,
    num_epochs: int,
    device: t.Union[str, torch.device],
    exp=None,
    logger=None,
) -> t.Dict[str, t.Dict[str, list]]:
    """Train for ``num_epochs``; returns the per-epoch train / val metric histories.

    On CUDA (``args.graph_step``, default on) the training step -- forward, fused losses / metrics, backward,
    DDP's gradient all-reduce, the packed metric all-reduce and Adam -- replays as ONE CUDA graph
    (``graph_step.GraphedTrainStep``); batches whose shape differs from the captured one (a short last batch)
    run the same step eagerly.  ``args.sync_stats`` turns on the global-batch-exact mode: every
    batch statistic (BatchNorm moments forward and backward, SILog moments) is all-reduced between the two
    phases of its kernel pair, so an N-GPU step equals the single-GPU step on the concatenated batch."""
    rank, _, world = vdist.env_world()
    is_main = rank == 0
    on_cuda = str(device).startswith("cuda")
    module.to(device)
    if on_cuda:
        module.model.to(memory_format=torch.channels_last)
    if world > 1:
        if getattr(args, "sync_stats", False):
            vdist.enable_stat_sync(module)
        vdist.wrap_data_parallel(module, torch.device(device).index)
    use_graph = on_cuda and bool(getattr(args, "graph_step", True))
    if on_cuda:  # capturable Adam with a device-tensor learning rate: the same object serves graph and eager steps
        from .graph_step import GraphedTrainStep, make_optimizer

        optimizer = make_optimizer(module.parameters(), args.lr, device)
    else:
        optimizer = torch.optim.Adam(module.parameters(), lr=args.lr)
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, patience=2, factor=0.9)
    module.model.train()
    exchange = _make_metric_exchange(module, world)
    graphed = None

    def eager_step(batch: dict) -> None:
        optimizer.zero_grad(set_to_none=True)
        loss = module.training_step(batch, batch_idx=0)
        loss.backward()
        if exchange is not None:
            exchange()
        optimizer.step()

    epoch_metrics = {"train": defaultdict(list), "val": defaultdict(list)}
    global_step = val_step = 0
    for epoch in range(num_epochs):
        if is_main:
            print(f"### Epoch {epoch + 1}/{num_epochs} ###\n---TRAIN---")
        for batch in datamodule.train_dataloader():
            if world > 1:
                batch = vdist.shard_batch(batch, rank, world)
            batch = module.transfer_batch_to_device(batch, device, 0)
            if use_graph and graphed is None:
                # DDP rebuilds its buckets during the first iterations: capture after they have settled
                graphed = GraphedTrainStep(module, optimizer, batch, warmup=11 if world > 1 else 3,
                                           after_backward=exchange, preserve_state=True, record_step_outputs=True)
            if graphed is not None and _same_shapes(batch, graphed.static):
                graphed(batch)
            else:
                eager_step(batch)
            scal = _step_scalars_to_host(module, "train")
            if is_main:
                _log(logger, exp, {f"step/train/{k}": v for k, v in scal.items()}, global_step)
                print_metrics("train", scal)
            global_step += 1
        train_epoch = module.on_train_epoch_end()
        for k, v in train_epoch.items():
            epoch_metrics["train"][k].append(v)
        if is_main:
            print_metrics("epoch", train_epoch)
            _log(logger, exp, {f"epoch/{k}": v for k, v in train_epoch.items()}, epoch)

        if (epoch + 1) % args.val_epoch_freq == 0:
            if is_main:
                print("---VAL---")
            val_loss = 0.0
            with torch.no_grad():  # BN stays in training mode here, as in the reference (F2)
                for batch in datamodule.val_dataloader():
                    if world > 1:
                        batch = vdist.shard_batch(batch, rank, world)
                    batch = module.transfer_batch_to_device(batch, device, 0)
                    module.validation_step(batch, batch_idx=0)
                    if exchange is not None:
                        exchange()
                    scal = _step_scalars_to_host(module, "val")
                    val_loss += scal["loss"]
                    if is_main:
                        _log(logger, exp, {f"step/val/{k}": v for k, v in scal.items()}, val_step)
                    val_step += 1
            val_epoch = module.on_validation_epoch_end()
            for k, v in val_epoch.items():
                epoch_metrics["val"][k].append(v)
            if is_main:
                print_metrics("epoch/val", val_epoch)
                _log(logger, exp, {f"epoch/{k}": v for k, v in val_epoch.items()}, epoch)
            # the SUM of val batch losses, as in the reference; under data parallelism it is the sum of the
            # GLOBAL losses, identical on every rank, so every replica takes the same plateau decision
            scheduler.step(val_loss)

        last = epoch == getattr(args, "num_epochs", num_epochs) - 1
        if is_main and logger is not None and ((epoch + 1) % args.save_epoch_freq == 0 or last):
            save_ckpt(module, optimizer, scheduler, epoch,
                      os.path.join(logger.log_dir, f"model_{epoch}.pt"),
                      os.path.join(logger.log_dir, f"session_{epoch}.pt"), exp=exp)
    return epoch_metrics


@torch.no_grad()
def predict(predict_dataloader, module: MTLModule, device, do_plot_preds: bool = False, exp=None,
            do_show_preds: bool = False):
    """Eval-mode predictions + metrics (training_lit.py:186-216; the one place running statistics
    are used, so the gate takes its single-pass folded-BN path)."""
    preds = []
    module.eval()
    module.to(device)
    for batch in predict_dataloader:
        batch = module.transfer_batch_to_device(batch, device, 0)
        preds.append(module.predict_step(batch, 0, 0))
    return preds, module.on_predict_epoch_end()


class _DirLogger:
    def __init__(self, log_dir: str):
        self.log_dir = log_dir
        os.makedirs(log_dir, exist_ok=True)

    def log_metrics(self, values: dict, step: int) -> None:
        pass


def main(argv: t.Optional[t.Sequence[str]] = None):
    args = parse_args(argv)
    rank, local_rank, world = vdist.init_distributed()
    if args.device.startswith("cuda") and world > 1:
        args.device = f"cuda:{local_rank}"
    torch.backends.cudnn.allow_tf32 = bool(args.conv_tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(args.conv_tf32)
    from . import ops

    ops.default_gate_precision = args.gate_precision
    shape = NYUV2 if args.dataset_name == "nyuv2" else CITYSCAPES
    torch.manual_seed(11)
    module = init_model(args, shape)
    datamodule = SyntheticDataModule(shape, args.batch_size * world).setup()
    logger = _DirLogger(os.path.join("lightning_logs", f"training-{args.model_name}", args.run_name or "run"))
    metrics = run_pipe(args, module, datamodule, args.num_epochs, args.device, exp=None, logger=logger)
    preds, predict_metrics = predict(datamodule.predict_dataloader(), module, args.device)
    if rank == 0:
        print_metrics("predict", predict_metrics)
    return metrics


if __name__ == "__main__":
    main()
