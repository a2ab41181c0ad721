"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
``include/vmtl_b200.h`` declares; the host-side surface mirrors the reference's class surface."""
import ctypes
import inspect
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(REPO, "include", "vmtl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vmtl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from vision_mtl_b200 import _lib, build

    path = build.build()
    lib = ctypes.CDLL(path)
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in vmtl_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table and header disagree"
    typed = _lib.load()
    assert typed.vmtl_version() >= 100
    assert typed.vmtl_strerror(-5).decode() == "workspace too small"
    assert typed.vmtl_strerror(0).decode() == "ok"
    # sizing helpers are host-only and must work without a GPU
    assert typed.vmtl_xstitch_bwd_workspace_bytes(2, 1 << 20, 32, 1) > 0
    # the tensor-core backward never materialises dz: its workspace holds per-CTA partials only; the CUDA-core
    # backward (precision 0) stages dz [M,N]
    assert 0 < typed.vmtl_gate_workspace_bytes(1 << 20, 128, 32, 1, 1) < 4 * (1 << 20) * 32
    assert typed.vmtl_gate_workspace_bytes(1 << 20, 128, 32, 0, 1) > 4 * (1 << 20) * 32
    assert typed.vmtl_gate_workspace_bytes(1 << 20, 100, 32, 1, 1) > 0   # any K: CUDA-core contraction
    assert typed.vmtl_gate_workspace_bytes(1 << 20, 128, 30, 1, 1) == 0  # N must be a multiple of 4
    assert typed.vmtl_loss_workspace_bytes(1 << 20) > 0


def test_no_cpu_fallback():
    """The product path must fail loudly off-GPU instead of silently computing elsewhere."""
    from vision_mtl_b200 import _lib, ops

    x = [torch.randn(1, 4, 2, 2), torch.randn(1, 4, 2, 2)]
    with pytest.raises(_lib.VmtlError):
        ops.cross_stitch(x, torch.rand(2, 2), "reference_diag")
    with pytest.raises(_lib.VmtlError):
        ops.confusion_accumulate(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.int64), 3)


def test_product_does_not_import_oracle():
    pkg = os.path.join(REPO, "vision_mtl_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_reference_class_surface():
    from vision_mtl_b200.lit_module import MTLModule
    from vision_mtl_b200.losses import SILogLoss
    from vision_mtl_b200.models import (AttentionModuleDecoder, AttentionModuleEncoder, BasicMTLModel,
                                        CrossStitchLayer, CSNet, MTANMiniUnet)

    def params(fn):
        return list(inspect.signature(fn).parameters)[1:]

    assert params(CrossStitchLayer.__init__)[:2] == ["num_tasks", "num_channels"]
    assert params(CSNet.__init__)[:2] == ["models", "channel_wise_stitching"]
    assert params(AttentionModuleEncoder.__init__) == ["shared_1_channels", "out_channels", "shared_2_channels",
                                                       "prev_layer_out_channels", "hidden_channels"]
    assert params(AttentionModuleEncoder.forward) == ["conv1_shared", "conv2_shared", "prev_layer_outs"]
    assert params(AttentionModuleDecoder.__init__) == ["shared_1_channels", "shared_2_channels",
                                                       "prev_layer_out_channels", "out_channels", "hidden_channels"]
    assert params(AttentionModuleDecoder.forward) == ["conv1_shared", "prev_layer_outs", "conv2_shared"]
    assert params(MTANMiniUnet.__init__) == ["in_channels", "map_tasks_to_num_channels",
                                             "task_subnets_hidden_channels", "encoder_first_channel",
                                             "encoder_num_channels"]
    assert params(BasicMTLModel.__init__) == ["segm_classes", "activation", "encoder_name", "encoder_weights",
                                              "decoder_first_channel", "num_decoder_layers", "in_channels"]
    assert params(SILogLoss.forward) == ["pred", "target", "mask", "interpolate", "min_depth"]
    assert params(MTLModule.__init__)[:7] == ["model", "num_classes", "optim_dict", "lr", "device",
                                              "loss_segm_weight", "loss_depth_weight"]
    for m in ("shared_step", "training_step", "validation_step", "test_step", "predict_step", "calc_losses",
              "calc_metrics", "postprocess_raw_out", "update_step_stats", "on_train_epoch_end",
              "on_validation_epoch_end", "on_predict_epoch_end", "transfer_batch_to_device",
              "configure_optimizers"):
        assert callable(getattr(MTLModule, m))
    layer = CrossStitchLayer(3, 8)
    assert layer.weights.shape == (3, 3, 8) and 0 <= float(layer.weights.min()) and float(layer.weights.max()) <= 1
    assert CrossStitchLayer(2).weights.shape == (2, 2)


def test_csnet_plan_and_state_dict_keys():
    from vision_mtl_b200.models import CSNet
    from vision_mtl_b200.utils.model_utils import get_model_with_dense_preds

    models = {"depth": get_model_with_dense_preds(1, None, dict(encoder_weights=None)),
              "segm": get_model_with_dense_preds(19, None, dict(encoder_weights=None))}
    net = CSNet(models, channel_wise_stitching=True)
    assert net.stitch_channels == [16, 24, 40, 80, 112, 160, 1072, 296, 152, 80, 32]
    keys = net.state_dict().keys()
    assert "cross_stitch_layers.0_encoder_model_blocks_1.weights" in keys
    assert "cross_stitch_layers.0_decoder_blocks_4.weights" in keys
    assert any(k.startswith("models.segm.0.encoder.model.conv_stem") for k in keys)
    ops_ = [op for op, _ in net._plan]
    # 11 stitch sites: 6 encoder sites, 4 decoder sites fused with their zero-pad + cat, and the last
    # decoder site fused with its nearest x2 up-sampling
    assert ops_.count("stitch") == 6 and ops_.count("cat_stitch") == 4 and ops_.count("up_stitch") == 1
    assert ops_.count("save_skip") == 4 and ops_.count("cat_skip") == 0 and ops_.count("upsample2") == 0


def test_host_side_rewrites_are_identities_off_gpu():
    """The two host-side rewrites of the drop-in modules must leave CPU / eval behaviour
    untouched: ``conv_without_bias`` keeps the bias unless a CUDA batch-statistics BatchNorm consumes the result, the
    deferred BatchNorm step counters add up exactly once per call, and the optimizer's layout predicate only accepts
    tensors whose elements pair up in memory."""
    from vision_mtl_b200 import ops
    from vision_mtl_b200.optim import _same_layout

    conv, bn = torch.nn.Conv2d(8, 16, 3, padding=1), torch.nn.BatchNorm2d(16)
    x = torch.randn(2, 8, 5, 7)
    y, cb = ops.conv_without_bias(conv, x, bn)  # CPU tensor: stock path
    assert cb is None and torch.equal(y, conv(x))
    bn.eval()
    y, cb = ops.conv_without_bias(conv, x, bn)
    assert cb is None
    bn.train()
    bn2 = torch.nn.BatchNorm2d(16, momentum=None)  # cumulative average: needs its count at call time
    with ops.deferred_batch_counters():
        ops._bump_batch_counter(bn)
        ops._bump_batch_counter(bn)
        ops._bump_batch_counter(bn2)
        assert int(bn.num_batches_tracked) == 0 and int(bn2.num_batches_tracked) == 1
    assert int(bn.num_batches_tracked) == 2 and int(bn2.num_batches_tracked) == 1
    ops._bump_batch_counter(bn)
    assert int(bn.num_batches_tracked) == 3
    a = torch.randn(8, 4, 3, 3)
    b = a.contiguous(memory_format=torch.channels_last)
    assert _same_layout(a, a.clone()) and _same_layout(b, b.clone()) and not _same_layout(a, b)
    c = torch.randn(8, 4, 1, 1)
    assert _same_layout(c, c.contiguous(memory_format=torch.channels_last))  # size-1 dimensions do not matter
    assert not _same_layout(a, a[:, ::2])


def test_conv3x_split_is_exact():
    """The hi / lo split behind the opt-in 3xTF32 convolutions: hi has 13 zero low mantissa bits, hi + lo == x exactly."""
    from vision_mtl_b200.conv3x import tf32_split

    x = torch.randn(10000) * torch.logspace(-20, 20, 10000)
    hi, lo = tf32_split(x)
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert torch.equal(hi + lo, x)
    assert float((lo.abs() / x.abs().clamp_min(1e-38)).max()) <= 2.0 ** -11
