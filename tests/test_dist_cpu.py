"""World-size-2 gloo tests (CPU) of the data-parallel host logic in ``vision_mtl_b200/dist.py``:
batch sharding, the single packed metric all-reduce and DDP gradient averaging."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import metrics_np as MN


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from vision_mtl_b200 import dist as vdist

    r, lr, w = vdist.init_distributed("gloo")
    assert (r, lr, w) == (rank, rank, world) and vdist.is_distributed()
    C, B = 7, 8
    g = torch.Generator().manual_seed(3)
    full = {"img": torch.randn(B, 3, 4, 4, generator=g), "mask": torch.randint(0, C, (B, 4, 4), generator=g),
            "depth": torch.rand(B, 4, 4, 1, generator=g)}
    pred = torch.randint(0, C, (B, 4, 4), generator=g)
    shard = vdist.shard_batch(full, rank, world)
    assert shard["img"].shape[0] == B // world
    lo, hi = rank * B // world, (rank + 1) * B // world
    assert torch.equal(shard["mask"], full["mask"][lo:hi])

    # ---- packed metric all-reduce: global confusion == confusion of the concatenated batch -----
    conf = torch.from_numpy(MN.confusion_matrix(pred[lo:hi].numpy(), shard["mask"].numpy(), C))
    loss = torch.tensor(1.0 + rank)
    dsum = torch.tensor([float(shard["depth"].numel()), 2.0 + rank, 5.0, 0.5 * (rank + 1)], dtype=torch.float64)
    stats = vdist.allreduce_step_stats(conf, loss, dsum)
    ref = MN.confusion_matrix(pred.numpy(), full["mask"].numpy(), C)
    assert np.array_equal(stats["confusion"].numpy(), ref), "global confusion matrix must be bit-exact"
    assert abs(float(stats["loss"]) - 1.5) < 1e-12 and float(stats["replicas"]) == world
    assert abs(float(stats["mae"]) - (2.0 + 3.0) / full["depth"].numel()) < 1e-12
    assert abs(float(stats["abs_rel"]) - 1.5 / 10.0) < 1e-12

    # ---- DDP gradient averaging through wrap_data_parallel -------------------------------------
    class Holder:  # stands in for MTLModule: wrap_data_parallel only touches `.model`
        pass

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.Tanh(), torch.nn.Conv2d(4, C, 1))
    ref_net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.Tanh(), torch.nn.Conv2d(4, C, 1))
    ref_net.load_state_dict(net.state_dict())
    holder = Holder()
    holder.model = net
    vdist.wrap_data_parallel(holder)
    assert type(holder.model).__name__ == "DistributedDataParallel"
    torch.nn.functional.cross_entropy(holder.model(shard["img"]), shard["mask"]).backward()
    torch.nn.functional.cross_entropy(ref_net(full["img"]), full["mask"]).backward()
    for p, q in zip(net.parameters(), ref_net.parameters()):
        assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-7), "DDP grads != single-process grads"
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_pack_unpack_roundtrip_single_process():
    from vision_mtl_b200 import dist as vdist

    C = 19
    conf = torch.randint(0, 2**40, (C, C), dtype=torch.int64)  # far beyond fp32 integers, exact in fp64
    buf = vdist.pack_step_stats(conf, torch.tensor(3.25))
    assert buf.dtype == torch.float64 and buf.numel() == C * C + vdist.STAT_EXTRA
    out = vdist.unpack_step_stats(buf, C)
    assert torch.equal(out["confusion"], conf) and float(out["loss"]) == 3.25
    out = vdist.allreduce_step_stats(conf, torch.tensor(3.25))  # no process group: identity
    assert torch.equal(out["confusion"], conf)
    assert vdist.env_world() == (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
                                 int(os.environ.get("WORLD_SIZE", 1)))
    with pytest.raises(ValueError):
        vdist.shard_batch({"img": torch.zeros(3, 1)}, 0, 2)
