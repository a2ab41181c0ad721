"""Multi-tensor Adam (vmtl_adam_step) against torch.optim.Adam -- the optimizer the reference builds at
training_lit.py:51 -- on the same parameters and gradients: parameters and both moments after several steps,
weight decay, a device learning rate changed between steps, parameters without gradients, channels_last
parameters, state_dict round trip into torch.optim.Adam and back, and CUDA-graph capture."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-6  # relative to the largest element; both sides are fp32 with different contraction of the same formula


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _params(dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(7,), (33, 5), (64, 32, 3, 3), (19, 32, 1, 1), (1,), (3, 70001), (128,), (2, 3, 5, 7)]
    ps = [torch.randn(*s, generator=g).to(dev).requires_grad_(True) for s in shapes]
    ps[2] = ps[2].detach().contiguous(memory_format=torch.channels_last).requires_grad_(True)  # a channels_last weight
    return ps


def _grads(ps, step, seed=0):
    g = torch.Generator().manual_seed(1000 * seed + step)
    out = []
    for p in ps:
        gr = torch.randn(p.shape, generator=g).to(p.device) * (0.1 + step)
        if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
            gr = gr.contiguous(memory_format=torch.channels_last)
        out.append(gr)
    return out


@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
@pytest.mark.parametrize("tensor_lr", [False, True])
def test_adam_matches_torch(weight_decay, tensor_lr):
    from vision_mtl_b200.optim import Adam

    dev = torch.device("cuda:0")
    mine, ref = _params(dev), _params(dev)
    lr = torch.tensor(3e-3, device=dev) if tensor_lr else 3e-3
    opt = Adam(mine, lr=lr, weight_decay=weight_decay)
    opt_ref = torch.optim.Adam(ref, lr=3e-3, weight_decay=weight_decay)
    for step in range(6):
        for ps in (mine, ref):
            for p, g in zip(ps, _grads(ps, step)):
                p.grad = g.clone()
        if step == 5:  # a parameter without a gradient is skipped (its moments and value stay)
            mine[4].grad = None
            ref[4].grad = None
        if step == 4:  # what ReduceLROnPlateau does
            if tensor_lr:
                opt.param_groups[0]["lr"].fill_(1e-3)
            else:
                opt.param_groups[0]["lr"] = 1e-3
            opt_ref.param_groups[0]["lr"] = 1e-3
        opt.step()
        opt_ref.step()
    torch.cuda.synchronize()
    assert float(opt.state[mine[0]]["step"]) == 6.0
    for a, b in zip(mine, ref):
        assert _rel(a.detach(), b.detach()) <= TOL
        assert _rel(opt.state[a]["exp_avg"], opt_ref.state[b]["exp_avg"]) <= TOL
        assert _rel(opt.state[a]["exp_avg_sq"], opt_ref.state[b]["exp_avg_sq"]) <= TOL


def test_adam_state_dict_round_trip():
    from vision_mtl_b200.optim import Adam

    dev = torch.device("cuda:0")
    mine, ref = _params(dev, 3), _params(dev, 3)
    opt, opt_ref = Adam(mine, lr=1e-3), torch.optim.Adam(ref, lr=1e-3)
    for step in range(2):
        for ps, o in ((mine, opt), (ref, opt_ref)):
            for p, g in zip(ps, _grads(ps, step, 3)):
                p.grad = g.clone()
            o.step()
    # my state -> torch.optim.Adam -> continues identically
    sd = copy.deepcopy(opt.state_dict())  # torch's load_state_dict keeps same-device tensors by reference
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    fresh = _params(dev, 3)
    with torch.no_grad():
        for a, b in zip(fresh, mine):
            a.copy_(b)
    opt_t = torch.optim.Adam(fresh, lr=1e-3, capturable=True)  # device-side step counters, like the loaded state
    opt_t.load_state_dict(sd)
    # torch state -> mine
    mine2 = _params(dev, 3)
    with torch.no_grad():
        for a, b in zip(mine2, ref):
            a.copy_(b)
    opt2 = Adam(mine2, lr=1e-3)
    opt2.load_state_dict(opt_ref.state_dict())
    for ps, o in ((fresh, opt_t), (mine2, opt2), (ref, opt_ref), (mine, opt)):
        for p, g in zip(ps, _grads(ps, 2, 3)):
            p.grad = g.clone()
        o.step()
    for a, b, c, d in zip(fresh, mine2, ref, mine):
        assert _rel(a.detach(), c.detach()) <= TOL
        assert _rel(b.detach(), c.detach()) <= TOL
        assert _rel(d.detach(), c.detach()) <= TOL


def test_adam_many_tensors_and_graph_capture():
    """More tensors than one launch's table holds (several launches, one step bump), captured into a CUDA graph:
    three replays equal three eager torch steps."""
    from vision_mtl_b200 import _lib
    from vision_mtl_b200.optim import Adam

    dev = torch.device("cuda:0")
    n = _lib.load().vmtl_adam_max_tensors_per_launch() + 37
    g = torch.Generator().manual_seed(5)
    base = [torch.randn(1 + (i * 7) % 50, generator=g) for i in range(n)]
    mine = [b.clone().to(dev).requires_grad_(True) for b in base]
    ref = [b.clone().to(dev).requires_grad_(True) for b in base]
    grads = [torch.randn_like(p) for p in mine]
    for p, r, gr in zip(mine, ref, grads):
        p.grad = gr.clone()
        r.grad = gr.clone()
    opt = Adam(mine, lr=torch.tensor(1e-2, device=dev))
    opt_ref = torch.optim.Adam(ref, lr=1e-2)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt.step()  # warm-up (also builds the state)
    torch.cuda.current_stream().wait_stream(side)
    opt_ref.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    opt_ref.step()  # the captured (and, for torch.cuda.graph, not executed) step ... replay it once below
    graph.replay()
    for _ in range(2):
        graph.replay()
        opt_ref.step()
    torch.cuda.synchronize()
    assert float(opt.state[mine[0]]["step"]) == 4.0
    for a, b in zip(mine, ref):
        assert _rel(a.detach(), b.detach()) <= TOL


def test_adam_rejects_cpu_parameters():
    from vision_mtl_b200 import _lib
    from vision_mtl_b200.optim import Adam

    p = torch.zeros(4, requires_grad=True)
    p.grad = torch.ones(4)
    with pytest.raises(_lib.VmtlError):
        Adam([p], lr=1e-3).step()
