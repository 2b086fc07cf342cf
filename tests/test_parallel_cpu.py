"""World-size-2 gloo tests (CPU) of the data-parallel host logic: parameter broadcast, phase-aware bucketed
gradient averaging.  No GPU kernel is involved."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _emulate_backward(reducer, contributions):
    """What autograd._NetFunction.backward + engine.Grads do on the GPU, with torch CPU ops: every (param, tensor) contribution
    is written into the parameter's arena slot (first one overwrites, later ones accumulate), then .grad is bound to the slot."""
    reducer.on_backward()
    touched = []
    for p, g in contributions:
        dst, acc = reducer.sink(p)
        dst.add_(g) if acc else dst.copy_(g)
        reducer.produced(p)
        touched.append((p, dst))
    for p, dst in touched:
        if p.grad is None or p.grad.data_ptr() != dst.data_ptr():
            p.grad = dst


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from heatnet_pub_b200 import parallel
        torch.manual_seed(100 + rank)                       # different init per rank on purpose
        seg = nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.Conv2d(8, 4, 1))
        critic = nn.Sequential(nn.Conv2d(4, 8, 4, 2, 1), nn.Conv2d(8, 1, 4, 2, 1))
        model = nn.ModuleDict({"seg": seg, "critic": critic})
        parallel.broadcast_parameters(model, 0)
        ref = [p.detach().clone() for p in model.parameters()]
        gathered = [None] * world
        dist.all_gather_object(gathered, [p.sum().item() for p in ref])
        assert gathered[0] == gathered[1], "parameters differ after broadcast"
        reducer = parallel.GradientReducer(model.parameters(), bucket_mb=0.0001)     # tiny buckets: several per step
        g = torch.Generator().manual_seed(7)
        x_all = torch.randn(4, 3, 12, 12, generator=g)
        x = parallel.shard_batch(x_all, rank, world)
        results = {}
        names = {id(p): k for k, p in model.named_parameters()}

        def check_average(local):
            live = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
            assert set(live) == set(local)
            for k in live:
                both = [None] * world
                dist.all_gather_object(both, local[k])
                assert torch.allclose(live[k], sum(both) / world, atol=1e-6), k
            return live

        # ---- (1) gradients produced by plain torch autograd (outside the arena): adopted, reduced after backward
        for phase in ("train_critic", "train_seg"):
            for p in seg.parameters():
                p.requires_grad = phase == "train_seg"
            for p in critic.parameters():
                p.requires_grad = phase == "train_critic"
            reducer.zero_grad()
            loss = critic(seg(x)).pow(2).mean()
            loss.backward()
            local = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
            reducer.finish()
            assert reducer.last_buckets > 1 and reducer.copied > 0, (reducer.last_buckets, reducer.copied, reducer.world)
            live = check_average(local)
            assert all(k.startswith("seg" if phase == "train_seg" else "critic") for k in live)
            assert all(p.grad.data_ptr() == reducer.view(p).data_ptr() for p in model.parameters() if p.grad is not None)
            results[phase] = sorted(live)

        # ---- (2) the arena protocol the GPU backward follows: two contributions per tensor (day + night pass), production order
        # = reverse registration order twice.  Step 1 learns the plan and reduces at finish(); steps 2.. fire every bucket the
        # moment its last contribution has been produced, i.e. DURING the backward.
        live_params = [p for p in reversed(list(seg.parameters()))]
        for step in range(3):
            gen = torch.Generator().manual_seed(1000 * rank + step)
            contribs = [(p, torch.randn(p.shape, generator=gen)) for p in live_params] + [(p, torch.randn(p.shape, generator=gen)) for p in live_params]
            local = {}
            for p, t in contribs:
                local[names[id(p)]] = local.get(names[id(p)], 0) + t
            reducer.zero_grad()
            reducer.on_forward()
            fired_during = []
            orig = reducer._fire
            reducer._fire = lambda rng, b, _o=orig: (fired_during.append((b, len(reducer.seq))), _o(rng, b))[1]
            _emulate_backward(reducer, contribs)
            n_during = len(fired_during)
            reducer.finish()
            reducer._fire = orig
            check_average(local)
            key = tuple(i for i, p in enumerate(reducer.order) if p.requires_grad)
            if step == 0:
                assert n_during == 0 and key in reducer.plans            # learning step: nothing overlapped
            else:
                plan = reducer.plans[key]
                assert n_during == len(plan.buckets) > 1                  # every bucket went out during the backward ...
                assert all(at > len(live_params) for _, at in fired_during)   # ... and only after the SECOND contribution
        results["plan_buckets"] = len(reducer.plans[key].buckets)

        # ---- (3) a gradient arriving after its bucket was reduced must raise, not train on a stale average
        reducer.zero_grad()
        reducer.on_forward()
        try:
            _emulate_backward(reducer, contribs + [(live_params[0], torch.ones(live_params[0].shape))])
            raised = False
        except RuntimeError as e:
            raised = "relearn" in str(e)
        assert raised
        # every rank is in the same state (the first buckets were reduced, then the error): forget the plan and carry on
        reducer.relearn()
        reducer.zero_grad()
        reducer.on_forward()
        _emulate_backward(reducer, contribs)
        reducer.finish()
        check_average(local)
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gradient_reducer_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    got = dict(q.get(timeout=10) for _ in range(world))
    assert got[0] == got[1]
    assert len(got[0]["train_seg"]) == 6 and len(got[0]["train_critic"]) == 4
    assert got[0]["plan_buckets"] > 1
