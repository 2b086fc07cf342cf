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
        reducer = parallel.GradientReducer(model.parameters(), bucket_mb=0.0005)     # tiny buckets: several per step
        g = torch.Generator().manual_seed(7)
        x_all = torch.randn(4, 3, 12, 12, generator=g)
        x = parallel.shard_batch(x_all, rank, world)
        results = {}
        for phase in ("train_critic", "train_seg"):
            for p in seg.parameters():
                p.requires_grad = phase == "train_seg"
            for p in critic.parameters():
                p.requires_grad = phase == "train_critic"
            for p in model.parameters():
                p.grad = None
            loss = critic(seg(x)).pow(2).mean()
            loss.backward()
            local = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
            reducer.reduce()
            assert reducer.last_buckets > 1
            live = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
            assert set(live) == set(local)
            assert all(k.startswith("seg" if phase == "train_seg" else "critic") for k in live)
            # averaged gradient == mean of the per-rank local gradients
            for k in live:
                both = [None] * world
                dist.all_gather_object(both, local[k])
                want = sum(both) / world
                assert torch.allclose(live[k], want, atol=1e-6), k
            results[phase] = sorted(live)
        # a training step with the all-reduce inside must not be captured as a CUDA graph under torch.distributed
        from heatnet_pub_b200 import graphs
        try:
            graphs.GraphedStep(lambda t: t, [torch.zeros(1)], module=model)
            refused = False
        except RuntimeError as e:
            refused = "collectives_inside=False" in str(e)
        assert refused, "GraphedStep must refuse a capture with collectives inside when world_size > 1"
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
