"""Batch-sharded data parallelism: one process per GPU, replicated weights, local BatchNorm statistics (what the
reference's nn.DataParallel does, cm/train_trgb_segnet_conf.py:234), and ONE exchange per step: a bucketed
all-reduce (sum, then / world) of whichever parameter set holds gradients in the current phase -- the seg net
(217 MB FP32) in train_seg, the critics (86 MB) in train_critic (conf_segnet.setPhase flips requires_grad, so the
live set is simply "parameters whose .grad is not None").  torch.distributed (NCCL over NVLink on GPUs, gloo in the
CPU tests) is the transport; inference and iou_eval shard over images with no collective at all.
"""
from typing import Iterable, List

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


def broadcast_parameters(module: torch.nn.Module, src: int = 0):
    """Rank `src`'s parameters and buffers (BN running statistics) to every rank, once at start-up."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


class GradientReducer:
    """`reducer.reduce()` after `loss.backward()`: averages the live gradients over the ranks in buckets of
    ~`bucket_mb` MB, each all-reduce launched asynchronously so later buckets are flattened while earlier ones are
    on the wire."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 32.0):
        self.params: List[torch.nn.Parameter] = list(params)
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.last_buckets = 0
        self.last_bytes = 0

    def _buckets(self, grads):
        bucket, size = [], 0
        for g in grads:
            nbytes = g.numel() * g.element_size()
            if bucket and (size + nbytes > self.bucket_bytes or g.dtype != bucket[0].dtype):
                yield bucket
                bucket, size = [], 0
            bucket.append(g)
            size += nbytes
        if bucket:
            yield bucket

    def reduce(self):
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size()
        if world == 1:
            return
        # reverse registration order = roughly the order gradients were produced in backward
        grads = [p.grad for p in reversed(self.params) if p.grad is not None]
        pending = []
        self.last_buckets, self.last_bytes = 0, 0
        for bucket in self._buckets(grads):
            flat = _flatten_dense_tensors(bucket)
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
            pending.append((work, flat, bucket))
            self.last_buckets += 1
            self.last_bytes += flat.numel() * flat.element_size()
        for work, flat, bucket in pending:
            work.wait()
            flat.div_(world)
            for g, synced in zip(bucket, _unflatten_dense_tensors(flat, bucket)):
                g.copy_(synced)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Even split of the batch dimension; day and night batches are split identically so pairs stay together."""
    n = t.shape[0]
    assert n % world == 0, "global batch must divide by the number of ranks"
    per = n // world
    return t[rank * per:(rank + 1) * per]
