"""Batch-sharded data parallelism: one process per GPU, replicated weights, local BatchNorm statistics (what the
reference's nn.DataParallel does, cm/train_trgb_segnet_conf.py:234), and ONE exchange per step: the gradient average of
whichever parameter set learns in the current phase -- the seg net (217 MB FP32) in train_seg, the critics (86 MB) in
train_critic (conf_segnet.setPhase flips requires_grad).  Inference and iou_eval shard over images with no collective.

How the exchange is organised (GradientReducer):

  * a persistent flat FP32 ARENA holds every parameter's gradient (reverse registration order ~ the order backward produces
    them).  The backward kernels write straight into it (engine.Grads.param_sink) and `p.grad` is a view of its slot, so
    there is no flatten / unflatten copy, gradient addresses never change (the fused optimizer's tables and a captured CUDA
    graph stay valid) and a bucket is one contiguous range;
  * the arena is cut into buckets of ~bucket_mb.  The first step of a phase runs un-overlapped and LEARNS the order in
    which the tape produces gradients (the same network runs on the day and on the night batch, so most tensors receive two
    contributions and only the second one is final).  From then on a bucket's all-reduce is enqueued on a communication
    stream the moment its last gradient has been written -- ordered by a CUDA event, no host synchronisation -- and overlaps
    the rest of backward; `finish()` makes the compute stream wait for the communication stream.  One plan per live
    parameter set, i.e. per phase;
  * on CUDA the collective is `ncclAllReduce(avg)` issued directly on that stream through the NCCL library torch already
    loaded (own communicator; the unique id travels over torch.distributed): stream-ordered, no work handles, no watchdog,
    therefore capturable -- graphs.GraphedStep replays the whole step, collectives included, for N > 1.  On CPU tensors (the
    gloo tests) torch.distributed.all_reduce is the transport.

A gradient that arrives after its bucket has been reduced (the step's control flow changed under the same live set)
raises instead of training on a stale average; `relearn()` forgets the plans.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist


def broadcast_parameters(module: torch.nn.Module, src: int = 0):
    """Rank `src`'s parameters and buffers (BN running statistics) to every rank, once at start-up."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Even split of the batch dimension; day and night batches are split identically so pairs stay together."""
    n = t.shape[0]
    assert n % world == 0, "global batch must divide by the number of ranks"
    per = n // world
    return t[rank * per:(rank + 1) * per]


# ---------------------------------------------------------------------------------------------------- NCCL, stream-ordered
class _NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_ubyte * 128)]      # c_ubyte, not c_char: a c_char array field reads back truncated at the first NUL


_NCCL_FLOAT32, _NCCL_SUM, _NCCL_AVG = 7, 0, 4


class NcclComm:
    """Own NCCL communicator over the ranks of the default process group, driven through ctypes on the libnccl torch has
    already mapped.  Calls are plain stream-ordered enqueues (`ncclAllReduce(..., stream)`)."""

    def __init__(self, device: torch.device):
        assert dist.is_initialized()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.lib = self._load()
        uid = _NcclUniqueId()
        if self.rank == 0:
            self._check(self.lib.ncclGetUniqueId(C.byref(uid)))
        raw = C.string_at(C.addressof(uid), 128) if self.rank == 0 else bytes(128)
        t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone().to(device)
        dist.broadcast(t, 0)
        raw = t.cpu().numpy().tobytes()
        assert len(raw) == 128
        C.memmove(C.addressof(uid), raw, 128)
        self.comm = C.c_void_p()
        with torch.cuda.device(device):
            self._check(self.lib.ncclCommInitRank(C.byref(self.comm), self.world, uid, self.rank))
        self.device = device

    @staticmethod
    def _load():
        last = None
        for name in ("libnccl.so.2", "libnccl.so"):
            try:
                lib = C.CDLL(name)
                break
            except OSError as e:
                last = e
        else:
            raise RuntimeError(f"NCCL library not found: {last}")
        lib.ncclGetErrorString.restype = C.c_char_p
        lib.ncclGetErrorString.argtypes = [C.c_int]
        lib.ncclGetUniqueId.argtypes = [C.POINTER(_NcclUniqueId)]
        lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _NcclUniqueId, C.c_int]
        lib.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.ncclCommDestroy.argtypes = [C.c_void_p]
        return lib

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("NCCL: " + self.lib.ncclGetErrorString(rc).decode())

    def all_reduce_avg_(self, t: torch.Tensor, stream: torch.cuda.Stream):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
        self._check(self.lib.ncclAllReduce(t.data_ptr(), t.data_ptr(), t.numel(), _NCCL_FLOAT32, _NCCL_AVG, self.comm, C.c_void_p(stream.cuda_stream)))

    def destroy(self):
        if self.comm:
            self.lib.ncclCommDestroy(self.comm)
            self.comm = C.c_void_p()


# ---------------------------------------------------------------------------------------------------- the reducer
class _Plan:
    """What one phase's backward looks like: the order in which parameter gradients are produced, the buckets of the live
    set and, for each bucket, the index of the production event after which it is complete."""
    __slots__ = ("seq", "buckets", "fire_at")

    def __init__(self, seq: List[int], buckets: List[tuple], fire_at: Dict[int, List[int]]):
        self.seq, self.buckets, self.fire_at = seq, buckets, fire_at


class GradientReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 32.0, overlap: bool = True):
        self.params: List[torch.nn.Parameter] = list(params)
        assert self.params, "GradientReducer: no parameters"
        dev = self.params[0].device
        self.bucket_bytes = max(int(bucket_mb * (1 << 20)), 4)
        self.overlap = overlap
        self.enabled = True                    # False: gradients stay local (bench: time the step without the exchange)
        # arena: reverse registration order, every slot 16-byte aligned (vector stores of the wgrad epilogues)
        self.slot: Dict[torch.nn.Parameter, tuple] = {}
        self.order = list(reversed(self.params))
        off = 0
        for p in self.order:
            assert p.dtype == torch.float32, "gradients are FP32 (the masters)"
            self.slot[p] = (off, p.numel())
            off += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self._index = {id(p): i for i, p in enumerate(self.order)}
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.nccl: Optional[NcclComm] = None
        self.comm_stream = None
        if dev.type == "cuda":
            self.comm_stream = torch.cuda.Stream(device=dev)
            if self.world > 1:
                self.nccl = NcclComm(dev)
        self.plans: Dict[tuple, _Plan] = {}
        self._reset_cycle()
        self.last_buckets = self.last_bytes = 0
        self.copied = 0                        # gradients that had to be copied into the arena (produced outside it)
        from . import engine as E
        E.grad_arena = self

    # ---- arena ---------------------------------------------------------------------------------------------------
    def view(self, p: torch.nn.Parameter) -> torch.Tensor:
        off, n = self.slot[p]
        return self.flat[off:off + n].view(p.shape)

    def owns(self, p) -> bool:
        return p in self.slot

    def zero_grad(self):
        """Start of a step: instead of optimizer.zero_grad() / `p.grad = None`."""
        for p in self.params:
            p.grad = None
        self._reset_cycle()

    def _reset_cycle(self):
        self.written = set()
        self.seq: List[int] = []
        self.fired = set()
        self.plan: Optional[_Plan] = None
        self._key = None
        self._in_backward = False
        self._cycle_buckets = self._cycle_bytes = 0

    # ---- called by the autograd glue (autograd._NetFunction) and engine.Grads -------------------------------------------
    def on_forward(self):
        if self._in_backward:              # a forward after a backward: a new step has begun without zero_grad()
            self._reset_cycle()

    def on_backward(self):
        if not self._in_backward:
            self._in_backward = True
            self._key = tuple(i for i, p in enumerate(self.order) if p.requires_grad)
            self.plan = self.plans.get(self._key) if (self.overlap and self.enabled and self.world > 1) else None

    def sink(self, p: torch.nn.Parameter):
        """-> (view of p's slot, accumulate): where the backward kernels write p's gradient.  The first contribution of a step
        overwrites the slot; later ones (second use of the module in this step, or a micro-batch accumulated into an
        existing .grad) add to it."""
        acc = id(p) in self.written or (p.grad is not None and p.grad.data_ptr() == self.flat.data_ptr() + 4 * self.slot[p][0])
        if self.plan is not None and self._bucket_of(p) in self.fired:
            raise RuntimeError("GradientReducer: a gradient arrived after its bucket was reduced -- the step's control flow differs "
                               "from the learned plan; call reducer.relearn()")
        return self.view(p), acc

    def produced(self, p: torch.nn.Parameter):
        """The kernel writing (a contribution to) p's gradient has been enqueued on the current stream."""
        self.written.add(id(p))
        i = len(self.seq)
        self.seq.append(self._index[id(p)])
        plan = self.plan
        if plan is None:
            return
        if i >= len(plan.seq) or plan.seq[i] != self.seq[i]:
            if self.fired:
                raise RuntimeError("GradientReducer: backward deviates from the learned plan after buckets were reduced; call "
                                   "reducer.relearn()")
            self.plan = None                   # fall back to the un-overlapped reduction in finish(), relearn there
            self.plans.pop(self._key, None)
            return
        for b in plan.fire_at.get(i, ()):
            self._fire(plan.buckets[b], b)

    def _bucket_of(self, p):
        if self.plan is None:
            return None
        off = self.slot[p][0]
        for b, (lo, hi) in enumerate(self.plan.buckets):
            if lo <= off < hi:
                return b
        return None

    # ---- the exchange ----------------------------------------------------------------------------------------------
    def _all_reduce_range(self, lo: int, hi: int, stream):
        buf = self.flat[lo:hi]
        if self.nccl is not None:
            self.nccl.all_reduce_avg_(buf, stream)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            buf.div_(self.world)

    def _fire(self, rng, b):
        self.fired.add(b)
        self._cycle_buckets += 1
        self._cycle_bytes += 4 * (rng[1] - rng[0])
        if self.comm_stream is None:
            self._all_reduce_range(rng[0], rng[1], None)
            return
        cur = torch.cuda.current_stream(self.flat.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.comm_stream.wait_event(ev)
        self._all_reduce_range(rng[0], rng[1], self.comm_stream)

    def _live_buckets(self, key):
        """Maximal runs of consecutive live slots, cut at ~bucket_bytes -> [(lo, hi)] element ranges of the arena."""
        live = set(key)
        buckets, lo, hi, prev = [], None, None, None
        for i, p in enumerate(self.order):
            if i not in live:
                continue
            off, n = self.slot[p]
            end = off + (n + 3) // 4 * 4
            if lo is not None and (prev != i - 1 or 4 * (hi - lo) >= self.bucket_bytes):
                buckets.append((lo, hi))
                lo = None
            if lo is None:
                lo = off
            hi, prev = end, i
        if lo is not None:
            buckets.append((lo, hi))
        return buckets

    def _learn(self, key):
        buckets = self._live_buckets(key)
        last = {}
        for i, idx in enumerate(self.seq):
            last[idx] = i
        fire_at: Dict[int, List[int]] = {}
        for b, (lo, hi) in enumerate(buckets):
            members = [i for i in key if lo <= self.slot[self.order[i]][0] < hi and i in last]
            if not members:
                continue                               # nothing of this bucket was produced (unused parameters): reduced in finish()
            fire_at.setdefault(max(last[i] for i in members), []).append(b)
        self.plans[key] = _Plan(list(self.seq), buckets, fire_at)

    def finish(self):
        """After loss.backward(): adopt gradients produced outside the arena, reduce what has not been reduced yet, and
        make the compute stream wait for the exchange.  Gradients are AVERAGES over the ranks afterwards."""
        key = self._key if self._key is not None else tuple(i for i, p in enumerate(self.order) if p.requires_grad)
        for p in self.params:                          # e.g. produced by plain torch autograd: move into the arena
            g = p.grad
            if g is not None and g.data_ptr() != self.flat.data_ptr() + 4 * self.slot[p][0]:
                v = self.view(p)
                v.copy_(g)
                p.grad = v
                self.copied += 1
        if self.world > 1 and self.enabled:
            plan = self.plan
            buckets = plan.buckets if plan is not None else self._live_buckets(key)
            for b, rng in enumerate(buckets):
                if b not in self.fired:
                    self._fire(rng, b)
            if self.comm_stream is not None:
                torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)
            if plan is None and self.overlap and self.seq and key not in self.plans:
                self._learn(key)
        self.last_buckets, self.last_bytes = self._cycle_buckets, self._cycle_bytes
        self._reset_cycle()
        self._in_backward = True                       # the next forward starts a new cycle

    # kept for callers of the round-1 API: reduce() == finish()
    reduce = finish

    def relearn(self):
        self.plans.clear()

    def close(self):
        from . import engine as E
        if E.grad_arena is self:
            E.grad_arena = None
        if self.nccl is not None:
            self.nccl.destroy()
            self.nccl = None
