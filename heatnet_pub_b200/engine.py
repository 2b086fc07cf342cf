"""Host-side op layer over the C ABI: NHWC activation views, parameter caches and one Python function per
kernel family.  torch is used for device memory (caching allocator), the current stream and parameter
storage only; every FLOP on the path runs in libheatnet_b200.so.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ACT_LEAKY, ACT_NONE, ACT_RELU, HN_BF16, HN_F32, HnConv, HnEpilogue, HnTensor

BN_EPS_DEFAULT = 1e-5
_TORCH_DTYPE = {HN_F32: torch.float32, HN_BF16: torch.bfloat16}
_HN_DTYPE = {torch.float32: HN_F32, torch.bfloat16: HN_BF16}

DEFAULT_PRECISION = os.environ.get("HEATNET_B200_PRECISION", "bf16")
# Storage type of the pre-normalisation conv output in train-mode BatchNorm2d on the BF16 path.  BF16 (default) is what
# torch.autocast stores too: 10 bytes/element less HBM traffic over the forward + backward of every BN layer, and the batch
# statistics are those of the stored (rounded) tensor, accumulated by the conv epilogue.  "1" keeps the tensor in FP32:
# (x - mean) * invstd is then free of the BF16 rounding of x (amplified by |mean|/std), closer to the FP32 reference.
BN_TRAIN_RAW_FP32 = os.environ.get("HEATNET_B200_BN_RAW_FP32", "0") != "0"
# Fused 2x-upsample + 3x3 conv (hn_upconv3x3_fwd) is used for inputs with at least this many channels.  Measured on B200
# (batch 16, 650x1920): with the exact-2x bilinear kernel at 3.1 TB/s the separate path wins everywhere -- up_1 (Cin 1024):
# fused 5.08 ms vs 0.82 + 3.60 ms unfused -- because the halo producers are CUDA-core work competing with the epilogue for
# issue slots.  The kernel stays (tests, HN_UPCONV_MIN_CIN=512 to enable): it is the right structure once the patch is
# produced by tensor cores (interpolation as a small GEMM), which is future work.
UPCONV_MIN_CIN = int(os.environ.get("HN_UPCONV_MIN_CIN", "1000000"))
STEM_FAST = os.environ.get("HN_NO_STEM_FAST") is None

# number of libheatnet_b200 kernels enqueued by this process (bench.py reports it as gpu_launches)
launch_count = 0
# bench.py sets this to a list to collect (description, start_event, end_event) around every conv launch
conv_timer = None
# ... and this one around every other op (bilinear, pools, layout conversion)
op_timer = None


class _timed:
    """`with _timed("bilinear"):` records CUDA events around the enclosed launches when op_timer is a list."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if op_timer is not None:
            self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if op_timer is not None:
            self.b.record()
            op_timer.append((self.name, self.a, self.b))
        return False


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def precision_dtype(precision: str) -> torch.dtype:
    if precision == "bf16":
        return torch.bfloat16
    if precision == "fp32":
        return torch.float32
    raise ValueError("precision must be 'bf16' or 'fp32'")


def refuse_autograd(module, *inputs):
    """The forward-only entry points build no autograd graph: refuse loudly instead of dropping gradients."""
    if not torch.is_grad_enabled():
        return
    if any(p.requires_grad for p in module.parameters()) or any(torch.is_tensor(t) and t.requires_grad for t in inputs):
        raise NotImplementedError(
            "heatnet_pub_b200: backward kernels are not wired into this entry point yet -- call under "
            "torch.no_grad() (inference / validation), or freeze the parameters")


class Act:
    """NHWC activation view: `buf` is a dense [N, H, W, LD] tensor, the view covers channels
    [coff, coff + c).  Channel slices of one buffer are how torch.cat(dim=1) becomes zero-copy."""
    __slots__ = ("buf", "c", "coff")

    def __init__(self, buf: torch.Tensor, c: Optional[int] = None, coff: int = 0):
        assert buf.dim() == 4 and buf.is_contiguous() and buf.is_cuda
        self.buf, self.coff = buf, coff
        self.c = buf.shape[3] - coff if c is None else c

    n = property(lambda s: s.buf.shape[0])
    h = property(lambda s: s.buf.shape[1])
    w = property(lambda s: s.buf.shape[2])
    ld = property(lambda s: s.buf.shape[3])
    dtype = property(lambda s: s.buf.dtype)

    def ptr(self) -> int:
        return self.buf.data_ptr() + self.coff * self.buf.element_size()

    def hn(self) -> HnTensor:
        return HnTensor(self.ptr(), _HN_DTYPE[self.buf.dtype], self.n, self.h, self.w, self.c, self.ld)

    def slice(self, coff: int, c: int) -> "Act":
        assert coff + c <= self.c
        return Act(self.buf, c, self.coff + coff)

    def nchw(self) -> torch.Tensor:
        """Zero-copy NCHW-shaped (channels-last strided) torch view, the layout returned to callers."""
        return self.buf[..., self.coff:self.coff + self.c].permute(0, 3, 1, 2)


def new_act(n, h, w, c, dtype, device, ld=None) -> Act:
    return Act(torch.empty((n, h, w, ld or c), dtype=dtype, device=device), c)


def act_from_view(t: torch.Tensor) -> Optional[Act]:
    """Recover the NHWC Act behind a tensor produced by Act.nchw() (no copy), else None."""
    if t.dim() != 4 or not t.is_cuda or t.dtype not in _HN_DTYPE:
        return None
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    if sc != 1 or sw < c or sh != w * sw or sn != h * w * sw:
        return None
    base = t._base if t._base is not None else None
    if base is None or base.dim() != 4 or not base.is_contiguous() or base.shape[:3] != (n, h, w) or base.shape[3] != sw:
        return None
    coff = (t.data_ptr() - base.data_ptr()) // t.element_size()
    if coff < 0 or coff + c > sw:
        return None
    return Act(base, c, coff)


# -------------------------------------------------------------------------------------------------- autograd tape
class Grads:
    """Gradient storage of one backward pass: one NHWC gradient buffer per activation buffer (same shape, compute
    dtype), with the channel ranges that already hold contributions.  Channel-slice views of one buffer (the
    zero-copy concats) therefore accumulate into slices of one gradient buffer."""

    def __init__(self):
        self.bufs = {}          # activation buf data_ptr -> [grad tensor, [(c0, c1), ...]]
        self.params = {}        # nn.Parameter -> gradient tensor (returned to torch.autograd)
        self.arena_params = {}  # nn.Parameter -> its slot of the gradient arena (assigned to .grad directly)

    def _entry(self, act: "Act", dtype=None):
        key = act.buf.data_ptr()
        e = self.bufs.get(key)
        if e is None:
            e = [torch.empty(act.buf.shape, dtype=dtype or act.buf.dtype, device=act.buf.device), []]
            self.bufs[key] = e
        return e

    @staticmethod
    def _covered(ranges, c0, c1):
        return any(a <= c0 and c1 <= b for a, b in ranges)

    @staticmethod
    def _touches(ranges, c0, c1):
        return any(a < c1 and c0 < b for a, b in ranges)

    def target(self, act: "Act", dtype=None):
        """-> (gradient view for act's channel range, already_initialised).  Call mark() after writing."""
        e = self._entry(act, dtype)
        c0, c1 = act.coff, act.coff + act.c
        if self._covered(e[1], c0, c1):
            inited = True
        elif not self._touches(e[1], c0, c1):
            inited = False
        else:                                   # partial overlap: zero the rest, then accumulate
            self._zero_uncovered(e, c0, c1)
            inited = True
        return Act(e[0], act.c, act.coff), inited

    def segments(self, act: "Act", align: int = 64):
        """Split act's channel range at the boundaries of what already holds gradient -> [(sub-view of the gradient, inited)] when
        every piece is `align`-channel aligned, else None.  A producer that can write channel sub-ranges separately (conv dgrad:
        rows of its weight pack) then overwrites the fresh pieces and accumulates into the others, instead of zero-filling the
        fresh part and accumulating over everything (the x5 slice of the 10240-channel PSP input already holds the critics'
        gradient when the bottleneck's dgrad arrives: that fill + read was 1.7 GB of traffic per step)."""
        e = self._entry(act)
        c0, c1 = act.coff, act.coff + act.c
        cuts = {c0, c1}
        for a, b in e[1]:
            if a < c1 and c0 < b:
                cuts.update((max(a, c0), min(b, c1)))
        cuts = sorted(cuts)
        if len(cuts) <= 2 or any((c - c0) % align for c in cuts):
            return None
        out = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            out.append((Act(e[0], b - a, a), self._covered(e[1], a, b)))
        return out

    def _zero_uncovered(self, e, c0, c1):
        c = c0
        for a, b in sorted(e[1]):
            if b <= c or a >= c1:
                continue
            if a > c:
                e[0][..., c:a].zero_()
            c = max(c, b)
        if c < c1:
            e[0][..., c:c1].zero_()
        e[1].append((c0, c1))

    def mark(self, act: "Act"):
        e = self._entry(act)
        c0, c1 = act.coff, act.coff + act.c
        if not self._covered(e[1], c0, c1):
            e[1].append((c0, c1))

    def get(self, act: "Act") -> Optional["Act"]:
        """Gradient of act, or None when nothing flowed into it."""
        e = self.bufs.get(act.buf.data_ptr())
        if e is None:
            return None
        c0, c1 = act.coff, act.coff + act.c
        if not self._touches(e[1], c0, c1):
            return None
        if not self._covered(e[1], c0, c1):
            self._zero_uncovered(e, c0, c1)
        return Act(e[0], act.c, act.coff)

    def param_sink(self, param):
        """-> (FP32 tensor of the parameter's shape the backward kernel writes into, accumulate flag).  With a gradient arena
        active (parallel.GradientReducer) that is the parameter's slot of the flat arena -- `.grad` becomes a view of it, no
        copy, static address; otherwise a tensor handed back to torch.autograd.  Call param_done() once the kernel is enqueued."""
        arena = grad_arena
        if arena is not None and arena.owns(param):
            t, acc = arena.sink(param)
            self.arena_params[param] = t
            return t, acc
        t = self.params.get(param)
        if t is not None:
            return t, True
        t = torch.empty(param.shape, dtype=torch.float32, device=param.device)
        self.params[param] = t
        return t, False

    def param_done(self, param):
        arena = grad_arena
        if arena is not None and param in self.arena_params:
            arena.produced(param)


class Tape:
    """Backward closures recorded by the ops below during a training-mode forward, replayed in reverse."""

    def __init__(self, input_needs_grad=False):
        self.ops = []
        self.live = set()        # data_ptrs of activation buffers whose gradient is needed

    def needs(self, act: Optional["Act"]) -> bool:
        return act is not None and act.buf.data_ptr() in self.live

    def require(self, act: "Act"):
        self.live.add(act.buf.data_ptr())

    def record(self, fn):
        self.ops.append(fn)

    def backward(self, grads: Grads):
        for fn in reversed(self.ops):
            fn(grads)
        self.ops = []


current_tape: Optional[Tape] = None
# > 1 while PSPNet.forward_pair runs: the batch holds that many equal groups of consecutive images (day batch, night batch) that
# share every convolution launch but are normalised by BatchNorm2d separately, exactly as consecutive forward calls would
bn_groups = 1
# parallel.GradientReducer registers itself here: parameter gradients are then written straight into its flat arena
grad_arena = None


def _wants(p) -> bool:
    return p is not None and p.requires_grad


# -------------------------------------------------------------------------------------------------- workspace
_workspace = {}
_capture_refs = None       # list while a CUDA graph is being captured: every workspace a captured kernel was pointed at


def begin_capture_refs() -> list:
    """Capture sites (PSPNet._forward_graph, graphs.GraphedStep) call this before capturing and keep the returned list with the
    graph: the grow-only workspace below is dropped when a later call needs a larger one, but a graph that baked the old
    address into its kernels must keep that block alive (and away from the allocator) for its own lifetime."""
    global _capture_refs
    _capture_refs = []
    return _capture_refs


def end_capture_refs():
    global _capture_refs
    _capture_refs = None


def workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch buffer per device; safe because every kernel is stream-ordered on one stream."""
    key = torch.device(device).index
    ws = _workspace.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _workspace[key] = None
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspace[key] = ws
    if _capture_refs is not None and not any(r is ws for r in _capture_refs):
        _capture_refs.append(ws)
    return ws


def _count(n=1):
    global launch_count
    launch_count += n


# -------------------------------------------------------------------------------------------------- layout
def from_nchw(t: torch.Tensor, dtype: torch.dtype, out: Optional[Act] = None) -> Act:
    """User NCHW tensor -> NHWC Act of the compute dtype (zero-copy when it already is one of our views and
    no destination slice is given)."""
    _lib.require_device()
    if out is None:
        a = act_from_view(t)
        if a is not None and a.dtype == dtype:
            return a
    src = t.detach()
    if src.dtype != torch.float32 or not src.is_contiguous():
        src = src.float().contiguous()
    n, c, h, w = src.shape
    if out is None:
        # wide odd channel counts (the 13-channel logits a critic consumes) get a 16-byte aligned pixel stride so that the vector /
        # TMA paths apply to the view and to its gradient; 1-4 channel images stay dense (the stems read them as they are)
        out = new_act(n, h, w, c, dtype, src.device, ld=(c + 7) // 8 * 8 if c > 4 else None)
    assert (out.n, out.h, out.w, out.c) == (n, h, w, c) and out.dtype == dtype
    with _timed("nchw_to_nhwc"):
        _lib.check(_lib.load().hn_nchw_to_nhwc(src.data_ptr(), C.byref(out.hn()), _stream()))
    _count()
    return out


def fuse_inputs(modal_1: torch.Tensor, modal_2: Optional[torch.Tensor], dtype: torch.dtype) -> Act:
    """Early fusion: torch.cat([modal_1, modal_2], 1) (cm/models/extractors.py:173) done by converting both
    NCHW inputs straight into channel slices of one NHWC buffer."""
    if modal_2 is None:
        return from_nchw(modal_1, dtype)
    n, c1, h, w = modal_1.shape
    c2 = modal_2.shape[1]
    cat = new_act(n, h, w, c1 + c2, dtype, modal_1.device)
    from_nchw(modal_1, dtype, cat.slice(0, c1))
    from_nchw(modal_2, dtype, cat.slice(c1, c2))
    return cat


def to_nchw_f32(a: Act) -> torch.Tensor:
    out = torch.empty((a.n, a.c, a.h, a.w), dtype=torch.float32, device=a.buf.device)
    with _timed("nhwc_to_nchw"):
        _lib.check(_lib.load().hn_nhwc_to_nchw(C.byref(a.hn()), out.data_ptr(), _stream()))
    _count()
    return out


# -------------------------------------------------------------------------------------------------- parameter caches
def _versions(*tensors):
    return tuple((t.data_ptr(), t._version) if t is not None else None for t in tensors)


# ---- multi-tensor repack: every plain pack built below is registered; after an optimizer step the first cache miss rebuilds ALL
# stale registered packs with one launch (hn_pack_weights_multi) instead of ~170 single-tensor launches per training step
MULTI_PACK = os.environ.get("HN_NO_MULTI_PACK") is None
_pack_registry = {}        # (id(conv), cache key) -> (weakref to conv, cache key, kind, phase)
_pack_tables = {}          # tuple of registry keys -> (jobs_dev, bmap_dev, n_blocks, host staging kept alive)


def _register_pack(conv, key, kind, phase=0):
    import weakref
    rk = (id(conv), key)
    if rk not in _pack_registry:
        _pack_registry[rk] = (weakref.ref(conv), key, kind, phase)


def repack_stale(device) -> int:
    """Rebuild every registered pack whose master weight has changed since it was packed, in place, in one launch.
    -> number of packs rebuilt (0: nothing stale, or fewer than two -- the caller's single-tensor path handles it)."""
    if not MULTI_PACK:
        return 0
    lib = _lib.load()
    stale = []
    for rk, (ref, key, kind, phase) in list(_pack_registry.items()):
        conv = ref()
        if conv is None:
            del _pack_registry[rk]
            continue
        hit = conv.__dict__.get("_hn_wcache", {}).get(key)
        if hit is None or conv.weight.device != device or key[-1] != device:
            continue
        ver = _versions(conv.weight)
        if hit[0] != ver:
            if hit[0][0][0] != ver[0][0] or not conv.weight.is_contiguous() or conv.weight.dtype != torch.float32:
                continue                        # the master moved / is not a dense FP32 tensor: rebuilt by the single-tensor path
            stale.append((rk, conv, key, kind, phase, hit, ver))
    if len(stale) < 2:
        return 0
    tkey = tuple(rk for rk, *_ in stale)
    tab = _pack_tables.get(tkey)
    if tab is None:
        import numpy as np
        chunk = lib.hn_pack_chunk()
        jobs = (_lib.HnPackJob * len(stale))()
        blocks = []
        for i, (rk, conv, key, kind, phase, hit, ver) in enumerate(stale):
            dtype = key[1] if kind != 2 else key[1]
            dst = hit[1] if kind != 2 else hit[1][phase]
            cout, cin, r, s = conv.weight.shape
            _lib.check(lib.hn_pack_job_init(C.byref(jobs[i]), conv.weight.data_ptr(), dst.data_ptr(), _HN_DTYPE[dtype], kind, cout, cin, r, s,
                                            conv.padding[0], phase))
            assert jobs[i].rows_pad * jobs[i].kpad == dst.numel(), "pack geometry mismatch"
            n = (dst.numel() + chunk - 1) // chunk
            blocks.append(np.stack([np.full(n, i, dtype=np.int32), np.arange(n, dtype=np.int32)], axis=1))
        bmap = np.ascontiguousarray(np.concatenate(blocks, axis=0))
        jobs_h = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8)
        bmap_h = torch.from_numpy(bmap)
        if not torch.cuda.is_current_stream_capturing():
            jobs_h, bmap_h = jobs_h.pin_memory(), bmap_h.pin_memory()
        jobs_d = torch.empty(jobs_h.shape, dtype=torch.uint8, device=device)
        bmap_d = torch.empty(bmap_h.shape, dtype=torch.int32, device=device)
        jobs_d.copy_(jobs_h, non_blocking=True)
        bmap_d.copy_(bmap_h, non_blocking=True)
        if len(_pack_tables) >= 8:
            _pack_tables.pop(next(iter(_pack_tables)))
        tab = (jobs_d, bmap_d, bmap.shape[0], (jobs_h, bmap_h))
        _pack_tables[tkey] = tab
    jobs_d, bmap_d, n_blocks, _ = tab
    _lib.check(lib.hn_pack_weights_multi(jobs_d.data_ptr(), bmap_d.data_ptr(), n_blocks, _stream()))
    _count()
    done = set()
    for rk, conv, key, kind, phase, hit, ver in stale:
        if (id(conv), key) in done:
            continue
        done.add((id(conv), key))
        conv.__dict__["_hn_wcache"][key] = (ver,) + tuple(hit[1:])
    return len(stale)


def packed_weight(conv: torch.nn.Conv2d, dtype: torch.dtype) -> torch.Tensor:
    """K-major pack [cout_pad][kpad] of the OIHW FP32 master weight, cached until the parameter changes."""
    lib = _lib.load()
    cache = conv.__dict__.setdefault("_hn_wcache", {})
    key = ("fwd", dtype, conv.weight.device)
    ver = _versions(conv.weight)
    hit = cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    if hit is not None and repack_stale(conv.weight.device) and cache[key][0] == ver:
        return cache[key][1]
    w = conv.weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    cout, cin, r, s = w.shape
    hdt = _HN_DTYPE[dtype]
    cout_pad, kpad = lib.hn_conv_cout_pad(cout, hdt), lib.hn_conv_kpad(cin, r, s)
    dst = hit[1] if hit is not None else torch.empty((cout_pad, kpad), dtype=dtype, device=w.device)
    _lib.check(lib.hn_pack_weight(w.data_ptr(), dst.data_ptr(), hdt, cout, cin, r, s, cout_pad, kpad, _stream()))
    _count()
    cache[key] = (ver, dst)
    _register_pack(conv, key, 0)
    return dst


def packed_weight_folded(conv: torch.nn.Conv2d, bn: Optional[torch.nn.BatchNorm2d], dtype: torch.dtype):
    """(pack, shift) for y = conv(x) folded with BatchNorm2d(eval) and the conv bias: the BN scale is multiplied into
    the filter rows before rounding to the compute dtype (what inference engines do), so the conv epilogue only adds a
    per-channel shift.  Cached on the versions of every tensor involved."""
    lib = _lib.load()
    scale, shift = folded_affine(conv, bn)
    if bn is None:
        return packed_weight(conv, dtype), shift
    cache = conv.__dict__.setdefault("_hn_wcache", {})
    key = ("folded", dtype, conv.weight.device)
    ver = _versions(conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    hit = cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1], shift
    w = conv.weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    cout, cin, r, s = w.shape
    hdt = _HN_DTYPE[dtype]
    cout_pad, kpad = lib.hn_conv_cout_pad(cout, hdt), lib.hn_conv_kpad(cin, r, s)
    dst = torch.empty((cout_pad, kpad), dtype=dtype, device=w.device)
    _lib.check(lib.hn_pack_weight_scaled(w.data_ptr(), scale.data_ptr(), dst.data_ptr(), hdt, cout, cin, r, s, cout_pad, kpad, _stream()))
    _count()
    cache[key] = (ver, dst)
    return dst, shift


def stem_ok(x: "Act", conv: torch.nn.Conv2d) -> bool:
    """7x7 stride-2 pad-3 stem on <= 4 input channels: runs without an im2col pass (hn_stem7x7s2_fwd)."""
    return (x.dtype == torch.bfloat16 and conv.kernel_size == (7, 7) and conv.stride == (2, 2) and conv.padding == (3, 3)
            and conv.dilation == (1, 1) and conv.in_channels <= 4 and conv.out_channels == 64 and STEM_FAST)


def packed_stem_weight(conv: torch.nn.Conv2d, bn: Optional[torch.nn.BatchNorm2d]):
    """[64][7][8][4] BF16 pack of a stem filter (BN(eval) scale folded in when `bn` is given) -> (pack, shift)."""
    lib = _lib.load()
    scale, shift = folded_affine(conv, bn) if (bn is not None or conv.bias is not None) else (None, None)
    cache = conv.__dict__.setdefault("_hn_wcache", {})
    key = ("stem", bn is not None, conv.weight.device)
    ver = _versions(conv.weight, conv.bias, *((bn.weight, bn.bias, bn.running_mean, bn.running_var) if bn is not None else ()))
    hit = cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1], shift
    w = conv.weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    dst = torch.empty((64, 224), dtype=torch.bfloat16, device=w.device)
    _lib.check(lib.hn_pack_stem_weight(w.data_ptr(), scale.data_ptr() if (bn is not None and scale is not None) else None, dst.data_ptr(),
                                       conv.out_channels, conv.in_channels, _stream()))
    _count()
    cache[key] = (ver, dst)
    return dst, shift


def stem_conv(x: "Act", conv: torch.nn.Conv2d, wp: torch.Tensor, shift, act=ACT_NONE, slope=0.0, slope_ptr=None,
              out: Optional["Act"] = None, out_dtype=None, stats: Optional[torch.Tensor] = None, stat_groups: int = 1) -> "Act":
    lib = _lib.load()
    ho, wo = conv_out_hw(x.h, x.w, conv)
    hpad, wpad = 2 * ho + 6, 2 * wo + 6
    flat = torch.empty((x.n * hpad * wpad * 4 + lib.hn_stem_pad_slack_bytes() // 2,), dtype=torch.bfloat16, device=x.buf.device)
    xpad = Act(flat[:x.n * hpad * wpad * 4].view(x.n, hpad, wpad, 4))         # + readable slack behind the last row (hn_stem_pad_slack_bytes)
    if out is None:
        out = new_act(x.n, ho, wo, conv.out_channels, out_dtype or x.dtype, x.buf.device)
    ep = _epilogue(None, shift, None, act, slope, slope_ptr, stats, stat_groups)
    timing = conv_timer is not None
    if timing:
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_a.record()
    _lib.check(lib.hn_stem_pad(C.byref(x.hn()), C.byref(xpad.hn()), _stream()))
    _lib.check(lib.hn_stem7x7s2_fwd(C.byref(xpad.hn()), wp.data_ptr(), conv.out_channels, C.byref(ep), C.byref(out.hn()), _stream()))
    if timing:
        ev_b.record()
        conv_timer.append((f"{x.c}->{conv.out_channels} k7 s2 d1 @{x.h}x{x.w}", ev_a, ev_b))
    _count(2)
    return out


def packed_weight_slice(conv: torch.nn.Conv2d, c0: int, c1: int, dtype: torch.dtype) -> torch.Tensor:
    """Pack of the input-channel slice [c0, c1) of a conv weight (the per-prior blocks of the PSP bottleneck)."""
    lib = _lib.load()
    cache = conv.__dict__.setdefault("_hn_wcache", {})
    key = ("slice", c0, c1, dtype, conv.weight.device)
    ver = _versions(conv.weight)
    hit = cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    w = conv.weight.detach()[:, c0:c1].float().contiguous()
    cout, cin, r, s = w.shape
    hdt = _HN_DTYPE[dtype]
    cout_pad, kpad = lib.hn_conv_cout_pad(cout, hdt), lib.hn_conv_kpad(cin, r, s)
    dst = torch.empty((cout_pad, kpad), dtype=dtype, device=w.device)
    _lib.check(lib.hn_pack_weight(w.data_ptr(), dst.data_ptr(), hdt, cout, cin, r, s, cout_pad, kpad, _stream()))
    _count()
    cache[key] = (ver, dst)
    return dst


def folded_affine(conv: torch.nn.Conv2d, bn: Optional[torch.nn.BatchNorm2d]):
    """(scale, shift) FP32 vectors for the conv epilogue: BatchNorm2d(eval) and the conv bias folded;
    (None, None) when the conv has neither."""
    bias = conv.bias
    if bn is None and bias is None:
        return None, None
    cache = conv.__dict__.setdefault("_hn_acache", {})
    if bn is not None:
        ver = _versions(bn.weight, bn.bias, bn.running_mean, bn.running_var, bias)
    else:
        ver = _versions(bias)
    hit = cache.get("fold")
    if hit is not None and hit[0] == ver:
        return hit[1], hit[2]
    cch = conv.out_channels
    dev = conv.weight.device
    scale = torch.empty(cch, dtype=torch.float32, device=dev)
    shift = torch.empty(cch, dtype=torch.float32, device=dev)
    p = lambda t: t.detach().data_ptr() if t is not None else None
    if bn is not None:
        rc = _lib.load().hn_bn_fold(p(bn.weight), p(bn.bias), p(bn.running_mean), p(bn.running_var), p(bias),
                                    float(bn.eps), scale.data_ptr(), shift.data_ptr(), cch, _stream())
    else:
        rc = _lib.load().hn_bn_fold(None, None, None, None, p(bias), BN_EPS_DEFAULT, scale.data_ptr(), shift.data_ptr(),
                                    cch, _stream())
        scale = None                       # bias only: no scale vector for the epilogue
    _lib.check(rc)
    _count()
    cache["fold"] = (ver, scale, shift)
    return scale, shift


# -------------------------------------------------------------------------------------------------- ops
def conv_out_hw(h, w, conv: torch.nn.Conv2d):
    k, s, p, d = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0]
    return (h + 2 * p - d * (k - 1) - 1) // s + 1, (w + 2 * p - d * (k - 1) - 1) // s + 1


def _epilogue(scale, shift, residual: Optional[Act], act, slope, slope_ptr, stats: Optional[torch.Tensor] = None, stat_groups: int = 1) -> HnEpilogue:
    ep = HnEpilogue()
    ep.stat_groups = stat_groups
    if stats is not None:        # FP64 [2, G*C]: per-(group, channel) sum / sum of squares accumulators
        c = stats.numel() // 2 if stats.dim() == 1 else stats.shape[1]
        ep.stat_sum, ep.stat_sqsum = stats.data_ptr(), stats.data_ptr() + 8 * c
    ep.scale = scale.data_ptr() if scale is not None else None
    ep.shift = shift.data_ptr() if shift is not None else None
    if residual is not None:
        ep.residual, ep.residual_ld = residual.ptr(), residual.ld
    ep.act, ep.slope = act, float(slope)
    ep.slope_ptr = slope_ptr.detach().data_ptr() if slope_ptr is not None else None
    return ep


def conv2d_raw(x: Act, wp: torch.Tensor, cout: int, k: int, stride: int, pad: int, dil: int, scale=None, shift=None,
               residual: Optional[Act] = None, act=ACT_NONE, slope=0.0, slope_ptr=None, out: Optional[Act] = None,
               out_dtype=None, out_hw=None, stats: Optional[torch.Tensor] = None, stat_groups: int = 1) -> Act:
    """One convolution launch on a packed weight matrix (see hn_conv2d_fwd)."""
    lib = _lib.load()
    ho = (x.h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    wo = (x.w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    if ho < 1 or wo < 1:
        raise RuntimeError(f"Calculated padded input size per channel: ({x.h + 2 * pad} x {x.w + 2 * pad}). "
                           f"Kernel size: ({k}, {k}). Kernel size can't be greater than actual input size")
    if out is None:
        out = new_act(x.n, ho, wo, cout, out_dtype or x.dtype, x.buf.device)
    assert (out.n, out.h, out.w, out.c) == (x.n, ho, wo, cout), ((out.n, out.h, out.w, out.c), (x.n, ho, wo, cout))
    if residual is not None:
        assert residual.dtype == out.dtype and (residual.n, residual.h, residual.w, residual.c) == (out.n, out.h, out.w, out.c)
    cv = HnConv(cout, k, k, stride, pad, dil)
    xh, yh = x.hn(), out.hn()
    ws_bytes = lib.hn_conv2d_workspace_bytes(C.byref(xh), C.byref(cv))
    ws_ptr = None
    if ws_bytes:
        ws_ptr = workspace(ws_bytes, x.buf.device).data_ptr()
        _count()
    ep = _epilogue(scale, shift, residual, act, slope, slope_ptr, stats, stat_groups)
    timing = conv_timer is not None
    if timing:
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_a.record()
    _lib.check(lib.hn_conv2d_fwd(C.byref(xh), wp.data_ptr(), C.byref(cv), C.byref(ep), C.byref(yh), ws_ptr, ws_bytes,
                                 _stream()))
    if timing:
        ev_b.record()
        conv_timer.append((f"{x.c}->{cout} k{k} s{stride} d{dil} @{x.h}x{x.w}", ev_a, ev_b))
    _count()
    return out


def conv2d(x: Act, conv: torch.nn.Conv2d, scale=None, shift=None, residual: Optional[Act] = None, act=ACT_NONE,
           slope=0.0, slope_ptr=None, out: Optional[Act] = None, out_dtype=None, stats: Optional[torch.Tensor] = None,
           stat_groups: int = 1) -> Act:
    """y = act(conv(x) * scale + shift + residual) in one launch (plus an im2col gather for strided /
    small-Cin shapes on the BF16 path)."""
    assert conv.groups == 1 and conv.kernel_size[0] == conv.kernel_size[1] and conv.stride[0] == conv.stride[1]
    assert conv.padding[0] == conv.padding[1] and conv.dilation[0] == conv.dilation[1]
    if x.c != conv.in_channels:
        raise RuntimeError(f"expected input with {conv.in_channels} channels, got {x.c}")
    return conv2d_raw(x, packed_weight(conv, x.dtype), conv.out_channels, conv.kernel_size[0], conv.stride[0], conv.padding[0],
                      conv.dilation[0], scale, shift, residual, act, slope, slope_ptr, out, out_dtype, stats=stats, stat_groups=stat_groups)


def affine_act(x: Act, scale, shift, residual: Optional[Act], act, slope=0.0, slope_ptr=None, out: Optional[Act] = None) -> Act:
    out = out or x
    ep = _epilogue(scale, shift, residual, act, slope, slope_ptr)
    _lib.check(_lib.load().hn_affine_act(C.byref(x.hn()), C.byref(ep), C.byref(out.hn()), _stream()))
    _count()
    return out


def batchnorm_train_affine(x: Act, bn: torch.nn.BatchNorm2d):
    """Batch statistics of x (FP64 accumulation) -> (scale, shift, mean, invstd) for the apply pass; updates the module's
    running statistics exactly like nn.BatchNorm2d in train mode (momentum, unbiased variance,
    num_batches_tracked).  One memset + ONE kernel (hn_bn_batch_stats: the last CTA finalizes)."""
    lib = _lib.load()
    dev = x.buf.device
    out = torch.empty((4, x.c), dtype=torch.float32, device=dev)     # scale, shift, save_mean, save_invstd
    track = bn.track_running_stats and bn.running_mean is not None
    p = lambda t: t.detach().data_ptr() if t is not None else None
    momentum = 0.0
    nbt = None
    if track:
        if bn.momentum is None:      # cumulative moving average: the factor depends on the counter's value
            bn.num_batches_tracked.add_(1)
            momentum = 1.0 / float(bn.num_batches_tracked.item())
        else:
            momentum = bn.momentum
            nbt = p(bn.num_batches_tracked)
    scratch = torch.empty((2 * x.c + 1,), dtype=torch.float64, device=dev)
    _lib.check(lib.hn_bn_batch_stats(C.byref(x.hn()), scratch.data_ptr(), p(bn.weight), p(bn.bias), float(bn.eps), float(momentum),
                                     p(bn.running_mean) if track else None, p(bn.running_var) if track else None, nbt,
                                     out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), _stream()))
    _count(2)
    if track:   # in-place kernel writes bypass torch's version counter; bump it so folded caches refresh
        bn.running_mean._bump_version() if hasattr(bn.running_mean, "_bump_version") else bn.running_mean.add_(0)
        bn.running_var._bump_version() if hasattr(bn.running_var, "_bump_version") else bn.running_var.add_(0)
    return out[0], out[1], out[2], out[3]


BN_FUSED_STATS = os.environ.get("HN_NO_FUSED_BN_STATS") is None
# Per-layer placement of the train-mode BatchNorm statistics on the BF16 engine: convolutions with fewer k-blocks than this leave
# them to a separate pass over the stored tensor (hn_channel_stats, one launch per statistics group) instead of the epilogue --
# their tiles are a handful of MMAs, the epilogue is the bound, and the statistics code roughly doubles it (measured: train_seg 62.96 ms at 0 = always
# the epilogue, 62.5 at 9, 62.2 at 17 = every 1x1 up to 1024 input channels and the 64-channel 3x3s; beyond that within run-to-run noise).
STATS_EPILOGUE_MIN_KB = int(os.environ.get("HN_STATS_EPILOGUE_MIN_KB", "17"))
_stats_chunk = {}          # device index -> [zeroed FP64 chunk, next free element]: one memset serves ~30 BN layers


def reset_stats_pool():
    """Forget the current accumulator chunk.  A chunk is zeroed once, when it is created, so a CUDA-graph capture must not carve
    slices from a chunk created before it (the memset would not be part of the graph and replays would keep accumulating), nor
    may eager code afterwards use a chunk that lives in a graph's private pool: graphs.GraphedStep calls this on both sides."""
    _stats_chunk.clear()


def _stats_alloc(c: int, device) -> torch.Tensor:
    """Zeroed FP64 [2, c] accumulator carved from a pooled chunk (each slice is handed out once; the chunk dies with its views)."""
    key = torch.device(device).index
    need = 2 * c
    ent = _stats_chunk.get(key)
    if ent is None or ent[1] + need > ent[0].numel():
        ent = [torch.zeros((max(32768, need),), dtype=torch.float64, device=device), 0]
        _stats_chunk[key] = ent
        _count()
    view = ent[0][ent[1]:ent[1] + need].view(2, c)
    ent[1] += need
    return view



def batchnorm_finalize(sums: torch.Tensor, count: int, bn: torch.nn.BatchNorm2d):
    """(scale, shift, mean, invstd) from the FP64 per-channel sums a convolution's epilogue accumulated (stats=...), plus the
    running-statistic update of nn.BatchNorm2d in train mode -- one tiny launch."""
    lib = _lib.load()
    cch = bn.num_features
    out = torch.empty((4, cch), dtype=torch.float32, device=sums.device)
    track = bn.track_running_stats and bn.running_mean is not None
    p = lambda t: t.detach().data_ptr() if t is not None else None
    momentum, nbt = 0.0, None
    if track:
        if bn.momentum is None:
            bn.num_batches_tracked.add_(1)
            momentum = 1.0 / float(bn.num_batches_tracked.item())
        else:
            momentum, nbt = bn.momentum, p(bn.num_batches_tracked)
    _lib.check(lib.hn_bn_finalize_tracked(sums.data_ptr(), sums.data_ptr() + 8 * cch, count, p(bn.weight), p(bn.bias), float(bn.eps),
                                          float(momentum), p(bn.running_mean) if track else None, p(bn.running_var) if track else None,
                                          nbt, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), cch, _stream()))
    _count()
    if track:
        bn.running_mean._bump_version() if hasattr(bn.running_mean, "_bump_version") else bn.running_mean.add_(0)
        bn.running_var._bump_version() if hasattr(bn.running_var, "_bump_version") else bn.running_var.add_(0)
    return out[0], out[1], out[2], out[3]


def batchnorm_apply_train(raw: Act, sums: torch.Tensor, bn: torch.nn.BatchNorm2d, residual: Optional[Act], act, slope, slope_ptr, out: Act,
                          groups: int = 1):
    """out = act(BN_train(raw) + residual) from the FP64 sums of a conv with fused statistics, and the running-statistic update of
    nn.BatchNorm2d -- one launch (hn_bn_apply_train).  -> (scale, shift, mean, invstd) for the backward."""
    cch = bn.num_features
    vec = torch.empty((4, groups * cch), dtype=torch.float32, device=sums.device)      # scale, shift, mean, invstd: [groups][C] each
    track = bn.track_running_stats and bn.running_mean is not None
    p = lambda t: t.detach().data_ptr() if t is not None else None
    ep = _epilogue(None, None, residual, act, slope, slope_ptr, stat_groups=groups)
    _lib.check(_lib.load().hn_bn_apply_train(C.byref(raw.hn()), sums.data_ptr(), sums.data_ptr() + 8 * groups * cch, (raw.n // groups) * raw.h * raw.w, p(bn.weight),
                                             p(bn.bias), float(bn.eps), float(bn.momentum), p(bn.running_mean) if track else None,
                                             p(bn.running_var) if track else None, p(bn.num_batches_tracked) if track else None, C.byref(ep),
                                             C.byref(out.hn()), vec[0].data_ptr(), vec[1].data_ptr(), vec[2].data_ptr(), vec[3].data_ptr(), _stream()))
    _count()
    if track:
        torch.autograd.graph.increment_version([bn.running_mean, bn.running_var])
    return vec[0], vec[1], vec[2], vec[3]


def _conv_param_backward(grads: Grads, tape: Tape, x: Act, conv, dz: Act):
    """Shared tail of every conv backward: weight / bias gradients and the input gradient."""
    if dz.dtype != x.dtype:                       # e.g. FP32 critic map gradient -> BF16 operands
        cast = new_act(dz.n, dz.h, dz.w, dz.c, x.dtype, dz.buf.device, ld=(dz.c + 7) // 8 * 8)
        accumulate(dz, cast, False)
        dz = cast
    if _wants(conv.weight):
        dst, acc = grads.param_sink(conv.weight)
        conv2d_wgrad(x, dz, conv, out=dst, accumulate=acc)
        grads.param_done(conv.weight)
    if _wants(conv.bias):
        dst, acc = grads.param_sink(conv.bias)
        vec_to_grad(channel_sums(dz)[0], conv.out_channels, out=dst, accumulate=acc)
        grads.param_done(conv.bias)
    if tape.needs(x):
        segs = grads.segments(x) if (conv.stride[0] == 1 and x.dtype == torch.bfloat16) else None
        if segs is not None:
            for sub, inited in segs:          # partially initialised target: per channel range, fresh pieces are written, the rest accumulated
                conv2d_dgrad(dz, conv, x.h, x.w, out=sub, accumulate=inited, cin_range=(sub.coff - x.coff, sub.coff - x.coff + sub.c))
        else:
            gx, inited = grads.target(x)
            conv2d_dgrad(dz, conv, x.h, x.w, out=gx, accumulate=inited)
        grads.mark(x)


def conv_bn_act(x: Act, conv, bn, act=ACT_NONE, slope=0.0, slope_ptr=None, residual: Optional[Act] = None,
                out: Optional[Act] = None, bn_training: Optional[bool] = None) -> Act:
    """conv -> BatchNorm2d -> (+residual) -> activation.
    eval BN: folded into the conv epilogue (1 launch).  train BN: conv (+bias) -> batch statistics -> one fused
    normalise + residual + activation pass.  With a tape active the backward closure is recorded."""
    tape = current_tape
    if bn_training is None:      # like nn.BatchNorm2d: batch statistics iff the BN module itself is in train mode
        bn_training = bn is not None and (bn.training or bn.running_mean is None)
    if bn is None or not bn_training:
        scale, shift = folded_affine(conv, bn)
        if residual is None and stem_ok(x, conv):
            wp, shift = packed_stem_weight(conv, bn)
            y = stem_conv(x, conv, wp, shift, act, slope, slope_ptr, out)
        elif bn is not None and x.dtype == torch.bfloat16:
            # BF16 engine: BN scale folded into the packed filter, the epilogue only adds the shift
            wp, shift = packed_weight_folded(conv, bn, x.dtype)
            y = conv2d_raw(x, wp, conv.out_channels, conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0], None, shift,
                           residual, act, slope, slope_ptr, out)
        else:
            y = conv2d(x, conv, scale, shift, residual, act, slope, slope_ptr, out)
        if tape is not None:
            _record_conv_affine(tape, x, conv, bn, scale, y, residual, act, slope, slope_ptr)
        return y
    scale, shift = folded_affine(conv, None)                      # conv bias only
    # The pre-normalisation tensor stays FP32: (x - mean) * invstd amplifies BF16 rounding of x by |mean|/std,
    # which is large for channels with little spatial variation.  Statistics and the normalise pass read the
    # FP32 values; only the normalised activation is rounded to the compute dtype.
    raw_fp32 = BN_TRAIN_RAW_FP32
    # BF16 engine with FP32 pre-normalisation output: the batch statistics are accumulated by the conv epilogue itself (FP32 per
    # 32-pixel block, FP64 atomics), so train-mode BN costs one tiny finalize launch instead of a pass over the tensor
    fused = BN_FUSED_STATS and x.dtype == torch.bfloat16 and conv.out_channels >= 17
    # statistics groups (PSPNet.forward_pair): the batch is the day batch followed by the night batch, convolved as one, normalised
    # per domain -- only on the fused-statistics path (BF16 engine)
    G = bn_groups if (bn_groups > 1 and x.n % bn_groups == 0) else 1
    if G > 1 and not (fused and bn.momentum is not None and conv.out_channels % 8 == 0):
        raise NotImplementedError("paired (grouped-statistics) forward needs the BF16 engine's fused BatchNorm statistics")
    sums = _stats_alloc(G * conv.out_channels, x.buf.device) if fused else None
    if stem_ok(x, conv):
        wp, shift = packed_stem_weight(conv, None)
        raw = stem_conv(x, conv, wp, shift, out_dtype=torch.float32 if raw_fp32 else None, stats=sums, stat_groups=G)
    else:
        kblocks = conv.kernel_size[0] * conv.kernel_size[1] * ((conv.in_channels + 63) // 64)
        in_epilogue = (not fused) or kblocks >= STATS_EPILOGUE_MIN_KB or conv.out_channels % 8 != 0
        raw = conv2d(x, conv, scale, shift, None, ACT_NONE, out_dtype=torch.float32 if raw_fp32 else None, stats=sums if in_epilogue else None,
                     stat_groups=G)
        if fused and not in_epilogue:        # the same [2][G][C] FP64 sums, from a pass over the stored tensor, one group at a time
            lib = _lib.load()
            ng, cch = raw.n // G, conv.out_channels
            for g in range(G):
                sub = Act(raw.buf[g * ng:(g + 1) * ng], raw.c, raw.coff)
                _lib.check(lib.hn_channel_stats(C.byref(sub.hn()), sums[0, g * cch:].data_ptr(), sums[1, g * cch:].data_ptr(), _stream()))
                _count(3)
    if out is None:
        in_place = raw.dtype == x.dtype and tape is None
        out = raw if in_place else new_act(raw.n, raw.h, raw.w, raw.c, x.dtype, x.buf.device)
    if fused and bn.momentum is not None and raw.c % 8 == 0 and raw.ld % 8 == 0 and out.ld % 8 == 0:
        # statistics came out of the conv epilogue: finalize + normalise + residual + activation in ONE launch
        bscale, bshift, mean, invstd = batchnorm_apply_train(raw, sums, bn, residual, act, slope, slope_ptr, out, groups=G)
        y = out
    else:
        assert G == 1
        if fused:
            bscale, bshift, mean, invstd = batchnorm_finalize(sums, raw.n * raw.h * raw.w, bn)
        else:
            bscale, bshift, mean, invstd = batchnorm_train_affine(raw, bn)
        y = affine_act(raw, bscale, bshift, residual, act, slope, slope_ptr, out=out)
    if tape is not None:
        live = tape.needs(x) or tape.needs(residual) or any(_wants(p) for p in (conv.weight, conv.bias, bn.weight, bn.bias, slope_ptr))
        if live:
            tape.require(y)

            def backward(grads: Grads):
                dout = grads.get(y)
                if dout is None:
                    return
                dres, dres_acc = None, False
                if tape.needs(residual):
                    dres, dres_acc = grads.target(residual)
                prelu = _wants(slope_ptr)
                # no residual: the pre-activation is recomputed from raw with the forward's own scale/shift (y is not streamed)
                fz = (bscale, bshift) if residual is None else (None, None)
                sinks = [grads.param_sink(q) if w else (None, False) for q, w in ((bn.bias, _wants(bn.bias)), (bn.weight, _wants(bn.weight)),
                                                                                (slope_ptr, prelu))]
                draw = bn_bwd(dout, y, raw, mean, invstd, bn.weight, act, slope, slope_ptr, dres, dres_acc, sinks, fz[0], fz[1], groups=G)
                if dres is not None:
                    grads.mark(residual)
                for q, (t, _) in zip((bn.bias, bn.weight, slope_ptr), sinks):
                    if t is not None:
                        grads.param_done(q)
                _conv_param_backward(grads, tape, x, conv, draw)

            tape.record(backward)
    return y


def _record_conv_affine(tape: Tape, x: Act, conv, bn, scale, y: Act, residual, act, slope, slope_ptr):
    """Backward of y = act(conv(x)*scale + shift + residual) with constant scale (conv bias / eval-mode BN)."""
    if bn is not None and (_wants(bn.weight) or _wants(bn.bias)):
        raise NotImplementedError("gradients of BatchNorm2d affine parameters need the BN module in train mode")
    live = tape.needs(x) or tape.needs(residual) or any(_wants(p) for p in (conv.weight, conv.bias, slope_ptr))
    if not live:
        return
    if _wants(slope_ptr):
        raise NotImplementedError("PReLU slope gradient is only implemented behind a train-mode BatchNorm2d")
    tape.require(y)

    def backward(grads: Grads):
        dout = grads.get(y)
        if dout is None:
            return
        sl = float(slope_ptr.detach().item()) if slope_ptr is not None else slope
        dz = act_bwd(dout, y, act, sl) if act != ACT_NONE else dout
        if tape.needs(residual):
            gr, inited = grads.target(residual)
            accumulate(dz, gr, inited)
            grads.mark(residual)
        if bn is not None:            # eval-mode BN: constant per-channel scale in front of the conv output
            dz = affine_act(dz, scale, None, None, ACT_NONE, out=new_act(dz.n, dz.h, dz.w, dz.c, dz.dtype, dz.buf.device))
        _conv_param_backward(grads, tape, x, conv, dz)

    tape.record(backward)


def maxpool3x3s2(x: Act, out: Optional[Act] = None) -> Act:
    tape = current_tape
    if tape is not None and tape.needs(x):
        y, idx = maxpool3x3s2_idx(x, out)
        tape.require(y)

        def backward(grads: Grads):
            dy = grads.get(y)
            if dy is None:
                return
            gx, inited = grads.target(x)
            maxpool3x3s2_bwd(dy, idx, gx, inited)
            grads.mark(x)

        tape.record(backward)
        return y
    ho, wo = (x.h - 1) // 2 + 1, (x.w - 1) // 2 + 1
    out = out or new_act(x.n, ho, wo, x.c, x.dtype, x.buf.device)
    with _timed("maxpool"):
        _lib.check(_lib.load().hn_maxpool3x3s2_fwd(C.byref(x.hn()), C.byref(out.hn()), _stream()))
    _count()
    return out


def pyramid_pool(x: Act, sizes: Sequence[int]):
    """All adaptive average pools of the PSP module in one pass -> list of dense Acts [N, s, s, C]."""
    lib = _lib.load()
    arr = (C.c_int32 * len(sizes))(*sizes)
    nb = sum(s * s for s in sizes)
    out = torch.empty((x.n * nb * x.c,), dtype=x.dtype, device=x.buf.device)
    xh = x.hn()
    ws_bytes = lib.hn_pyramid_pool_workspace_bytes(C.byref(xh), arr, len(sizes))
    ws = workspace(ws_bytes, x.buf.device)
    with _timed("pyramid_pool"):
        _lib.check(lib.hn_pyramid_pool_fwd(C.byref(xh), arr, len(sizes), out.data_ptr(), ws.data_ptr(), ws_bytes, _stream()))
    _count(2)
    acts, off = [], 0
    for s in sizes:
        cnt = x.n * s * s * x.c
        acts.append(Act(out[off:off + cnt].view(x.n, s, s, x.c)))
        off += cnt
    tape = current_tape
    if tape is not None and tape.needs(x):
        for a in acts:
            tape.require(a)

        def backward(grads: Grads):
            gs = [grads.get(a) for a in acts]
            if all(g is None for g in gs):
                return
            dpool = torch.zeros_like(out)        # the per-size blocks, in the forward's layout
            o = 0
            for a, g in zip(acts, gs):
                cnt = a.buf.numel()
                if g is not None:
                    accumulate(g, Act(dpool[o:o + cnt].view(a.buf.shape)), False)
                o += cnt
            gx, inited = grads.target(x)
            pyramid_pool_bwd(dpool, list(sizes), gx, inited)
            grads.mark(x)

        tape.record(backward)
    return acts


def bilinear(x: Act, h: int, w: int, out: Optional[Act] = None, out_dtype=None) -> Act:
    out = out or new_act(x.n, h, w, x.c, out_dtype or x.dtype, x.buf.device)
    assert (out.n, out.h, out.w, out.c) == (x.n, h, w, x.c)
    with _timed("bilinear"):
        _lib.check(_lib.load().hn_bilinear_fwd(C.byref(x.hn()), C.byref(out.hn()), _stream()))
    _count()
    tape = current_tape
    if tape is not None and tape.needs(x):
        y = out
        tape.require(y)

        def backward(grads: Grads):
            dy = grads.get(y)
            if dy is None:
                return
            gx, inited = grads.target(x)
            bilinear_bwd(dy, gx, inited)
            grads.mark(x)

        tape.record(backward)
    return out


_dropout_state = {}        # device index -> int64 [2] on the device: {seed, masks drawn so far}


def dropout_seed(seed: Optional[int] = None, device=None):
    """(Re)seed the Dropout2d generator of `device`: the same seed reproduces the same sequence of channel masks.  Without an
    explicit seed the state is created from torch.initial_seed() (so torch.manual_seed() before the first training forward
    is enough).  The state lives in device memory and the mask kernel advances it itself: replays of a captured CUDA graph
    draw a fresh mask each time (hn_dropout2d_scale)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if seed is None:
        seed = torch.initial_seed()
    st = torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=dev)
    _dropout_state[dev.index] = st
    return st


def dropout2d(x: Act, p: float, mask: Optional[torch.Tensor] = None) -> Act:
    """nn.Dropout2d in train mode: whole channels of each image zeroed with probability p, survivors scaled by
    1/(1-p); p = 1 zeroes everything (torch does the same).  `mask` (N, C) of {0,1} keep flags may be injected for
    parity runs (the reference's RNG stream cannot be reproduced by another kernel); otherwise the [N, C] scale vector
    comes from the keyed Philox kernel (dropout_seed)."""
    if not 0.0 <= p <= 1.0:
        raise ValueError(f"dropout probability has to be between 0 and 1, but got {p}")
    if mask is not None:
        assert tuple(mask.shape) == (x.n, x.c), f"injected Dropout2d mask must be (N, C) = {(x.n, x.c)}, got {tuple(mask.shape)}"
        keep = 1.0 / (1.0 - p) if p < 1.0 else 0.0
        scale = (mask.to(device=x.buf.device, dtype=torch.float32) * keep).contiguous()
    else:
        st = _dropout_state.get(x.buf.device.index)
        if st is None:
            st = dropout_seed(None, x.buf.device)
        scale = torch.empty((x.n, x.c), dtype=torch.float32, device=x.buf.device)
        _lib.check(_lib.load().hn_dropout2d_scale(st.data_ptr(), x.n * x.c, float(p), scale.data_ptr(), _stream()))
        _count()
    ep = _epilogue(scale, None, None, ACT_NONE, 0.0, None)
    ep.per_image = 1
    out = new_act(x.n, x.h, x.w, x.c, x.dtype, x.buf.device)
    _lib.check(_lib.load().hn_affine_act(C.byref(x.hn()), C.byref(ep), C.byref(out.hn()), _stream()))
    _count()
    tape = current_tape
    if tape is not None and tape.needs(x):
        y = out
        tape.require(y)

        def backward(grads: Grads):
            dy = grads.get(y)
            if dy is None:
                return
            gx, inited = grads.target(x)
            if inited:
                tmp = new_act(x.n, x.h, x.w, x.c, dy.dtype, x.buf.device)
                _lib.check(_lib.load().hn_affine_act(C.byref(dy.hn()), C.byref(ep), C.byref(tmp.hn()), _stream()))
                accumulate(tmp, gx, True)
                _count()
            else:
                _lib.check(_lib.load().hn_affine_act(C.byref(dy.hn()), C.byref(ep), C.byref(gx.hn()), _stream()))
                _count()
            grads.mark(x)
            _keep = scale          # the mask vector must outlive the closure's ep pointer

        tape.record(backward)
    return out


# -------------------------------------------------------------------------------------------------- backward ops
def record_projected_bottleneck(tape: Tape, feats: Act, priors: Sequence[Act], y: Act, bott: torch.nn.Conv2d):
    """Backward of the PSP bottleneck in its projected form (pspnet.PSPModule._run_projected_train):
        y = relu(W_f feats + b + sum_s up(W_s prior_s)),   W = [W_0 | ... | W_{S-1} | W_f] the column blocks of bott.weight.
    One closure for the whole block, because the five column blocks of dW share one gradient sink:
        dz = dy * (y > 0);  db = sum dz;  dW_f = dz^T feats,  dfeats += W_f^T dz;
        dproj_s = up^T(dz)  (the adjoint of the bilinear upsample, [N, s, s, Cout]);  dW_s = dproj_s^T prior_s,  dprior_s = W_s^T dproj_s."""
    import types
    f = feats.c
    S = len(priors)
    cout = bott.out_channels
    live = tape.needs(feats) or any(tape.needs(q) for q in priors) or _wants(bott.weight) or _wants(bott.bias)
    if not live:
        return
    tape.require(y)
    block = types.SimpleNamespace(out_channels=cout, in_channels=f, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0), dilation=(1, 1))

    def backward(grads: Grads):
        dout = grads.get(y)
        if dout is None:
            return
        dz = act_bwd(dout, y, ACT_RELU)
        if _wants(bott.bias):
            dst, acc = grads.param_sink(bott.bias)
            vec_to_grad(channel_sums(dz)[0], cout, out=dst, accumulate=acc)
            grads.param_done(bott.bias)
        wsink, wacc = (None, False)
        if _wants(bott.weight):
            wsink, wacc = grads.param_sink(bott.weight)
            wsink = wsink.view(cout, (S + 1) * f)

        def block_backward(x: Act, g: Act, i: int):
            if wsink is not None:
                dw = conv2d_wgrad(x, g, block)                           # [cout, f, 1, 1]
                col = wsink[:, i * f:(i + 1) * f]
                col.add_(dw.view(cout, f)) if wacc else col.copy_(dw.view(cout, f))
                _count()
            if tape.needs(x):
                segs = grads.segments(x) if i == S else None
                if segs is not None:
                    for sub, inited in segs:
                        a = sub.coff - x.coff
                        conv2d_dgrad(g, bott, x.h, x.w, out=sub, accumulate=inited, cin_range=(i * f + a, i * f + a + sub.c))
                else:
                    gx, inited = grads.target(x)
                    conv2d_dgrad(g, bott, x.h, x.w, out=gx, accumulate=inited, cin_range=(i * f, (i + 1) * f))
                grads.mark(x)

        for i, q in enumerate(priors):
            dproj = new_act(q.n, q.h, q.w, cout, dz.dtype, dz.buf.device)
            bilinear_bwd(dz, dproj, False)
            block_backward(q, dproj, i)
        block_backward(feats, dz, S)
        if wsink is not None:
            grads.param_done(bott.weight)

    tape.record(backward)


def packed_weight_dgrad(conv: torch.nn.Conv2d, dtype: torch.dtype) -> torch.Tensor:
    """Pack for dgrad-as-forward-conv: [cin_pad][kpad'], k' = ((R-1-r)*S + (S-1-s))*Cout + o (flipped taps, in/out
    channels swapped), cached like the forward pack."""
    lib = _lib.load()
    cache = conv.__dict__.setdefault("_hn_wcache", {})
    key = ("dgrad", dtype, conv.weight.device)
    ver = _versions(conv.weight)
    hit = cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    if hit is not None and repack_stale(conv.weight.device) and cache[key][0] == ver:
        return cache[key][1]
    w = conv.weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    cout, cin, r, s = w.shape
    hdt = _HN_DTYPE[dtype]
    cin_pad, kpad = lib.hn_conv_cout_pad(cin, hdt), lib.hn_conv_kpad(cout, r, s)
    dst = hit[1] if hit is not None else torch.empty((cin_pad, kpad), dtype=dtype, device=w.device)
    _lib.check(lib.hn_pack_weight_dgrad(w.data_ptr(), dst.data_ptr(), hdt, cout, cin, r, s, cin_pad, kpad, _stream()))
    _count()
    cache[key] = (ver, dst)
    _register_pack(conv, key, 1)
    return dst


def packed_weight_dgrad_phases(conv: torch.nn.Conv2d, dtype: torch.dtype):
    """The four parity-phase sub-filter packs of a stride-2 convolution's dgrad (hn_pack_weight_dgrad_phase), cached like the other
    packs -> (list of 4 tensors or None for phases no tap reaches, ctypes array of their addresses)."""
    lib = _lib.load()
    cache = conv.__dict__.setdefault("_hn_wcache", {})
    key = ("dgrad_s2", dtype, conv.weight.device)
    ver = _versions(conv.weight)
    hit = cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1], hit[2]
    if hit is not None and repack_stale(conv.weight.device) and cache[key][0] == ver:
        return cache[key][1], cache[key][2]
    w = conv.weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    cout, cin, r, s = w.shape
    hdt = _HN_DTYPE[dtype]
    cin_pad = lib.hn_conv_cout_pad(cin, hdt)
    packs = []
    for phase in range(4):
        kpad = lib.hn_dgrad_s2_phase_kpad(cout, r, s, conv.padding[0], phase)
        if kpad == 0:
            packs.append(None)
            continue
        dst = hit[1][phase] if hit is not None else torch.empty((cin_pad, kpad), dtype=dtype, device=w.device)
        _lib.check(lib.hn_pack_weight_dgrad_phase(w.data_ptr(), dst.data_ptr(), hdt, cout, cin, r, s, conv.padding[0], phase, cin_pad, _stream()))
        _count()
        packs.append(dst)
    ptrs = (C.c_void_p * 4)(*[t.data_ptr() if t is not None else None for t in packs])
    cache[key] = (ver, packs, ptrs)
    for phase in range(4):
        if packs[phase] is not None:
            _pack_registry.setdefault((id(conv), key, phase), (__import__("weakref").ref(conv), key, 2, phase))
    return packs, ptrs


DGRAD_S2_PHASES = os.environ.get("HN_NO_DGRAD_PHASES") is None


def conv2d_dgrad(dy: Act, conv: torch.nn.Conv2d, in_h: int, in_w: int, out: Optional[Act] = None,
                 accumulate: bool = False, cin_range=None) -> Act:
    """dX of a convolution.  Stride 1: a stride-1 convolution of dY with the flipped, channel-transposed filter.  Stride 2 on the
    BF16 engine: four parity phases, each a stride-1 correlation of dY with a sub-filter, written straight onto its sub-lattice of
    dX (hn_conv2d_dgrad_s2) -- no zero-inserted gradient.  Otherwise (FP32 parity path, odd shapes): zero insertion (hn_dilate)
    followed by the stride-1 form.  `accumulate` adds into `out` through the epilogue's residual input."""
    k, st, pad, dil = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0]
    lib = _lib.load()
    if st == 2 and DGRAD_S2_PHASES and dy.dtype == torch.bfloat16:
        if out is None:
            out, accumulate = new_act(dy.n, in_h, in_w, conv.in_channels, dy.dtype, dy.buf.device, ld=(conv.in_channels + 7) // 8 * 8), False
        cv = HnConv(conv.out_channels, k, k, st, pad, dil)
        dyh, oh = dy.hn(), out.hn()
        if (out.h, out.w) == (in_h, in_w) and lib.hn_conv2d_dgrad_s2_ok(C.byref(dyh), C.byref(cv), C.byref(oh)):
            packs, ptrs = packed_weight_dgrad_phases(conv, dy.dtype)
            timing = conv_timer is not None
            if timing:
                ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev_a.record()
            _lib.check(lib.hn_conv2d_dgrad_s2(C.byref(dyh), ptrs, C.byref(cv), C.byref(oh), int(accumulate), _stream()))
            if timing:
                ev_b.record()
                conv_timer.append((f"dgrad {conv.in_channels}<-{conv.out_channels} k{k} s2 @{in_h}x{in_w}", ev_a, ev_b))
            _count(4)
            return out
    padp = dil * (k - 1) - pad
    assert padp >= 0, "dgrad needs pad <= dil*(k-1)"
    g = dy
    # the (zero-inserted) gradient map must be exactly this large for a stride-1 conv with pad' to return in_h x in_w
    hu, wu = in_h - dil * (k - 1) + 2 * pad, in_w - dil * (k - 1) + 2 * pad
    if st > 1 or (hu, wu) != (dy.h, dy.w):
        up = new_act(dy.n, hu, wu, dy.c, dy.dtype, dy.buf.device, ld=(dy.c + 7) // 8 * 8)
        _lib.check(lib.hn_dilate(C.byref(dy.hn()), st, C.byref(up.hn()), _stream()))
        _count()
        g = up
    wp = packed_weight_dgrad(conv, dy.dtype)
    res = out if accumulate else None
    cin = conv.in_channels
    if cin_range is not None:       # only input channels [a, b): a row range of the pack (rows = input channels), stride-1 form only
        a, b = cin_range
        assert st == 1 and a % 64 == 0 and (b - a) % 64 == 0 and out is not None and out.c == b - a
        wp, cin = wp[a:b], b - a
    return conv2d_raw(g, wp, cin, k, 1, padp, dil, residual=res, out=out)


def conv2d_wgrad(x: Act, dy: Act, conv: torch.nn.Conv2d, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """-> OIHW FP32 weight gradient, written to (accumulate: added to) `out` when given, else a new tensor."""
    lib = _lib.load()
    hdt = _HN_DTYPE[x.dtype]
    cout, cin, k = conv.out_channels, conv.in_channels, conv.kernel_size[0]
    cout_pad, kpad = lib.hn_conv_cout_pad(cout, hdt), lib.hn_conv_kpad(cin, k, k)
    if out is None:
        out, accumulate = torch.empty((cout, cin, k, k), dtype=torch.float32, device=x.buf.device), False
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == cout * cin * k * k
    cv = HnConv(cout, k, k, conv.stride[0], conv.padding[0], conv.dilation[0])
    xh, dh = x.hn(), dy.hn()
    ws_bytes = lib.hn_conv2d_wgrad_workspace_bytes(C.byref(xh), C.byref(cv))
    ws_ptr = workspace(ws_bytes, x.buf.device).data_ptr() if ws_bytes else None
    # 1x1 without padding: the packed [Cout][Cin] matrix IS the OIHW gradient -- the kernel accumulates straight into `out`
    direct = k == 1 and kpad == cin and cout_pad == cout
    packed = out if direct else torch.empty((cout_pad, kpad), dtype=torch.float32, device=x.buf.device)
    _lib.check(lib.hn_conv2d_wgrad(C.byref(xh), C.byref(dh), C.byref(cv), packed.data_ptr(), 0 if (direct and accumulate) else 1, ws_ptr,
                                   ws_bytes, _stream()))
    _count(2 + (1 if ws_bytes else 0))
    if not direct:
        _lib.check(lib.hn_unpack_wgrad(packed.data_ptr(), out.data_ptr(), cout, cin, k, k, kpad, int(accumulate), _stream()))
        _count()
    return out


def act_bwd(dout: Act, out: Act, act, slope=0.0) -> Act:
    dz = new_act(dout.n, dout.h, dout.w, dout.c, dout.dtype, dout.buf.device, ld=dout.ld if dout.c % 8 else None)
    _lib.check(_lib.load().hn_act_bwd(C.byref(dout.hn()), C.byref(out.hn()), act, float(slope), C.byref(dz.hn()), _stream()))
    _count()
    return dz


def channel_sums(x: Act) -> torch.Tensor:
    """FP64 [2, C]: per-channel sum and sum of squares over all pixels."""
    sums = torch.empty((2, x.c), dtype=torch.float64, device=x.buf.device)
    _lib.check(_lib.load().hn_channel_stats(C.byref(x.hn()), sums[0].data_ptr(), sums[1].data_ptr(), _stream()))
    _count(3)
    return sums


def vec_to_grad(src_f64: torch.Tensor, n: int, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    if out is None:
        out, accumulate = torch.empty((n,), dtype=torch.float32, device=src_f64.device), False
    _lib.check(_lib.load().hn_vec_to_grad(src_f64.data_ptr(), out.data_ptr(), n, int(accumulate), _stream()))
    _count()
    return out


def bn_bwd(dout: Act, out: Act, raw: Act, mean, invstd, gamma, act, slope=0.0, slope_ptr=None, dres: Optional[Act] = None,
           dres_accumulate=False, sinks=((None, False),) * 3, fwd_scale=None, fwd_shift=None, groups: int = 1) -> Act:
    """-> draw.  `sinks` = (tensor or None, accumulate) for dbeta, dgamma, dPReLU-slope: FP32 parameter gradients written by the
    apply kernel itself.  The kernel has ONE accumulate flag: sinks that disagree with it are routed through a scratch vector."""
    cch = dout.c
    dev = dout.buf.device
    sums = torch.empty(((2 * cch + 1) * groups,), dtype=torch.float64, device=dev)
    draw = new_act(dout.n, dout.h, dout.w, cch, dout.dtype, dev)
    wanted = [(t, a) for t, a in sinks if t is not None]
    acc = bool(wanted) and all(a for _, a in wanted)
    ptrs, fixups = [], []
    for t, a in sinks:
        if t is None:
            ptrs.append(None)
        elif a == acc:
            ptrs.append(t.data_ptr())
        else:                                  # mixed flags (cannot happen for one BN layer's own parameters in practice)
            tmp = torch.empty_like(t)
            ptrs.append(tmp.data_ptr())
            fixups.append((t, tmp, a))
    p = lambda t: t.detach().data_ptr() if t is not None else None
    _lib.check(_lib.load().hn_bn_bwd(C.byref(dout.hn()), C.byref(out.hn()), C.byref(raw.hn()), p(mean), p(invstd), p(gamma), act,
                                    float(slope), p(slope_ptr), sums.data_ptr(), C.byref(draw.hn()),
                                    C.byref(dres.hn()) if dres is not None else None, int(dres_accumulate),
                                    int(ptrs[2] is not None), ptrs[0], ptrs[1], ptrs[2], int(acc), p(fwd_scale), p(fwd_shift), int(groups), _stream()))
    _count(3)
    for t, tmp, a in fixups:
        t.add_(tmp) if a else t.copy_(tmp)
    return draw


def accumulate(x: Act, y: Act, add: bool):
    _lib.check(_lib.load().hn_accumulate(C.byref(x.hn()), C.byref(y.hn()), int(add), _stream()))
    _count()


def maxpool3x3s2_idx(x: Act, out: Optional[Act] = None):
    ho, wo = (x.h - 1) // 2 + 1, (x.w - 1) // 2 + 1
    out = out or new_act(x.n, ho, wo, x.c, x.dtype, x.buf.device)
    idx = torch.empty((x.n, ho, wo, x.c), dtype=torch.uint8, device=x.buf.device)
    _lib.check(_lib.load().hn_maxpool3x3s2_fwd_idx(C.byref(x.hn()), C.byref(out.hn()), idx.data_ptr(), _stream()))
    _count()
    return out, idx


def maxpool3x3s2_bwd(dy: Act, idx: torch.Tensor, dx: Act, add: bool):
    _lib.check(_lib.load().hn_maxpool3x3s2_bwd(C.byref(dy.hn()), idx.data_ptr(), C.byref(dx.hn()), int(add), _stream()))
    _count()


def bilinear_bwd(dy: Act, dx: Act, add: bool):
    lib = _lib.load()
    dyh, dxh = dy.hn(), dx.hn()
    ws_bytes = lib.hn_bilinear_bwd_workspace_bytes(C.byref(dyh), C.byref(dxh))
    ws_ptr = workspace(ws_bytes, dy.buf.device).data_ptr() if ws_bytes else None
    _lib.check(lib.hn_bilinear_bwd(C.byref(dyh), C.byref(dxh), int(add), ws_ptr, ws_bytes, _stream()))
    _count(2 if ws_bytes else 1)


def pyramid_pool_bwd(dpool: torch.Tensor, sizes: Sequence[int], dx: Act, add: bool):
    arr = (C.c_int32 * len(sizes))(*sizes)
    _lib.check(_lib.load().hn_pyramid_pool_bwd(dpool.data_ptr(), arr, len(sizes), C.byref(dx.hn()), int(add), _stream()))
    _count()


def bilinear_sum(xs: Sequence[Act], h: int, w: int) -> Act:
    """sum_i upsample(xs[i]) -> [N, h, w, C] in one pass (FP32 accumulation, one write)."""
    x0 = xs[0]
    out = new_act(x0.n, h, w, x0.c, x0.dtype, x0.buf.device)
    arr = (HnTensor * len(xs))(*[x.hn() for x in xs])
    with _timed("bilinear_sum"):
        _lib.check(_lib.load().hn_bilinear_sum_fwd(arr, len(xs), C.byref(out.hn()), _stream()))
    _count()
    return out


def upconv3x3(x: Act, conv: torch.nn.Conv2d, scale=None, shift=None, act=ACT_NONE, slope=0.0, slope_ptr=None, out_dtype=None,
              wp=None) -> Act:
    """act(conv3x3(bilinear_2x(x)) * scale + shift) in ONE kernel: the upsampled tensor is never written to HBM."""
    lib = _lib.load()
    assert conv.kernel_size == (3, 3) and conv.stride == (1, 1) and conv.padding == (1, 1) and conv.dilation == (1, 1)
    out = new_act(x.n, 2 * x.h, 2 * x.w, conv.out_channels, out_dtype or x.dtype, x.buf.device)
    wp = wp if wp is not None else packed_weight(conv, x.dtype)
    cv = HnConv(conv.out_channels, 3, 3, 1, 1, 1)
    ep = _epilogue(scale, shift, None, act, slope, slope_ptr)
    timing = conv_timer is not None
    if timing:
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_a.record()
    _lib.check(lib.hn_upconv3x3_fwd(C.byref(x.hn()), wp.data_ptr(), C.byref(cv), C.byref(ep), C.byref(out.hn()), _stream()))
    if timing:
        ev_b.record()
        conv_timer.append((f"{x.c}->{conv.out_channels} k3 s1 d1 @{2 * x.h}x{2 * x.w}", ev_a, ev_b))
    _count()
    return out


HEAD_FUSION = os.environ.get("HN_NO_HEAD_FUSION") is None


def conv3x3_head_ok(x: Act, conv: torch.nn.Conv2d, bn, head: torch.nn.Conv2d) -> bool:
    """conv3x3 (+ eval-mode BN) + activation + 1x1 classifier in one launch (hn_conv3x3_head_fwd)?"""
    return (HEAD_FUSION and current_tape is None and x.dtype == torch.bfloat16 and x.c % 64 == 0 and x.ld % 8 == 0
            and conv.kernel_size == (3, 3) and conv.stride == (1, 1) and conv.padding == (1, 1) and conv.dilation == (1, 1)
            and conv.out_channels == 64 and (bn is None or not (bn.training or bn.running_mean is None))
            and head.kernel_size == (1, 1) and head.stride == (1, 1) and head.padding == (0, 0) and head.in_channels == 64
            and head.out_channels <= 16)


def conv3x3_head(x: Act, conv: torch.nn.Conv2d, bn, head: torch.nn.Conv2d, act=ACT_NONE, slope=0.0, slope_ptr=None) -> torch.Tensor:
    """-> NCHW FP32 logits = head(act(bn(conv(x)))) without the intermediate activation in HBM (cm/models/pspnet.py:72-75, eval)."""
    lib = _lib.load()
    if bn is not None:
        wp, shift = packed_weight_folded(conv, bn, x.dtype)
    else:
        wp, (_, shift) = packed_weight(conv, x.dtype), folded_affine(conv, None)
    # the classifier travels as kernel parameters: host copies, refreshed when the parameters change (one D2H sync then)
    cache = head.__dict__.setdefault("_hn_wcache", {})
    ver = _versions(head.weight, head.bias)
    hit = cache.get("host")
    if hit is None or hit[0] != ver:
        hw = head.weight.detach().float().reshape(head.out_channels, 64).cpu().contiguous()
        hb = head.bias.detach().float().cpu().contiguous() if head.bias is not None else None
        hit = (ver, hw, hb)
        cache["host"] = hit
    hw, hb = hit[1], hit[2]
    out = torch.empty((x.n, head.out_channels, x.h, x.w), dtype=torch.float32, device=x.buf.device)
    cv = HnConv(64, 3, 3, 1, 1, 1)
    ep = _epilogue(None, shift, None, act, slope, slope_ptr)
    timing = conv_timer is not None
    if timing:
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_a.record()
    _lib.check(lib.hn_conv3x3_head_fwd(C.byref(x.hn()), wp.data_ptr(), C.byref(cv), C.byref(ep), hw.data_ptr(),
                                       hb.data_ptr() if hb is not None else None, head.out_channels, out.data_ptr(), _stream()))
    if timing:
        ev_b.record()
        conv_timer.append((f"{x.c}->64 k3 s1 d1 @{x.h}x{x.w} + head 64->{head.out_channels}", ev_a, ev_b))
    _count()
    return out


def upconv3x3_ok(x: Act, conv: torch.nn.Conv2d) -> bool:
    # the halo-patch producers are CUDA-core work: hidden behind the tensor pipe only when a tile has many k-blocks
    return (x.dtype == torch.bfloat16 and x.c % 64 == 0 and x.ld % 8 == 0 and conv.out_channels >= 33
            and x.c >= UPCONV_MIN_CIN)
