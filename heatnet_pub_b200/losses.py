"""Loss criteria of the training step on the B200 kernel library: forward and backward in one pass.

Drop-in `nn.Module`s for the criteria the reference trainers build
(`models/confusion_maximization/train_trgb_segnet_conf.py:238-245`: `BCEWithLogitsLoss()`, `MSELoss()`,
`CrossEntropyLoss()`; `scripts/main.py:223`: `CrossEntropyLoss(ignore_index=13)`), applied exactly where the
reference applies them (`:437-446`, `:452`, `:529-546`).  Only the configurations on that path are implemented:
mean reduction, no class weights.  The critic criteria accept the reference's `torch.full_like(c, 1)` target tensors
as well as a plain float (which saves materialising and reading the constant map).  Anything else raises.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from . import engine as E

KIND_MSE, KIND_BCE_LOGITS = 0, 1


def _scratch(device):
    lib = _lib.load()
    nbytes = lib.hn_loss_scratch_bytes()
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def _dense_f32(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        _lib.require_device()
        if logits.dim() != 4 or labels.dim() != 3 or labels.shape != (logits.shape[0], logits.shape[2], logits.shape[3]):
            raise ValueError(f"Expected logits (N,K,H,W) and labels (N,H,W); got {tuple(logits.shape)} and {tuple(labels.shape)}")
        if labels.dtype != torch.int64:
            raise RuntimeError("expected scalar type Long for the target")
        lib = _lib.load()
        x = _dense_f32(logits)
        lab = labels.contiguous()
        n, k, h, w = x.shape
        need_grad = ctx.needs_input_grad[0]
        dx = torch.empty_like(x) if need_grad else None
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        flags = torch.zeros(1, dtype=torch.int32, device=x.device)
        scratch, nbytes = _scratch(x.device)
        _lib.check(lib.hn_ce_loss_fwd_bwd(x.data_ptr(), lab.data_ptr(), n, k, h * w, int(ignore_index), 1.0, loss.data_ptr(),
                                          dx.data_ptr() if need_grad else None, flags.data_ptr(), scratch.data_ptr(), nbytes, E._stream()))
        E._count(4)
        ctx.dx, ctx.in_dtype = dx, logits.dtype
        _CrossEntropy.last_flags = flags
        return loss

    @staticmethod
    def backward(ctx, gout):
        dx = ctx.dx
        ctx.dx = None
        g = gout.detach().to(torch.float32).contiguous()
        _lib.check(_lib.load().hn_scale_by_scalar(dx.data_ptr(), dx.numel(), g.data_ptr(), E._stream()))
        E._count()
        return dx.to(ctx.in_dtype), None, None


class _CriticLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, kind):
        _lib.require_device()
        lib = _lib.load()
        xd = _dense_f32(x)
        tptr, tconst = None, 0.0
        if torch.is_tensor(target):
            if target.shape != x.shape:
                raise ValueError(f"Target size ({tuple(target.shape)}) must be the same as input size ({tuple(x.shape)})")
            if target.requires_grad:
                raise NotImplementedError("gradients w.r.t. the target are not on the hot path")
            tdense = _dense_f32(target)
            tptr = tdense.data_ptr()
        else:
            tconst = float(target)
        need_grad = ctx.needs_input_grad[0]
        dx = torch.empty_like(xd) if need_grad else None
        loss = torch.empty((), dtype=torch.float32, device=xd.device)
        scratch, nbytes = _scratch(xd.device)
        _lib.check(lib.hn_critic_loss_fwd_bwd(xd.data_ptr(), tptr, tconst, xd.numel(), int(kind), 1.0, loss.data_ptr(),
                                              dx.data_ptr() if need_grad else None, scratch.data_ptr(), nbytes, E._stream()))
        E._count(2)
        ctx.dx, ctx.in_dtype, ctx.shape = dx, x.dtype, x.shape
        return loss

    @staticmethod
    def backward(ctx, gout):
        dx = ctx.dx
        ctx.dx = None
        g = gout.detach().to(torch.float32).contiguous()
        _lib.check(_lib.load().hn_scale_by_scalar(dx.data_ptr(), dx.numel(), g.data_ptr(), E._stream()))
        E._count()
        return dx.view(ctx.shape).to(ctx.in_dtype), None, None


class CrossEntropyLoss(nn.Module):
    """`torch.nn.CrossEntropyLoss(ignore_index=...)` on (N,K,H,W) logits and (N,H,W) int64 labels (mean reduction)."""

    def __init__(self, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction='mean', label_smoothing=0.0):
        super().__init__()
        if weight is not None or reduction != 'mean' or label_smoothing != 0.0 or size_average is not None or reduce is not None:
            raise NotImplementedError("heatnet_pub_b200.losses.CrossEntropyLoss: only unweighted mean reduction is on the hot path")
        self.ignore_index = ignore_index

    def forward(self, input, target):
        loss = _CrossEntropy.apply(input, target, self.ignore_index)
        self.last_flags = _CrossEntropy.last_flags
        return loss

    def check_labels(self):
        """torch raises a device-side assert for labels outside [0,K) other than ignore_index; the kernel skips them and
        raises a flag instead.  This reads the flag of the last forward (one 4-byte D2H) and raises IndexError."""
        if getattr(self, "last_flags", None) is not None and int(self.last_flags.item()) & 1:
            raise IndexError("Target is out of bounds")



class MSELoss(nn.Module):
    """`torch.nn.MSELoss()` of a critic map; `criterion(c, 1.0)` == the reference's `criterion(c, torch.full_like(c, 1))`."""

    def __init__(self, size_average=None, reduce=None, reduction='mean'):
        super().__init__()
        if reduction != 'mean' or size_average is not None or reduce is not None:
            raise NotImplementedError("heatnet_pub_b200.losses.MSELoss: only mean reduction is on the hot path")

    def forward(self, input, target):
        return _CriticLoss.apply(input, target, KIND_MSE)


class BCEWithLogitsLoss(nn.Module):
    """`torch.nn.BCEWithLogitsLoss()` of a critic map against a target tensor or constant in [0, 1]."""

    def __init__(self, weight=None, size_average=None, reduce=None, reduction='mean', pos_weight=None):
        super().__init__()
        if weight is not None or pos_weight is not None or reduction != 'mean' or size_average is not None or reduce is not None:
            raise NotImplementedError("heatnet_pub_b200.losses.BCEWithLogitsLoss: only unweighted mean reduction is on the hot path")

    def forward(self, input, target):
        return _CriticLoss.apply(input, target, KIND_BCE_LOGITS)
