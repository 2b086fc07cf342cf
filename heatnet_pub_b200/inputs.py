"""Device-side input pipeline: what the HeatNet loaders do per pixel on the CPU before the H2D copy
(cm/thermal_loader.py:649-659 IR clip + range scaling, :715-728 F.to_tensor / F.normalize; cm/train_trgb_segnet_conf.py:82-86
rectDropTensor), done on the GPU after it.  The host ships uint8 RGB (N,H,W,3) and uint16 IR (N,H,W) frames -- 5 bytes per
pixel instead of 16 -- and gets back NCHW-shaped tensors that are channels-last views of the NHWC activations the network's stems
consume (zero-copy into PSPNet.forward / conv_segnet.forward).

The normalisations are evaluated once per possible input value with the reference's own arithmetic (torch FP32 for the RGB path,
numpy FP64 for the IR range scaling) into lookup tables; the kernels gather.  FP32 outputs are therefore bit-identical to the
reference's loader output."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import engine as E

IR_MINVAL, IR_MAXVAL = 21800, 25000            # cm/thermal_loader.py:650-651

_lut_cache = {}


def _rgb_lut(mean, std, device):
    key = ("rgb", tuple(mean), tuple(std), device)
    lut = _lut_cache.get(key)
    if lut is None:
        v = torch.arange(256, dtype=torch.uint8).view(256, 1, 1).expand(256, 1, 3).contiguous().numpy()       # HWC "image" of all byte values
        t = torch.from_numpy(v).permute(2, 0, 1).contiguous().to(torch.float32).div(255)                        # F.to_tensor
        m = torch.as_tensor(mean, dtype=torch.float32).view(-1, 1, 1)
        s = torch.as_tensor(std, dtype=torch.float32).view(-1, 1, 1)
        t = t.sub(m).div(s)                                                                                    # F.normalize
        lut = t.reshape(3, 256).contiguous().to(device)
        _lut_cache[key] = lut
    return lut


def _ir_lut(minval, maxval, mean, std, device):
    key = ("ir", minval, maxval, float(mean), float(std), device)
    lut = _lut_cache.get(key)
    if lut is None:
        k = np.arange(maxval - minval + 1, dtype=np.int64)
        t = k / (maxval - minval)                                        # numpy true division: float64 (thermal_loader.py:658)
        t = torch.from_numpy(t)                                          # F.to_tensor keeps float64
        t = t.sub(mean).div(std)                                         # F.normalize(mean=[0.5], std=[0.5])
        lut = t.to(torch.float32).contiguous().to(device)                # the trainer's .float() on the way to the model
        _lut_cache[key] = lut
    return lut


def _dtype(precision):
    return E.precision_dtype(precision or E.DEFAULT_PRECISION)


def prepare_rgb(rgb_u8: torch.Tensor, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), precision=None) -> torch.Tensor:
    """uint8 (N, H, W, 3) CUDA frames -> normalised (N, 3, H, W) tensor (channels-last view of an NHWC activation)."""
    _lib.require_device()
    assert rgb_u8.is_cuda and rgb_u8.dtype == torch.uint8 and rgb_u8.dim() == 4 and rgb_u8.shape[3] == 3, "rgb must be uint8 (N, H, W, 3) on the GPU"
    src = rgb_u8.contiguous()
    n, h, w, _ = src.shape
    out = E.new_act(n, h, w, 3, _dtype(precision), src.device)
    lut = _rgb_lut(mean, std, src.device)
    _lib.check(_lib.load().hn_prepare_rgb_u8(src.data_ptr(), lut.data_ptr(), C.byref(out.hn()), E._stream()))
    E._count()
    return out.nchw()


def prepare_ir(ir: torch.Tensor, minval=IR_MINVAL, maxval=IR_MAXVAL, mean=0.5, std=0.5, precision=None) -> torch.Tensor:
    """uint16 / int32 (N, H, W) CUDA thermal counts -> clipped, range-scaled, normalised (N, 1, H, W) tensor."""
    _lib.require_device()
    assert ir.is_cuda and ir.dim() == 3 and ir.dtype in (torch.uint16, torch.int16, torch.int32), "ir must be uint16 / int32 (N, H, W) on the GPU"
    assert ir.dtype != torch.int16 or maxval < 32768, "int16 storage cannot hold the requested range; use uint16 or int32"
    src = ir.contiguous()
    n, h, w = src.shape
    out = E.new_act(n, h, w, 1, _dtype(precision), src.device)
    lut = _ir_lut(int(minval), int(maxval), mean, std, src.device)
    bits = 32 if src.dtype == torch.int32 else 16          # non-negative int16 values read as uint16 are unchanged
    _lib.check(_lib.load().hn_prepare_ir(src.data_ptr(), bits, int(minval), int(maxval), lut.data_ptr(), C.byref(out.hn()), E._stream()))
    E._count()
    return out.nchw()


def rectDropTensor(tensor: torch.Tensor, params: torch.Tensor) -> torch.Tensor:
    """cm/train_trgb_segnet_conf.py:82-86: tensor[i, :, r0:r0+dh, c0:c0+dw] = 0 for every image, in place, one launch.
    Works on the views returned by prepare_rgb / prepare_ir and on plain contiguous NCHW FP32 tensors (via an NHWC round trip
    only if the tensor is not one of our views)."""
    _lib.require_device()
    act = E.act_from_view(tensor)
    p = params.to(device=tensor.device, dtype=torch.int32).contiguous()
    if act is None:                      # foreign NCHW tensor: same semantics with torch indexing, like the reference
        pi = p.cpu()
        for i in range(tensor.size(0)):
            tensor[i, :, pi[i, 0]:(pi[i, 0] + pi[i, 2]), pi[i, 1]:(pi[i, 1] + pi[i, 3])] = 0
        return tensor
    _lib.check(_lib.load().hn_rect_drop(C.byref(act.hn()), p.data_ptr(), E._stream()))
    E._count()
    return tensor


def _as_act(t: torch.Tensor):
    """NHWC view behind one of our channels-last tensors, or -- for a plain contiguous (N, 1, H, W) tensor, whose NCHW and NHWC
    layouts coincide -- a zero-copy reinterpretation."""
    act = E.act_from_view(t)
    if act is None and t.dim() == 4 and t.shape[1] == 1 and t.is_contiguous() and t.is_cuda and t.dtype in (torch.float32, torch.bfloat16):
        act = E.Act(t.view(t.shape[0], t.shape[2], t.shape[3], 1))
    return act


def ir_scale_aug(ir: torch.Tensor, scale: float) -> torch.Tensor:
    """cm/train_trgb_segnet_conf.py:404-406 `ir_day = scale * ir_day`, in place, one launch."""
    _lib.require_device()
    act = _as_act(ir)
    if act is None:
        return ir.mul_(scale)
    _lib.check(_lib.load().hn_label_scale(C.byref(act.hn()), None, None, 0, float(scale), None, E._stream()))
    E._count()
    return ir


def smartAugment(ir_day: torch.Tensor, label_day: torch.Tensor, rng=None) -> torch.Tensor:
    """cm/train_trgb_segnet_conf.py:101-110: every class present in `label_day` gets its own factor ~ U(0.1, 1.0) and the IR pixels
    of that class are multiplied by it.  The factors are drawn exactly like the reference draws them -- `random.uniform(0.1, 1.0)`
    once per entry of `torch.unique(label_day)`, in ascending label order -- so the same `random.seed()` gives the same augmentation;
    the per-class torch.where passes are one gather-multiply launch (in place)."""
    import random as _random
    _lib.require_device()
    rng = rng or _random
    labels = label_day.to(torch.int64).contiguous()
    present = torch.unique(labels).tolist()                      # ascending, like the reference's loop order
    k = (max(present) + 1) if present else 1
    if present and min(present) < 0:
        raise IndexError("smartAugment: negative label")
    factors = [1.0] * k
    for l in present:
        factors[int(l)] = rng.uniform(0.1, 1.0)
    act = _as_act(ir_day)
    if act is None:                                               # foreign layout: the reference's own loop
        for l in present:
            for b in range(ir_day.shape[0]):
                ir_day[b] = torch.where(label_day[b] == l, ir_day[b] * factors[int(l)], ir_day[b])
        return ir_day
    assert labels.shape == (act.n, act.h, act.w), "label map must be (N, H, W) matching the IR tensor"
    fdev = torch.tensor(factors, dtype=torch.float32, device=ir_day.device)
    _lib.check(_lib.load().hn_label_scale(C.byref(act.hn()), labels.data_ptr(), fdev.data_ptr(), k, 1.0, None, E._stream()))
    E._count()
    return ir_day
