"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the per-pixel input arithmetic of the HeatNet loaders and trainer:
    cm/thermal_loader.py:649-659   IR clip to [21800, 25000] and (x - min) / (max - min)   (numpy, float64)
    cm/thermal_loader.py:715-728   F.to_tensor + F.normalize(mean .5, std .5) for RGB (FP32) and IR (FP64 -> .float())
    cm/train_trgb_segnet_conf.py:82-86   rectDropTensor
    cm/train_trgb_segnet_conf.py:101-110,404-406   smartAugment / ir_scale_aug
torchvision's to_tensor / normalize are restated in plain torch (uint8 HWC -> CHW float / 255; (t - mean) / std) and checked
against torchvision itself in tests/test_oracle.py when it is importable.  Parity pinned by that check and by construction
(the product path evaluates the same expressions once per input value); the reference ships no fixtures for its loaders."""
import numpy as np
import torch

IR_MINVAL, IR_MAXVAL = 21800, 25000


def to_tensor(pic: np.ndarray) -> torch.Tensor:
    """torchvision.transforms.functional.to_tensor for ndarrays: HWC -> CHW; uint8 is scaled by 1/255 in FP32."""
    if pic.ndim == 2:
        pic = pic[:, :, None]
    img = torch.from_numpy(np.ascontiguousarray(pic.transpose((2, 0, 1))))
    if img.dtype == torch.uint8:
        return img.to(torch.float32).div(255)
    return img


def normalize(t: torch.Tensor, mean, std) -> torch.Tensor:
    mean = torch.as_tensor(mean, dtype=t.dtype).view(-1, 1, 1)
    std = torch.as_tensor(std, dtype=t.dtype).view(-1, 1, 1)
    return t.clone().sub_(mean).div_(std)


def load_rgb(rgb_u8_hwc: np.ndarray, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> torch.Tensor:
    """thermal_loader.py:716-722: F.to_tensor(rgb) then F.normalize -> (3, H, W) FP32."""
    return normalize(to_tensor(rgb_u8_hwc), mean, std)


def load_ir(ir_counts_hw: np.ndarray, minval=IR_MINVAL, maxval=IR_MAXVAL) -> torch.Tensor:
    """thermal_loader.py:649-659,724-725 and the trainer's .float(): -> (1, H, W) FP32."""
    ir = ir_counts_hw.astype(np.int64).copy()
    ir[ir < minval] = minval
    ir[ir > maxval] = maxval
    ir = (ir - minval) / (maxval - minval)                     # float64
    t = normalize(to_tensor(ir), [0.5], [0.5])                  # float64
    return t.float()


def rect_drop_tensor(tensor: torch.Tensor, params: torch.Tensor) -> torch.Tensor:
    params = params.int()
    for i in range(tensor.size(0)):
        tensor[i, :, params[i, 0]:(params[i, 0] + params[i, 2]), params[i, 1]:(params[i, 1] + params[i, 3])] = 0
    return tensor


def smart_augment(ir_day: torch.Tensor, label_day: torch.Tensor, rng) -> torch.Tensor:
    """train_trgb_segnet_conf.py:101-110, with the module-level `random` replaced by the generator passed in."""
    label_indices = torch.unique(label_day)
    for l in label_indices:
        factor = rng.uniform(0.1, 1.0)
        for b in range(ir_day.shape[0]):
            ir_day[b] = torch.where(label_day[b] == l, ir_day[b] * factor, ir_day[b])
    return ir_day


def ir_scale_aug(ir_day: torch.Tensor, scale: float) -> torch.Tensor:
    """train_trgb_segnet_conf.py:404-406."""
    return scale * ir_day
