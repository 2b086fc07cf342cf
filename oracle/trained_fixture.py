"""Separated-logits weight fixture for the argmax-agreement gate.  TEST INFRASTRUCTURE ONLY.

With random-init weights the top-2 logits of most pixels are near ties, so BF16 argmax agreement measures tie-breaking
noise, not kernel quality: torch's own BF16 autocast of the reference graph scores the same as any other BF16
implementation there.  The north star's ">= 99.9 % argmax agreement" is only meaningful on weights that actually
separate classes.  No checkpoint can be downloaded and a 217 MB state dict cannot be committed, so the fixture is
GENERATED IN THE JOB: the oracle graph (oracle/heatnet_oracle.py = the reference's graph, stock torch ops) is trained
for a few hundred Adam steps with stock torch on the GPU on a synthetic task whose labels are a function of the input --
class-coloured blocks: every class owns an RGB colour and an IR level, an image is a grid of blocks, the label map is
the block's class.  Nothing of the product (heatnet_pub_b200) is involved in making it.

The result is cached under /tmp for the lifetime of the box (tests, smoke() and bench.py's parity leg share it).
"""
from __future__ import annotations

import hashlib
import math
import os
import time
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import heatnet_oracle as O

N_CLASSES = 13
RECIPE = dict(steps=500, batch=4, height=160, width=320, block=80, lr=1e-3, noise=0.08, seed=20260)


def class_palette(device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """13 well-separated RGB colours in [-1, 1]^3 and 13 IR levels in [-1, 1] (class k -> colour k, level k)."""
    g = torch.Generator().manual_seed(13)
    levels = torch.tensor([-0.9, 0.0, 0.9])
    grid = torch.cartesian_prod(levels, levels, levels)                # 27 candidates, pairwise distance >= 0.9
    rgb = grid[torch.randperm(27, generator=g)[:N_CLASSES]]
    ir = torch.linspace(-0.9, 0.9, N_CLASSES)[torch.randperm(N_CLASSES, generator=g)]
    return rgb.to(device), ir.to(device)


def block_batch(batch: int, height: int, width: int, block: int, seed: int, noise: float = 0.08, device="cpu"):
    """-> rgb (B,3,H,W), ir (B,1,H,W) in [-1,1], label (B,H,W) int64: a grid of `block`-sized squares with a random
    class each (grid origin jittered per image so block edges fall anywhere), colours + uniform noise."""
    g = torch.Generator().manual_seed(seed)
    pal_rgb, pal_ir = class_palette()
    gh, gw = height // block + 2, width // block + 2
    cls = torch.randint(0, N_CLASSES, (batch, gh, gw), generator=g)
    oy = torch.randint(0, block, (batch,), generator=g)
    ox = torch.randint(0, block, (batch,), generator=g)
    ys = (torch.arange(height)[None, :] + oy[:, None]) // block          # (B, H)
    xs = (torch.arange(width)[None, :] + ox[:, None]) // block           # (B, W)
    label = cls[torch.arange(batch)[:, None, None], ys[:, :, None], xs[:, None, :]]
    rgb = pal_rgb[label].permute(0, 3, 1, 2) + (torch.rand(batch, 3, height, width, generator=g) * 2 - 1) * noise
    ir = pal_ir[label][:, None] + (torch.rand(batch, 1, height, width, generator=g) * 2 - 1) * noise
    return rgb.clamp(-1, 1).to(device), ir.clamp(-1, 1).to(device), label.to(device)


def _cache_path() -> str:
    key = hashlib.sha1(repr(sorted(RECIPE.items())).encode() + open(__file__, "rb").read()).hexdigest()[:16]
    return os.path.join(os.environ.get("HEATNET_FIXTURE_DIR", "/tmp"), f"heatnet_b200_separated_{key}.pt")


def train(device="cuda", log=None, **overrides) -> Dict[str, torch.Tensor]:
    """Train the oracle graph on class-coloured blocks with stock torch (FP32, TF32 allowed: this is fixture
    generation, not a parity measurement) -> CPU state dict incl. BN running statistics."""
    r = dict(RECIPE, **overrides)
    dev = torch.device(device)
    sd = {k: v.to(dev) for k, v in O.recipe_fill(O.pspnet_state_dict(True, 4), seed=0).items()}
    params = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    opt = torch.optim.Adam(params, lr=r["lr"])
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    t0 = time.time()
    try:
        for step in range(r["steps"]):
            for grp in opt.param_groups:              # cosine decay to 5 % of the base rate: sharpens the margins at the end
                grp["lr"] = r["lr"] * (0.05 + 0.95 * 0.5 * (1.0 + math.cos(math.pi * step / r["steps"])))
            rgb, ir, label = block_batch(r["batch"], r["height"], r["width"], r["block"], r["seed"] + step, r["noise"], dev)
            logits = O.pspnet_forward(sd, rgb, ir, late_fusion=True, training=True, dropout=False)[0]
            loss = F.cross_entropy(logits, label)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            if log and (step % 50 == 0 or step == r["steps"] - 1):
                acc = (logits.argmax(1) == label).float().mean().item()
                log(f"fixture step {step}: loss {loss.item():.4f}, train-mode pixel accuracy {acc:.4f}, {time.time() - t0:.1f} s")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    return {k: v.detach().cpu() for k, v in sd.items()}


def separated_state_dict(device="cuda", log=None) -> Dict[str, torch.Tensor]:
    """The cached fixture (trained on first use)."""
    path = _cache_path()
    if os.path.exists(path):
        try:
            return torch.load(path, map_location="cpu")
        except Exception:
            os.unlink(path)
    sd = train(device, log)
    tmp = path + f".{os.getpid()}.tmp"
    torch.save(sd, tmp)
    os.replace(tmp, path)
    return sd


def eval_frames(batch: int, height: int, width: int, block: int = 160, seed: int = 77, device="cpu"):
    """Evaluation frames of the same distribution at any size (the net is fully convolutional)."""
    return block_batch(batch, height, width, block, seed, RECIPE["noise"], device)


def agreement(a: torch.Tensor, b: torch.Tensor) -> float:
    return (a.argmax(1) == b.argmax(1)).float().mean().item()


def oracle_logits(sd: Dict[str, torch.Tensor], rgb: torch.Tensor, ir: torch.Tensor, autocast_bf16: bool = False) -> torch.Tensor:
    """Eval-mode oracle forward on the device of `rgb` (TF32 off): FP32 = the reference result, autocast = the noise floor
    stock PyTorch itself has in BF16."""
    dev = rgb.device
    s = {k: v.to(dev) for k, v in sd.items()}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad(), torch.autocast(dev.type, dtype=torch.bfloat16, enabled=autocast_bf16):
            out = O.pspnet_forward(s, rgb, ir, late_fusion=True, training=False)[0]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    return out.float()
