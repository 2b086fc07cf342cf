"""CPU oracle for the HeatNet dense-segmentation hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *restatement* of the reference's algorithm for the path named by
BASELINE.json -> north_star.  It is a checker: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  Nothing under
`heatnet_pub_b200/` imports it, and the product path fails loudly when its CUDA library is missing.

The reference executes every FLOP inside a third-party dependency that is not vendored under
/root/reference: PyTorch (ATen + oneDNN on CPU) -- the reference pins no version
(`environment.yml` is empty); this container has torch 2.11.0+cu128 / numpy 2.3.5.  The restatement
therefore states the network *graph* and its numerical semantics as plain functions over a flat
`state_dict` (key -> FP32 tensor), calling the same published torch.nn.functional primitives in FP32
on the CPU.  Each primitive is additionally restated in plain C in `oracle/heatnet_oracle.c`
and cross-checked in `tests/test_oracle.py`.

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the reference itself, produced in the build container by importing
/root/reference (script: tests/golden/make_golden.py, fixtures: tests/golden/*.npz).

All `file:line` citations are relative to /root/reference.  "cm/" abbreviates
models/confusion_maximization/.
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

BN_EPS = 1e-5       # nn.BatchNorm2d default, cm/models/extractors.py:72
BN_MOMENTUM = 0.1
RESNET50_LAYERS = (3, 4, 6, 3)   # cm/models/extractors.py:391
PSP_SIZES = (1, 2, 3, 6)         # cm/models/build_net.py:21
DROP_P1, DROP_P2 = 0.3, 0.15     # cm/models/pspnet.py:49,55


# ----------------------------------------------------------------------------------------------
# primitives (torch CPU FP32; C restatements of the same semantics live in heatnet_oracle.c)
# ----------------------------------------------------------------------------------------------
def conv2d(x: Tensor, sd: SD, name: str, stride: int = 1, padding: int = 0, dilation: int = 1) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride, padding, dilation)


def batchnorm(x: Tensor, sd: SD, name: str, training: bool) -> Tensor:
    """nn.BatchNorm2d: eps 1e-5, momentum 0.1; train mode normalises with the biased batch variance and
    updates running_var with the unbiased one, num_batches_tracked += 1."""
    if training:
        sd[name + ".num_batches_tracked"] += 1
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], training, BN_MOMENTUM, BN_EPS)


def upsample_bilinear(x: Tensor, size: Tuple[int, int]) -> Tensor:
    """F.upsample(mode='bilinear') == F.interpolate(align_corners=False) (cm/models/pspnet.py:23,39)."""
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


def dropout2d(x: Tensor, p: float, training: bool, mask: Optional[Tensor]) -> Tensor:
    """nn.Dropout2d: per-(b,c) channel mask scaled by 1/(1-p).  torch's RNG stream cannot be matched
    by another implementation, so parity runs inject `mask` (B,C) of {0,1} or disable dropout."""
    if not training or p == 0.0:
        return x
    if mask is None:
        return F.dropout2d(x, p, True)
    if p >= 1.0:                                   # torch: every channel dropped, zeros (not 0 * inf)
        return torch.zeros_like(x)
    return x * (mask.to(x.dtype) / (1.0 - p))[:, :, None, None]


# ----------------------------------------------------------------------------------------------
# encoder: cm/models/extractors.py:66-198
# ----------------------------------------------------------------------------------------------
def bottleneck(x: Tensor, sd: SD, p: str, stride: int, dilation: int, training: bool) -> Tensor:
    """Bottleneck.forward, cm/models/extractors.py:82-102.  1x1 -> BN -> ReLU -> 3x3(stride, dil,
    pad=dil) -> BN -> ReLU -> 1x1 -> BN -> (+ downsample(x) or x) -> ReLU."""
    out = F.relu(batchnorm(conv2d(x, sd, p + "conv1"), sd, p + "bn1", training))
    out = conv2d(out, sd, p + "conv2", stride=stride, padding=dilation, dilation=dilation)
    out = F.relu(batchnorm(out, sd, p + "bn2", training))
    out = batchnorm(conv2d(out, sd, p + "conv3"), sd, p + "bn3", training)
    if (p + "downsample.0.weight") in sd:
        residual = batchnorm(conv2d(x, sd, p + "downsample.0", stride=stride), sd, p + "downsample.1", training)
    else:
        residual = x
    return F.relu(out + residual)


def res_layer(x: Tensor, sd: SD, p: str, blocks: int, stride: int, dilation: int, training: bool) -> Tensor:
    """_make_layer, cm/models/extractors.py:155-170: block 0 takes the stride and (quirk) dilation 1,
    blocks 1.. take stride 1 and the layer's dilation."""
    x = bottleneck(x, sd, f"{p}.0.", stride, 1, training)
    for i in range(1, blocks):
        x = bottleneck(x, sd, f"{p}.{i}.", 1, dilation, training)
    return x


def stem(x: Tensor, sd: SD, conv: str, bn: str, training: bool) -> Tensor:
    x = F.relu(batchnorm(conv2d(x, sd, conv, stride=2, padding=3), sd, bn, training))
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=1)     # cm/models/extractors.py:128


def resnet_forward(sd: SD, modal_1: Tensor, modal_2: Optional[Tensor], late_fusion: bool, training: bool,
                   p: str = "feats.") -> List[Tensor]:
    """ResNet.forward, cm/models/extractors.py:172-198 -> [x5, x4, x3, x2, x1]."""
    L = RESNET50_LAYERS
    if late_fusion:
        x_1 = stem(modal_1, sd, p + "conv1", p + "bn1", training)
        # reference order of BN updates: stem, stem_2, layer1, layer1_2, layer2, layer2_2 (:173-190)
        x_1_ir = stem(modal_2, sd, p + "conv1_2", p + "bn1_2", training)
        x_2 = res_layer(x_1, sd, p + "layer1", L[0], 1, 1, training)
        x_2_ir = res_layer(x_1_ir, sd, p + "layer1_2", L[0], 1, 1, training)
        x_3 = res_layer(x_2, sd, p + "layer2", L[1], 2, 1, training)
        x_3_ir = res_layer(x_2_ir, sd, p + "layer2_2", L[1], 2, 1, training)
        x3c = torch.cat([x_3, x_3_ir], 1)
        x_4 = res_layer(x3c, sd, p + "layer3", L[2], 1, 2, training)
        x_5 = res_layer(x_4, sd, p + "layer4", L[3], 1, 4, training)
        return [x_5, x_4, x3c, torch.cat([x_2, x_2_ir], 1), torch.cat([x_1, x_1_ir], 1)]
    x = modal_1 if modal_2 is None else torch.cat([modal_1, modal_2], 1)
    x_1 = stem(x, sd, p + "conv1", p + "bn1", training)
    x_2 = res_layer(x_1, sd, p + "layer1", L[0], 1, 1, training)
    x_3 = res_layer(x_2, sd, p + "layer2", L[1], 2, 1, training)
    x_4 = res_layer(x_3, sd, p + "layer3", L[2], 1, 2, training)
    x_5 = res_layer(x_4, sd, p + "layer4", L[3], 1, 4, training)
    return [x_5, x_4, x_3, x_2, x_1]


# ----------------------------------------------------------------------------------------------
# PSP head + decoder: cm/models/pspnet.py:8-76
# ----------------------------------------------------------------------------------------------
def psp_module(feats: Tensor, sd: SD, p: str = "psp.") -> Tensor:
    """PSPModule.forward, cm/models/pspnet.py:21-25."""
    h, w = feats.shape[2:]
    priors = []
    for i, s in enumerate(PSP_SIZES):
        pooled = F.adaptive_avg_pool2d(feats, (s, s))
        priors.append(upsample_bilinear(conv2d(pooled, sd, f"{p}stages.{i}.1"), (h, w)))
    priors.append(feats)
    return F.relu(conv2d(torch.cat(priors, 1), sd, p + "bottleneck"))


def psp_upsample(x: Tensor, sd: SD, p: str, training: bool) -> Tensor:
    """PSPUpsample.forward, cm/models/pspnet.py:37-40: bilinear 2x -> 3x3 conv(+bias) -> BN -> PReLU."""
    x = upsample_bilinear(x, (2 * x.shape[2], 2 * x.shape[3]))
    x = batchnorm(conv2d(x, sd, p + "conv.0", padding=1), sd, p + "conv.1", training)
    return F.prelu(x, sd[p + "conv.2.weight"])


def pspnet_forward(sd: SD, modal_1: Tensor, modal_2: Optional[Tensor] = None, late_fusion: bool = True,
                   training: bool = False, dropout_masks: Optional[Sequence[Tensor]] = None,
                   dropout: bool = True):
    """PSPNet.forward, cm/models/pspnet.py:60-76 -> (logits, [logits, x5, x4, x3, x2, x1], None).

    `dropout_masks` = 4 keep-masks (B,1024),(B,256),(B,64),(B,64) for drop_1 and the three drop_2 calls;
    `dropout=False` disables Dropout2d while keeping BN in train mode (parity runs)."""
    f = resnet_forward(sd, modal_1, modal_2, late_fusion, training)
    dm = list(dropout_masks) if dropout_masks is not None else [None] * 4
    tr = training and dropout
    p = psp_module(f[0], sd)
    p = dropout2d(p, DROP_P1, tr, dm[0])
    p = psp_upsample(p, sd, "up_1.", training)
    p = dropout2d(p, DROP_P2, tr, dm[1])
    p = psp_upsample(p, sd, "up_2.", training)
    p = dropout2d(p, DROP_P2, tr, dm[2])
    p = psp_upsample(p, sd, "up_3.", training)
    p = dropout2d(p, DROP_P2, tr, dm[3])
    out = conv2d(p, sd, "final.0")
    return out, [out, f[0], f[1], f[2], f[3], f[4]], None


# ----------------------------------------------------------------------------------------------
# domain critics: cm/discriminator_model.py:35-64
# ----------------------------------------------------------------------------------------------
def fc_discriminator(x: Tensor, sd: SD, p: str) -> Tensor:
    """FCDiscriminator.forward: 5x conv 4x4 s2 p1 (+bias), LeakyReLU(0.2) between, bilinear x32."""
    for name in ("conv1", "conv2", "conv3", "conv4"):
        x = F.leaky_relu(conv2d(x, sd, p + name, stride=2, padding=1), 0.2)
    x = conv2d(x, sd, p + "classifier", stride=2, padding=1)
    return upsample_bilinear(x, (32 * x.shape[2], 32 * x.shape[3]))


def conf_segnet_forward(sd: SD, input_a: Sequence[Tensor], input_b: Sequence[Tensor], num_critics: int = 6,
                        late_fusion: bool = True, training: bool = True, dropout: bool = True,
                        dropout_masks_a=None, dropout_masks_b=None) -> dict:
    """conv_segnet.forward, cm/models/conf_segnet.py:106-140 (arch='pspnet', disc_arch='cyclegan',
    no feedback_seg / input_adapter).  Keys carry the 'trgb_segnet.' / 'critics.<i>.' prefixes of the
    reference module tree."""
    seg = _SubDict(sd, "trgb_segnet.")
    pred_a, inter_a, cert_a = pspnet_forward(seg, *input_a, late_fusion=late_fusion, training=training,
                                             dropout=dropout, dropout_masks=dropout_masks_a)
    pred_b, inter_b, cert_b = pspnet_forward(seg, *input_b, late_fusion=late_fusion, training=training,
                                             dropout=dropout, dropout_masks=dropout_masks_b)
    out = {"critics_a": [], "critics_b": []}
    for i in range(num_critics):
        out["critics_a"].append(fc_discriminator(inter_a[i], sd, f"critics.{i}."))
        out["critics_b"].append(fc_discriminator(inter_b[i], sd, f"critics.{i}."))
    out.update(pred_label_a=pred_a, pred_label_b=pred_b, cert_a=cert_a, cert_b=cert_b, inter_f_b=inter_b)
    return out


class _SubDict(dict):
    """View of a flat state dict under a key prefix that shares tensor storage (BN buffers update in place)."""

    def __init__(self, sd: SD, prefix: str):
        super().__init__({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)})


# ----------------------------------------------------------------------------------------------
# losses of the training step: cm/train_trgb_segnet_conf.py:237-245, 437-446, 452, 529-546
# ----------------------------------------------------------------------------------------------
def total_critics_loss(out: dict) -> Tensor:
    """train_critic objective (:437-446): sum_i MSE(c_a_i, 1) + sum_i MSE(c_b_i, 0), each its own mean."""
    la = sum(F.mse_loss(c, torch.ones_like(c)) for c in out["critics_a"])
    lb = sum(F.mse_loss(c, torch.zeros_like(c)) for c in out["critics_b"])
    return la + lb


def train_seg_loss(out: dict, label_day: Tensor, conf_weight: float = 0.1,
                   critic_weights: Optional[Sequence[float]] = None, multidir: bool = False):
    """train_seg objective (:452, 529-546): CE(pred_day, label) + conf_weight * confusion loss where both
    domains' critic maps are pushed to label 1 (day target 0 with --multidir).  The reference multiplies
    by bilinear-interpolated ones (identity) before the mean."""
    seg_loss = F.cross_entropy(out["pred_label_a"], label_day)
    n = len(out["critics_a"])
    w = list(critic_weights) if critic_weights is not None else [1.0] * n
    conf = 0.0
    for m, c in enumerate(out["critics_a"]):
        tgt = torch.zeros_like(c) if multidir else torch.ones_like(c)
        conf = conf + F.mse_loss(c, tgt) * w[m]
    for m, c in enumerate(out["critics_b"]):
        conf = conf + F.mse_loss(c, torch.ones_like(c)) * w[m]
    return seg_loss + conf_weight * conf, seg_loss, conf


# ----------------------------------------------------------------------------------------------
# state-dict construction (key set / shapes of the reference module tree) and the weight recipe
# ----------------------------------------------------------------------------------------------
def _bn_keys(sd, name, c):
    sd[name + ".weight"] = torch.ones(c)
    sd[name + ".bias"] = torch.zeros(c)
    sd[name + ".running_mean"] = torch.zeros(c)
    sd[name + ".running_var"] = torch.ones(c)
    sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)


def _layer_keys(sd, p, inplanes, planes, blocks, stride):
    """Key order follows nn.Module registration order in Bottleneck.__init__ (cm/models/extractors.py:69-80)."""
    for i in range(blocks):
        cin = inplanes if i == 0 else planes * 4
        b = f"{p}.{i}."
        sd[b + "conv1.weight"] = torch.zeros(planes, cin, 1, 1)
        _bn_keys(sd, b + "bn1", planes)
        sd[b + "conv2.weight"] = torch.zeros(planes, planes, 3, 3)
        _bn_keys(sd, b + "bn2", planes)
        sd[b + "conv3.weight"] = torch.zeros(planes * 4, planes, 1, 1)
        _bn_keys(sd, b + "bn3", planes * 4)
        if i == 0 and (stride != 1 or inplanes != planes * 4):
            sd[b + "downsample.0.weight"] = torch.zeros(planes * 4, inplanes, 1, 1)
            _bn_keys(sd, b + "downsample.1", planes * 4)
    return planes * 4


def pspnet_state_dict(late_fusion: bool = True, in_channels: int = 4, n_classes: int = 13) -> SD:
    """Zero-filled state dict with the reference PSPNet-ResNet50's keys, order and shapes
    (488 entries late fusion / 350 early)."""
    sd: SD = {}
    p = "feats."
    if late_fusion:
        sd[p + "conv1.weight"] = torch.zeros(64, 3, 7, 7); _bn_keys(sd, p + "bn1", 64)
        sd[p + "conv1_2.weight"] = torch.zeros(64, 1, 7, 7); _bn_keys(sd, p + "bn1_2", 64)
    else:
        sd[p + "conv1.weight"] = torch.zeros(64, in_channels, 7, 7); _bn_keys(sd, p + "bn1", 64)
    L = RESNET50_LAYERS
    c = _layer_keys(sd, p + "layer1", 64, 64, L[0], 1)
    if late_fusion:
        _layer_keys(sd, p + "layer1_2", 64, 64, L[0], 1)
    c2 = _layer_keys(sd, p + "layer2", c, 128, L[1], 2)
    if late_fusion:
        _layer_keys(sd, p + "layer2_2", c, 128, L[1], 2)
        c2 *= 2                                  # cm/models/extractors.py:142-143
    c3 = _layer_keys(sd, p + "layer3", c2, 256, L[2], 1)
    _layer_keys(sd, p + "layer4", c3, 512, L[3], 1)
    for i in range(len(PSP_SIZES)):
        sd[f"psp.stages.{i}.1.weight"] = torch.zeros(2048, 2048, 1, 1)
    sd["psp.bottleneck.weight"] = torch.zeros(1024, 2048 * 5, 1, 1)
    sd["psp.bottleneck.bias"] = torch.zeros(1024)
    for name, ci, co in (("up_1", 1024, 256), ("up_2", 256, 64), ("up_3", 64, 64)):
        sd[name + ".conv.0.weight"] = torch.zeros(co, ci, 3, 3)
        sd[name + ".conv.0.bias"] = torch.zeros(co)
        _bn_keys(sd, name + ".conv.1", co)
        sd[name + ".conv.2.weight"] = torch.full((1,), 0.25)
    sd["final.0.weight"] = torch.zeros(n_classes, 64, 1, 1)
    sd["final.0.bias"] = torch.zeros(n_classes)
    return sd


def critic_state_dict(num_classes: int, ndf: int = 64, prefix: str = "") -> SD:
    sd: SD = {}
    chans = [num_classes, ndf, ndf * 2, ndf * 4, ndf * 8, 1]
    for name, ci, co in zip(("conv1", "conv2", "conv3", "conv4", "classifier"), chans[:-1], chans[1:]):
        sd[f"{prefix}{name}.weight"] = torch.zeros(co, ci, 4, 4)
        sd[f"{prefix}{name}.bias"] = torch.zeros(co)
    return sd


CRITIC_CHANNELS_LATE = (13, 2048, 1024, 1024, 512, 128)    # cm/models/conf_segnet.py:46
CRITIC_CHANNELS_EARLY = (13, 2048, 1024, 512, 256, 64)     # cm/models/conf_segnet.py:49


def conf_segnet_state_dict(late_fusion: bool = True, num_critics: int = 6) -> SD:
    sd = {"trgb_segnet." + k: v for k, v in pspnet_state_dict(late_fusion, 4).items()}
    ch = CRITIC_CHANNELS_LATE if late_fusion else CRITIC_CHANNELS_EARLY
    for i in range(num_critics):
        sd.update(critic_state_dict(ch[i], prefix=f"critics.{i}."))
    return sd


def recipe_fill(sd: SD, seed: int = 0, residual_gain: float = 0.5) -> SD:
    """Deterministic, construction-order-independent weights: every tensor is drawn from its own
    generator seeded by crc32(key) ^ seed, so the same recipe applied to the reference module's
    state_dict (tests/golden/make_golden.py) and to any other implementation gives identical weights.

    conv weights ~ N(0, sqrt(2/(k*k*Cin))) (fan-in He init: keeps activations O(1) through all 85 convs
    and the critics; the extractor's own fan-out init, cm/models/extractors.py:148-151, does not);
    BN gamma ~ U(0.8,1.2) (x residual_gain on the block-closing bn3 so the eval-mode residual stream
    stays bounded), beta ~ N(0,0.05), running_mean ~ N(0,0.05), running_var ~ U(0.6,1.4); conv bias
    ~ N(0,0.05); PReLU slope 0.25 (+U(0,0.1))."""
    for k, v in sd.items():
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) ^ seed) & 0x7FFFFFFF)
        if k.endswith("num_batches_tracked"):
            v.zero_()
        elif v.dim() == 4:
            _, ci, kh, kw = v.shape
            v.copy_(torch.randn(v.shape, generator=g) * (2.0 / (kh * kw * ci)) ** 0.5)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) * 0.8 + 0.6)
        elif k.endswith("running_mean") or k.endswith(".bias"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.05)
        elif v.numel() == 1:                       # PReLU slope
            v.copy_(0.25 + 0.1 * torch.rand(v.shape, generator=g))
        else:                                      # BN gamma
            gain = residual_gain if ".bn3." in k else 1.0
            v.copy_((torch.rand(v.shape, generator=g) * 0.4 + 0.8) * gain)
    return sd


def synthetic_inputs(batch: int, height: int, width: int, seed: int = 1203412412):
    """RGB and IR in U(-1,1) as the loaders produce after (x-0.5)/0.5 (cm/thermal_loader.py:649-659);
    seed = the reference's own (scripts/main.py:126)."""
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(batch, 3, height, width, generator=g) * 2 - 1
    ir = torch.rand(batch, 1, height, width, generator=g) * 2 - 1
    return rgb, ir
