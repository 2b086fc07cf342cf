"""PSPNet pyramid-pooling head and PSPUpsample decoder on the B200 kernel library.

Host-side mirror of the reference's `models/pspnet.py` (top level, RGB, `forward(x) -> logits`) and
`models/confusion_maximization/models/pspnet.py` (HeatNet, `forward(modal_1, modal_2) ->
(logits, [logits, x5, x4, x3, x2, x1], None)`): same class names, constructor arguments, attribute
names and state_dict.  Reference: cm/models/pspnet.py:8-76, models/pspnet.py:8-75.
"""
import os

import torch
from torch import nn

from . import autograd
from . import engine as E
from . import extractors
from .engine import ACT_LEAKY, ACT_NONE, ACT_RELU, Act
from .extractors import _KernelModule


# BF16 training runs the PSP bottleneck in its projected form (PSPModule._run_projected_train); HN_NO_PSP_PROJECTED_TRAIN=1 keeps the
# literal 10240-channel concat formulation (the FP32 parity path always uses it)
PSP_PROJECTED_TRAIN = os.environ.get("HN_NO_PSP_PROJECTED_TRAIN") is None


class PSPModule(_KernelModule):
    """cm/models/pspnet.py:8-25.  One pass over the features produces all pyramid bins; each prior is
    upsampled straight into its channel slice of the (features*(len(sizes)+1))-channel buffer the encoder's
    last block also wrote into, so the 10240-channel torch.cat never happens; bias + ReLU sit in the
    bottleneck GEMM's epilogue."""

    def __init__(self, features, out_features=1024, sizes=(1, 2, 3, 6)):
        super().__init__()
        self.stages = []
        self.stages = nn.ModuleList([self._make_stage(features, size) for size in sizes])
        self.bottleneck = nn.Conv2d(features * (len(sizes) + 1), out_features, kernel_size=1)
        self.relu = nn.ReLU()

    def _make_stage(self, features, size):
        prior = nn.AdaptiveAvgPool2d(output_size=(size, size))
        conv = nn.Conv2d(features, features, kernel_size=1, bias=False)
        return nn.Sequential(prior, conv)

    def sizes(self):
        out = []
        for st in self.stages:
            s = st[0].output_size
            s = s if isinstance(s, int) else s[0]
            out.append(int(s))
        return out

    def alloc_cat(self, n, h, w, dtype, device) -> Act:
        return E.new_act(n, h, w, self.bottleneck.in_channels, dtype, device)

    def feats_slice(self, cat: Act) -> Act:
        f = self.stages[0][1].in_channels
        return cat.slice(len(self.stages) * f, f)

    def _run_cat(self, cat: Act) -> Act:
        """`cat` already holds the features in its last channel slice."""
        feats = self.feats_slice(cat)
        f = feats.c
        pooled = E.pyramid_pool(feats, self.sizes())
        for i, st in enumerate(self.stages):
            prior = E.conv_bn_act(pooled[i], st[1], None)
            E.bilinear(prior, cat.h, cat.w, out=cat.slice(i * f, f))
        return E.conv_bn_act(cat, self.bottleneck, None, ACT_RELU)

    def _run_projected(self, feats: Act) -> Act:
        """Same function without the 10240-channel tensor (inference; no tape): a 1x1 convolution commutes with
        bilinear upsampling, so
            bottleneck(cat(up(conv_s(pool_s f)), f)) = W_f f + b + sum_s up(W_s conv_s(pool_s f)),
        where W_s / W_f are the column blocks of the bottleneck weight.  The 10240->1024 GEMM shrinks to 2048->1024
        plus four GEMMs on <= 36 pixels per image, the four 2048-channel upsamples become one 1024-channel pass that
        the main GEMM's epilogue adds as its residual.  Summation order differs from the reference (FP32
        re-association only); throughput figures are still quoted against the reference's algorithmic FLOPs."""
        f = feats.c
        bott = self.bottleneck
        pooled = E.pyramid_pool(feats, self.sizes())
        projected = []
        for i, st in enumerate(self.stages):
            prior = E.conv2d(pooled[i], st[1])
            wp = E.packed_weight_slice(bott, i * f, (i + 1) * f, feats.dtype)
            projected.append(E.conv2d_raw(prior, wp, bott.out_channels, 1, 1, 0, 1))
        r = E.bilinear_sum(projected, feats.h, feats.w)
        scale, shift = E.folded_affine(bott, None)
        wf = E.packed_weight_slice(bott, len(self.stages) * f, (len(self.stages) + 1) * f, feats.dtype)
        return E.conv2d_raw(feats, wf, bott.out_channels, 1, 1, 0, 1, scale, shift, residual=r, act=ACT_RELU)

    def _run_projected_train(self, feats: Act) -> Act:
        """The projected form with a tape (BF16 training): the same five GEMMs as `_run_projected`, whose backward
        (engine.record_projected_bottleneck) needs a fifth of the FLOPs of the 10240-channel formulation in each of forward, dgrad
        and wgrad (2.15 -> 0.43 TFLOP per 32 images at 320x640) and never builds the 10240-channel tensor or its gradient."""
        f = feats.c
        bott = self.bottleneck
        pooled = E.pyramid_pool(feats, self.sizes())
        priors, projected = [], []
        for i, st in enumerate(self.stages):
            prior = E.conv_bn_act(pooled[i], st[1], None)
            priors.append(prior)
            wp = E.packed_weight_slice(bott, i * f, (i + 1) * f, feats.dtype)
            projected.append(E.conv2d_raw(prior, wp, bott.out_channels, 1, 1, 0, 1))
        r = E.bilinear_sum(projected, feats.h, feats.w)
        scale, shift = E.folded_affine(bott, None)
        wf = E.packed_weight_slice(bott, len(self.stages) * f, (len(self.stages) + 1) * f, feats.dtype)
        y = E.conv2d_raw(feats, wf, bott.out_channels, 1, 1, 0, 1, scale, shift, residual=r, act=ACT_RELU)
        E.record_projected_bottleneck(E.current_tape, feats, priors, y, bott)
        return y

    def projected_train_ok(self, dtype) -> bool:
        return (PSP_PROJECTED_TRAIN and dtype == torch.bfloat16 and self.bottleneck.kernel_size == (1, 1) and not self.bottleneck.padding[0]
                and self.stages[0][1].in_channels % 64 == 0)

    def _run(self, feats: Act) -> Act:
        if E.current_tape is None and not self.bottleneck.padding[0]:
            return self._run_projected(feats)
        if self.projected_train_ok(feats.dtype):
            return self._run_projected_train(feats)
        cat = self.alloc_cat(feats.n, feats.h, feats.w, feats.dtype, feats.buf.device)
        dst = self.feats_slice(cat)
        dst.buf[..., dst.coff:dst.coff + dst.c].copy_(feats.nchw().permute(0, 2, 3, 1))   # stand-alone use only
        return self._run_cat(cat)


class PSPUpsample(_KernelModule):
    """cm/models/pspnet.py:28-40: bilinear 2x -> 3x3 conv (+bias) -> BN -> PReLU."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, 3, padding=1),
            nn.BatchNorm2d(out_channels),
            nn.PReLU()
        )

    def _run(self, x: Act) -> Act:
        conv, bn, prelu = self.conv[0], self.conv[1], self.conv[2]
        bn_train = bn.training or bn.running_mean is None
        if E.current_tape is None and E.upconv3x3_ok(x, conv):
            # inference / validation: upsample fused into the conv's operand producer (no 2x tensor in HBM)
            if not bn_train:
                wp, shift = E.packed_weight_folded(conv, bn, x.dtype)
                return E.upconv3x3(x, conv, None, shift, ACT_LEAKY, slope_ptr=prelu.weight, wp=wp)
            scale, shift = E.folded_affine(conv, None)
            raw = E.upconv3x3(x, conv, scale, shift, out_dtype=torch.float32)
            bscale, bshift, _, _ = E.batchnorm_train_affine(raw, bn)
            return E.affine_act(raw, bscale, bshift, None, ACT_LEAKY, slope_ptr=prelu.weight,
                                out=E.new_act(raw.n, raw.h, raw.w, raw.c, x.dtype, x.buf.device))
        p = E.bilinear(x, 2 * x.h, 2 * x.w)
        return E.conv_bn_act(p, conv, bn, ACT_LEAKY, slope_ptr=prelu.weight)


class PSPNet(_KernelModule):
    """HeatNet PSPNet, cm/models/pspnet.py:43-76."""

    def __init__(self, n_classes=13, sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet34',
                 pretrained=True, late_fusion=False, in_channels=3):
        super().__init__()
        self.feats = getattr(extractors, backend)(pretrained, late_fusion=late_fusion, in_channels=in_channels)
        self.psp = PSPModule(psp_size, 1024, sizes)
        self.drop_1 = nn.Dropout2d(p=0.3)

        self.up_1 = PSPUpsample(1024, 256)
        self.up_2 = PSPUpsample(256, 64)
        self.up_3 = PSPUpsample(64, 64)

        self.drop_2 = nn.Dropout2d(p=0.15)
        self.final = nn.Sequential(
            nn.Conv2d(64, n_classes, kernel_size=1)
        )

    def set_precision(self, precision: str):
        """'bf16' (tcgen05 tensor-core convolutions) or 'fp32' (CUDA-core parity path, 1e-4 vs the reference)."""
        E.precision_dtype(precision)
        for m in self.modules():
            if isinstance(m, _KernelModule):
                m.precision = precision
        return self

    def _dropout(self, p: Act, drop: nn.Dropout2d) -> Act:
        if not (drop.training and drop.p > 0.0):
            return p
        masks = getattr(self, "_injected_dropout_masks", None)
        return E.dropout2d(p, drop.p, masks.pop(0) if masks else None)

    def _run_full(self, m1: Act, m2: Act = None):
        if E.current_tape is None:
            f = self.feats._run_taps(m1, m2)
            p = self.psp._run_projected(f[0])
        elif self.psp.projected_train_ok(m1.dtype):       # BF16 training: projected bottleneck with its own backward
            f = self.feats._run_taps(m1, m2)
            p = self.psp._run_projected_train(f[0])
        else:       # FP32 parity path: the literal concat formulation, whose backward the tape records op by op
            h8, w8 = self._h8w8(m1.h, m1.w)
            cat = self.psp.alloc_cat(m1.n, h8, w8, m1.dtype, m1.buf.device)
            f = self.feats._run_taps(m1, m2, x5_out=self.psp.feats_slice(cat))
            p = self.psp._run_cat(cat)
        p = self._dropout(p, self.drop_1)
        p = self.up_1._run(p)
        p = self._dropout(p, self.drop_2)
        p = self.up_2._run(p)
        p = self._dropout(p, self.drop_2)
        fin = self.final[0]
        up3 = self.up_3.conv
        drop_off = not (self.drop_2.training and self.drop_2.p > 0.0)
        if drop_off and not E.upconv3x3_ok(p, up3[0]):
            p = E.bilinear(p, 2 * p.h, 2 * p.w)
            if E.conv3x3_head_ok(p, up3[0], up3[1], fin):
                # inference: up_3's conv + BN + PReLU and the classifier in one kernel, logits written NCHW FP32 directly
                return E.conv3x3_head(p, up3[0], up3[1], fin, ACT_LEAKY, slope_ptr=up3[2].weight), f
            p = E.conv_bn_act(p, up3[0], up3[1], ACT_LEAKY, slope_ptr=up3[2].weight)
        else:
            p = self.up_3._run(p)
        p = self._dropout(p, self.drop_2)
        # logits are produced in FP32 on both paths (NHWC, channel stride padded to a multiple of 8)
        logits = E.new_act(p.n, p.h, p.w, fin.out_channels, torch.float32, p.buf.device, ld=(fin.out_channels + 7) // 8 * 8)
        E.conv_bn_act(p, fin, None, out=logits)
        return logits, f

    def _h8w8(self, h, w):
        h2, w2 = E.conv_out_hw(h, w, self.feats.conv1)
        h4, w4 = (h2 - 1) // 2 + 1, (w2 - 1) // 2 + 1
        st = self.feats.layer2[0].stride
        return (h4 - 1) // st + 1, (w4 - 1) // st + 1

    def forward(self, modal_1, modal_2=None):
        if autograd.needs_autograd(self, modal_1, modal_2):
            outs = autograd.apply(self._autograd_runner, [modal_1, modal_2], self)
            return outs[0], list(outs), None
        if getattr(self, "_graph_enabled", False) and self._graph_ok(modal_1, modal_2):
            return self._forward_graph(modal_1, modal_2)
        return self._forward_eager(modal_1, modal_2)

    def forward_pair(self, input_a, input_b):
        """`forward(*input_a)` and `forward(*input_b)` (the day and the night batch of an adversarial step,
        cm/models/conf_segnet.py:114-115) as ONE batch: every convolution, pooling and upsampling launch covers both domains
        (half as many launches, twice the tiles per launch, weights fetched once, one wgrad per layer instead of two accumulated
        ones), while every train-mode BatchNorm2d still normalises each domain with its own batch statistics and applies its
        two running-statistic updates in call order -- the results equal two consecutive calls up to floating-point
        summation order.  -> (logits, [logits, x5, x4, x3, x2, x1]) for the concatenated batch: rows [:B] are input_a's."""
        ins = [torch.cat([a, b], dim=0) for a, b in zip(input_a, input_b)]
        prev = E.bn_groups
        E.bn_groups = 2
        try:
            logits, taps, _ = self.forward(*ins)
        finally:
            E.bn_groups = prev
        return logits, taps

    def pair_ok(self, input_a, input_b) -> bool:
        """Can forward_pair replace two forward calls here?  Same shapes on the GPU; with BatchNorm2d in train mode the grouped
        statistics live on the BF16 engine and the flattened 1x1 convolutions need whole 32-pixel blocks per domain."""
        if len(input_a) != len(input_b) or any(a.shape != b.shape or not (a.is_cuda and b.is_cuda) for a, b in zip(input_a, input_b)):
            return False
        if not any(m.training for m in self.modules() if isinstance(m, nn.BatchNorm2d)):
            return True
        if (self.precision or E.DEFAULT_PRECISION) != "bf16" or E.BN_TRAIN_RAW_FP32 or not E.BN_FUSED_STATS:
            return False
        n, _, h, w = input_a[0].shape
        h2, w2 = E.conv_out_hw(h, w, self.feats.conv1)
        h4, w4 = (h2 - 1) // 2 + 1, (w2 - 1) // 2 + 1
        h8, w8 = self._h8w8(h, w)
        return (n * h4 * w4) % 32 == 0 and (n * h8 * w8) % 32 == 0

    def _forward_eager(self, modal_1, modal_2=None):
        m1, m2 = self.feats._inputs(modal_1, modal_2)
        logits, f = self._run_full(m1, m2)
        out = logits if torch.is_tensor(logits) else E.to_nchw_f32(logits)      # the reference's NCHW FP32 logits
        return out, [out, f[0].nchw(), f[1].nchw(), f[2].nchw(), f[3].nchw(), f[4].nchw()], None

    # ---- CUDA-graph replay of the inference forward -------------------------------------------------------------------
    # A forward is ~105 kernel launches issued from Python (~30 us each): below ~10 images of 320x640 per call the step is
    # launch-bound (3.2 ms whatever the batch; the kernels of a batch-1 forward take ~0.5 ms).  set_cuda_graph(True) captures
    # the launch sequence once per (input shapes, precision, parameter versions) and replays it; inputs are copied into static
    # buffers, the logits are returned as a fresh tensor, the feature taps are views of the graph's static buffers (valid
    # until the next call with the same shapes).  Eval mode, torch.no_grad() only; anything else takes the eager path.
    def set_cuda_graph(self, enabled: bool = True):
        self._graph_enabled = bool(enabled)
        self._graphs = {}
        return self

    def _graph_ok(self, modal_1, modal_2):
        if self.training or torch.is_grad_enabled() or E.conv_timer is not None or E.op_timer is not None:
            return False
        if any(m.training for m in self.modules() if isinstance(m, (nn.BatchNorm2d, nn.Dropout2d))):
            return False
        return modal_1.is_cuda and (modal_2 is None or modal_2.is_cuda)

    MAX_GRAPHS = 2            # live graphs per module (one per input shape): each pool holds a full set of activations

    def _graph_key(self, modal_1, modal_2):
        """-> (shape key, state version): a graph is reused for the same shapes while no parameter / buffer has changed."""
        ver = (sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers()), next(self.parameters()).data_ptr())
        return (tuple(modal_1.shape), modal_1.dtype, None if modal_2 is None else (tuple(modal_2.shape), modal_2.dtype),
                modal_1.device, self.precision or E.DEFAULT_PRECISION), ver

    def _forward_graph(self, modal_1, modal_2):
        key, ver = self._graph_key(modal_1, modal_2)
        entry = self._graphs.get(key)
        if entry is not None and entry[5] != ver:     # weights or BN buffers changed: the captured weight packs are stale
            del self._graphs[key]
            entry = None
        if entry is None:
            while len(self._graphs) >= self.MAX_GRAPHS:
                del self._graphs[next(iter(self._graphs))]      # oldest first (dicts keep insertion order)
            s1 = torch.empty_like(modal_1, memory_format=torch.contiguous_format)
            s2 = None if modal_2 is None else torch.empty_like(modal_2, memory_format=torch.contiguous_format)
            s1.copy_(modal_1)
            if s2 is not None:
                s2.copy_(modal_2)
            side = torch.cuda.Stream(device=modal_1.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):             # warm-up: weight packs, folded BN vectors, workspaces, smem attributes
                for _ in range(2):
                    self._forward_eager(s1, s2)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(modal_1.device)
            graph = torch.cuda.CUDAGraph()
            l0 = E.launch_count
            E.reset_stats_pool()
            ws_refs = E.begin_capture_refs()          # the scratch buffers the captured kernels point at live as long as the graph
            try:
                with torch.cuda.graph(graph):
                    out = self._forward_eager(s1, s2)
            finally:
                E.end_capture_refs()
            E.reset_stats_pool()
            entry = (graph, s1, s2, out, E.launch_count - l0, ver, ws_refs)
            self._graphs[key] = entry
        graph, s1, s2, out, launches = entry[:5]
        s1.copy_(modal_1)
        if s2 is not None:
            s2.copy_(modal_2)
        graph.replay()
        E.launch_count += launches
        logits = out[0].clone()
        return logits, [logits] + list(out[1][1:]), None

    def _autograd_runner(self, tape, inputs):
        """Forward under a tape: -> (input acts, output acts, output tensors) for autograd._NetFunction."""
        modal_1, modal_2 = inputs
        m1, m2 = self.feats._inputs(modal_1, modal_2)
        if modal_1.requires_grad:
            tape.require(m1)
        if m2 is not None and modal_2.requires_grad:
            tape.require(m2)
        logits, f = self._run_full(m1, m2)
        out = E.to_nchw_f32(logits)
        if self.feats.late_fusion or modal_2 is None:
            in_acts = [m1, m2]
        else:                                                   # early fusion: the two inputs are channel slices of m1
            c1 = modal_1.shape[1]
            in_acts = [m1.slice(0, c1), m1.slice(c1, m1.c - c1)]
        return in_acts, [logits] + list(f), [out] + [t.nchw() for t in f]


class PSPNetRGB(PSPNet):
    """Top-level models/pspnet.py:43-75 signature: `PSPNet(n_classes=18, ..., pretrained=True)`, `forward(x) -> logits`."""

    def __init__(self, n_classes=18, sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet34',
                 pretrained=True):
        super().__init__(n_classes, sizes, psp_size, deep_features_size, backend, pretrained, late_fusion=False, in_channels=3)

    def forward(self, x):
        return super().forward(x)[0]
