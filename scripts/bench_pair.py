"""Tuning aid: per-layer time of the narrow-output convolutions with and without CTA pairs (cta_group::2).
Run twice on the GPU box and compare:   HN_NO_PAIR=1 python scripts/bench_pair.py ; python scripts/bench_pair.py
Every layer is also checked against torch's own convolution of the same BF16 operands (FP32 math on the GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn as nn
import torch.nn.functional as F

from heatnet_pub_b200 import engine as E

# name, cin, cout, k, stride, pad, dil, n, h, w, residual           (inference shapes of the 16 x 650 x 1920 workload)
LAYERS = [
    ("layer1.0 conv1 1x1 64->64", 64, 64, 1, 1, 0, 1, 16, 163, 480, False),
    ("layer1.x conv1 1x1 256->64", 256, 64, 1, 1, 0, 1, 16, 163, 480, False),
    ("layer1 conv2 3x3 64->64", 64, 64, 3, 1, 1, 1, 16, 163, 480, False),
    ("layer1 conv3 1x1 64->256 +res", 64, 256, 1, 1, 0, 1, 16, 163, 480, True),
    ("layer2.0 conv1 1x1 256->128", 256, 128, 1, 1, 0, 1, 16, 163, 480, False),
    ("layer2.0 conv2 3x3 s2 128->128", 128, 128, 3, 2, 1, 1, 16, 163, 480, False),
    ("layer2.x conv1 1x1 512->128", 512, 128, 1, 1, 0, 1, 16, 82, 240, False),
    ("layer2.x conv2 3x3 128->128", 128, 128, 3, 1, 1, 1, 16, 82, 240, False),
    ("layer2 conv3 1x1 128->512 +res", 128, 512, 1, 1, 0, 1, 16, 82, 240, True),
    ("up_3 class 3x3 64->64 full res (4 images)", 64, 64, 3, 1, 1, 1, 4, 650, 1920, False),
    ("up_2 3x3 256->64 half res (8 images)", 256, 64, 3, 1, 1, 1, 8, 325, 960, False),
    ("layer3 class 3x3 d2 256->256", 256, 256, 3, 1, 2, 2, 16, 82, 240, False),
    ("layer4 3x3 d4 512->512", 512, 512, 3, 1, 4, 4, 16, 82, 240, False),
    ("layer3 conv3 1x1 256->1024 +res", 256, 1024, 1, 1, 0, 1, 16, 82, 240, True),
    ("layer4 conv1 1x1 2048->512", 2048, 512, 1, 1, 0, 1, 16, 82, 240, False),
    ("layer3 conv1 1x1 1024->256", 1024, 256, 1, 1, 0, 1, 16, 82, 240, False),
    ("odd tile count 3x3 d2 64->64", 64, 64, 3, 1, 2, 2, 3, 37, 41, True),
]


def timed(fn, reps=10):
    """GPU time per call: `reps` calls captured in one CUDA graph (no host launch overhead between them), replayed 3 times"""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        g.replay()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / (3 * reps)


def head_layer():
    """up_3 + classifier in one kernel (hn_conv3x3_head_fwd) at 4 x 650 x 1920"""
    n, h, w = 4, 650, 1920
    conv = nn.Conv2d(64, 64, 3, 1, 1).cuda()
    bn = nn.BatchNorm2d(64).cuda().eval()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.1)
        bn.running_var.uniform_(0.5, 1.5)
        conv.weight.mul_(3.0)
    prelu = nn.PReLU().cuda()
    head = nn.Conv2d(64, 13, 1).cuda()
    x = E.new_act(n, h, w, 64, torch.bfloat16, "cuda")
    x.buf.normal_()
    fn = lambda: E.conv3x3_head(x, conv, bn, head, E.ACT_LEAKY, slope_ptr=prelu.weight)
    y = fn()
    torch.cuda.synchronize()
    with torch.no_grad():
        wf, shift = conv.weight * (bn.weight / torch.sqrt(bn.running_var + bn.eps))[:, None, None, None], None
        a = F.conv2d(x.nchw()[:1].float(), conv.weight.float(), conv.bias, 1, 1)
        a = F.prelu(bn(a), prelu.weight)
        ref = head(a.bfloat16().float())
    err = ((y[:1] - ref).abs().max() / ref.abs().max()).item()
    ms = timed(fn)
    tf = 2.0 * n * h * w * 64 * (64 * 9 + 13) / ms / 1e9
    print(f"{'up_3 + classifier head (4 images)':45s} {ms:8.3f} ms  {tf:7.1f} TFLOP/s   rel err {err:.2e} {'OK' if err < 3e-2 else 'MISMATCH'}", flush=True)


def main():
    torch.manual_seed(0)
    print("HN_NO_PAIR =", os.environ.get("HN_NO_PAIR"), " HN_NO_HALO =", os.environ.get("HN_NO_HALO"))
    for name, cin, cout, k, stride, pad, dil, n, h, w, with_res in LAYERS:
        conv = nn.Conv2d(cin, cout, k, stride, pad, dil, bias=True).cuda()
        with torch.no_grad():
            conv.weight.mul_(3.0)
        x = E.new_act(n, h, w, cin, torch.bfloat16, "cuda")
        x.buf.normal_()
        scale, shift = E.folded_affine(conv, None)
        ho = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
        wo = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
        res = None
        if with_res:
            res = E.new_act(n, ho, wo, cout, torch.bfloat16, "cuda")
            res.buf.normal_()
        fn = lambda: E.conv2d(x, conv, scale, shift, residual=res, act=E.ACT_RELU)
        y = fn()
        torch.cuda.synchronize()
        ref = F.conv2d(x.nchw().float(), conv.weight.detach().bfloat16().float(), conv.bias, stride, pad, dil)
        if res is not None:
            ref = ref + res.nchw().float()
        ref = F.relu(ref)
        err = ((y.nchw().float() - ref).abs().max() / ref.abs().max()).item()
        ms = timed(fn)
        tf = 2.0 * n * ho * wo * cout * cin * k * k / ms / 1e9
        print(f"{name:45s} {ms:8.3f} ms  {tf:7.1f} TFLOP/s   rel err {err:.2e} {'OK' if err < 1e-2 else 'MISMATCH'}", flush=True)
        del x, y, ref, res
    head_layer()


if __name__ == "__main__":
    main()
