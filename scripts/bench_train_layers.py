"""Tuning aid: GPU time of the 1x1 convolutions of a training step (32 images of 320x640) with / without the fused BatchNorm statistics,
as the train-mode forward runs them (BF16 raw output, no activation).  Usage (GPU box): python scripts/bench_train_layers.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import torch
import torch.nn as nn

from heatnet_pub_b200 import engine as E
from bench_pair import timed

# name, cin, cout, k, pad, dil, n, h, w
LAYERS = [
    ("layer1 conv3 64->256 @80x160", 64, 256, 1, 0, 1, 32, 80, 160),
    ("layer1 conv1 256->64 @80x160", 256, 64, 1, 0, 1, 32, 80, 160),
    ("layer2 conv3 128->512 @40x80", 128, 512, 1, 0, 1, 32, 40, 80),
    ("layer3 conv3 256->1024 @40x80", 256, 1024, 1, 0, 1, 32, 40, 80),
    ("layer3 conv1 1024->256 @40x80", 1024, 256, 1, 0, 1, 32, 40, 80),
    ("layer4 conv3 512->2048 @40x80", 512, 2048, 1, 0, 1, 32, 40, 80),
    ("layer4 conv1 2048->512 @40x80", 2048, 512, 1, 0, 1, 32, 40, 80),
    ("layer3 conv2 3x3 d2 256->256 @40x80", 256, 256, 3, 2, 2, 32, 40, 80),
    ("layer4 conv2 3x3 d4 512->512 @40x80", 512, 512, 3, 4, 4, 32, 40, 80),
]

if __name__ == "__main__":
    torch.manual_seed(0)
    print("HN_NO_PAIR =", os.environ.get("HN_NO_PAIR"), "HN_PAIR_MIN_KB =", os.environ.get("HN_PAIR_MIN_KB"))
    for name, cin, cout, k, pad, dil, n, h, w in LAYERS:
        conv = nn.Conv2d(cin, cout, k, 1, pad, dil, bias=False).cuda()
        x = E.new_act(n, h, w, cin, torch.bfloat16, "cuda")
        x.buf.normal_()
        out = E.new_act(n, h, w, cout, torch.bfloat16, "cuda")
        sums = torch.zeros((2, 2 * cout), dtype=torch.float64, device="cuda")
        t_plain = timed(lambda: E.conv2d(x, conv, None, None, out=out))
        t_stats = timed(lambda: E.conv2d(x, conv, None, None, out=out, stats=sums, stat_groups=2))
        gb = (x.buf.numel() + out.buf.numel()) * 2 / 1e9
        tf = 2.0 * n * h * w * cout * cin * k * k / 1e9
        print(f"{name:40s} plain {t_plain * 1e3:7.1f} us ({tf / t_plain:6.0f} TFLOP/s, {gb / t_plain * 1e3:5.2f} TB/s)   with statistics {t_stats * 1e3:7.1f} us", flush=True)
