"""Launches the representative CTA-pair convolution layers once each inside a cudaProfiler range, for
`ncu --profile-from-start off --set full`.  Usage (GPU box): see scripts/gpu_ncu_pair.sh"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn as nn

from heatnet_pub_b200 import engine as E

# name, cin, cout, k, pad, dil, n, h, w
LAYERS = [
    ("up_2 3x3 256->64 (halo kernel, pair, N=64, streamed weights)", 256, 64, 3, 1, 1, 16, 325, 960),
    ("layer1 conv2 3x3 64->64 (halo kernel, pair, N=64, resident weights)", 64, 64, 3, 1, 1, 16, 163, 480),
    ("layer2 conv2 3x3 128->128 (halo kernel, pair, N=128)", 128, 128, 3, 1, 1, 16, 82, 240),
    ("layer3 conv2 3x3 d2 256->256 (generic kernel, pair, N=256)", 256, 256, 3, 2, 2, 16, 82, 240),
    ("layer4 conv2 3x3 d4 512->512 (generic kernel, pair, N=256)", 512, 512, 3, 4, 4, 16, 82, 240),
    ("layer4 conv1 1x1 2048->512 (generic kernel, pair, N=256, flat)", 2048, 512, 1, 0, 1, 16, 82, 240),
]

if __name__ == "__main__":
    torch.manual_seed(0)
    fns = []
    for name, cin, cout, k, pad, dil, n, h, w in LAYERS:
        conv = nn.Conv2d(cin, cout, k, 1, pad, dil, bias=True).cuda()
        x = E.new_act(n, h, w, cin, torch.bfloat16, "cuda")
        x.buf.normal_()
        scale, shift = E.folded_affine(conv, None)
        fns.append((name, (lambda x=x, conv=conv, scale=scale, shift=shift: E.conv2d(x, conv, scale, shift, act=E.ACT_RELU))))
    # the fused up_3 + classifier kernel
    conv = nn.Conv2d(64, 64, 3, 1, 1).cuda()
    bn = nn.BatchNorm2d(64).cuda().eval()
    prelu = nn.PReLU().cuda()
    head = nn.Conv2d(64, 13, 1).cuda()
    xh = E.new_act(16, 650, 1920, 64, torch.bfloat16, "cuda")
    xh.buf.normal_()
    fns.append(("up_3 + classifier 64->64->13 (halo kernel, pair, N=64, pair-issued head)",
                lambda: E.conv3x3_head(xh, conv, bn, head, E.ACT_LEAKY, slope_ptr=prelu.weight)))
    for _, fn in fns:
        for _ in range(2):
            fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for name, fn in fns:
        fn()
        print(name)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
