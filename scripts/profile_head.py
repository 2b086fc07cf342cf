"""Tuning aid for the fused conv3x3 + classifier kernel (hn_conv3x3_head_fwd): per-tile cycles of the epilogue role from
a -DHN_PROFILE_ROLES build (see profile_roles.py).  Usage (GPU box): python scripts/profile_head.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import profile_roles as PR

if __name__ == "__main__":
    if not os.path.exists(PR.PROF_LIB) or "--rebuild" in sys.argv:
        PR.build()
    os.environ["HEATNET_B200_LIB"] = PR.PROF_LIB
    import torch
    import torch.nn as nn
    from heatnet_pub_b200 import _lib, engine as E
    lib = _lib.load()
    n, h, w = 16, 656, 1920
    conv = nn.Conv2d(64, 64, 3, 1, 1).cuda()
    bn = nn.BatchNorm2d(64).cuda().eval()
    prelu = nn.PReLU().cuda()
    head = nn.Conv2d(64, 13, 1).cuda()
    x = E.new_act(n, h, w, 64, torch.bfloat16, "cuda")
    x.buf.normal_()
    tiles = n * ((h + 15) // 16) * ((w + 7) // 8)
    buf = (C.c_ulonglong * 16)()
    for name, fn in (("fused head", lambda: E.conv3x3_head(x, conv, bn, head, E.ACT_LEAKY, slope_ptr=prelu.weight)),
                     ("conv only", lambda: E.conv_bn_act(x, conv, bn, E.ACT_LEAKY, slope_ptr=prelu.weight))):
        for it in range(3):
            lib.hn_prof_read(buf, 1)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            fn()
            ev1.record()
            torch.cuda.synchronize()
        lib.hn_prof_read(buf, 1)
        v = list(buf)
        print(f"{name}: {ev0.elapsed_time(ev1):.3f} ms; per tile (cycles, epilogue warp 0 of each CTA): wait_tfull={v[5] / tiles:.0f} "
              f"wait_store_drain={v[6] / tiles:.0f} epi_total={v[8] / tiles:.0f}; tiles/CTA={tiles / 148:.0f}")
