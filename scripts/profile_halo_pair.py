"""Tuning aid: where does the MMA-issuing warp of the halo-patch convolution kernel wait, with and without CTA pairs?
Needs the -DHN_PROFILE_ROLES build (scripts/profile_roles.py).  Usage (GPU box):
    HN_NO_PAIR=1 python scripts/profile_halo_pair.py ; python scripts/profile_halo_pair.py"""
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
    print("HN_NO_PAIR =", os.environ.get("HN_NO_PAIR"))
    shapes = [(64, 64, 4, 650, 1920), (256, 64, 8, 325, 960), (128, 128, 16, 82, 240), (64, 64, 16, 163, 480)]
    for (cin, cout, n, h, w) in shapes:
        conv = nn.Conv2d(cin, cout, 3, 1, 1, bias=False).cuda()
        x = E.new_act(n, h, w, cin, torch.bfloat16, "cuda")
        x.buf.normal_()
        shift = torch.zeros(cout, device="cuda")
        buf = (C.c_ulonglong * 16)()
        for it in range(3):
            lib.hn_prof_read(buf, 1)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            E.conv2d(x, conv, None, shift, act=E.ACT_RELU)
            ev1.record()
            torch.cuda.synchronize()
        lib.hn_prof_read(buf, 1)
        v = list(buf)
        tiles, ctas = max(v[9], 1), max(v[10], 1)
        mmas = 36 * (cin // 64)
        print(f"{cin}->{cout} 3x3 @{n}x{h}x{w}: {ev0.elapsed_time(ev1):.3f} ms, {tiles / ctas:.0f} tiles per issuing CTA; issuer cycles per tile: "
              f"total={v[4] / tiles:.0f} ({v[4] / tiles / mmas:.1f} per MMA), wait_patch={v[2] / tiles:.0f}, wait_weights={v[11] / tiles:.0f}, "
              f"wait_accumulator={v[3] / tiles:.0f}; epilogue warp 0 (sum over the CTAs that flush, per issuer tile): wait_tfull={v[5] / tiles:.0f} "
              f"store_drain={v[6] / tiles:.0f} tmem_ld={v[12] / tiles:.0f} fence+store={v[13] / tiles:.0f} release={v[14] / tiles:.0f} total={v[8] / tiles:.0f}")
