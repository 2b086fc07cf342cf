"""Tuning aid: where do the roles of conv_tc_kernel wait?  Builds a -DHN_PROFILE_ROLES copy of the library (the device
code needs -rdc for the extern counter array, so the profiling copy is a separate .so) and prints, per layer shape,
the average cycles per tile each role spent blocked.  Usage (GPU box): python scripts/profile_roles.py"""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PROF_LIB = os.path.join(ROOT, "heatnet_pub_b200", "libheatnet_b200_prof.so")


def build():
    import __graft_entry__ as ge
    src = sorted(glob.glob(os.path.join(ge.CSRC, "*.cu")))
    subprocess.check_call(["nvcc"] + ge.NVCC_FLAGS + ["-DHN_PROFILE_ROLES", "-rdc=true"] + src + ["-o", PROF_LIB])


if __name__ == "__main__":
    if not os.path.exists(PROF_LIB) or "--rebuild" in sys.argv:
        build()
    os.environ["HEATNET_B200_LIB"] = PROF_LIB
    os.environ["HN_NO_HALO"] = "1"
    import torch
    import torch.nn as nn
    from heatnet_pub_b200 import _lib, engine as E
    lib = _lib.load()
    names = ["prod_wait_empty", "prod_total", "mma_wait_full", "mma_wait_tempty", "mma_total", "epi_wait_tfull", "epi_wait_store_drain",
             "epi_wait_residual", "epi_total", "tiles", "ctas"]
    if os.environ.get("HN_EPI_DBG"):
        print("HN_EPI_DBG =", os.environ["HN_EPI_DBG"], "(timing experiment: results are NOT valid convolutions)")
    only = os.environ.get("HN_PROF_ONLY")
    shapes = [  # cin, cout, k, dil, n, h, w, residual
        (64, 64, 3, 1, 16, 656, 1920, False), (256, 64, 3, 1, 16, 328, 960, False), (64, 256, 1, 1, 16, 163, 480, True),
        (512, 2048, 1, 1, 16, 82, 240, True), (256, 1024, 1, 1, 16, 82, 240, True), (64, 64, 3, 1, 16, 163, 480, False),
        (512, 512, 3, 4, 16, 82, 240, False), (2048, 1024, 1, 1, 16, 82, 240, False)]
    if only:
        shapes = [shapes[int(i)] for i in only.split(",")]
    for (cin, cout, k, dil, n, h, w, res) in shapes:
        conv = nn.Conv2d(cin, cout, k, 1, dil * (k // 2), dil, bias=False).cuda()
        x = E.new_act(n, h, w, cin, torch.bfloat16, "cuda")
        x.buf.normal_()
        r = E.new_act(n, h, w, cout, torch.bfloat16, "cuda") if res else None
        scale = None; shift = torch.zeros(cout, device="cuda")
        buf = (C.c_ulonglong * 16)()
        for it in range(3):
            lib.hn_prof_read(buf, 1)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            E.conv2d(x, conv, scale, shift, residual=r, act=E.ACT_RELU)
            ev1.record()
            torch.cuda.synchronize()
        lib.hn_prof_read(buf, 1)
        v = list(buf)
        tiles, ctas = max(v[9], 1), max(v[10], 1)
        per_tile = {nme: v[i] / tiles for i, nme in enumerate(names[:9])}
        print(f"{cin}->{cout} k{k} d{dil} @{h}x{w} res={res}: {ev0.elapsed_time(ev1):.3f} ms, {tiles / ctas:.0f} tiles/CTA; cycles per tile: "
              + ", ".join(f"{k2}={val:.0f}" for k2, val in per_tile.items()))
