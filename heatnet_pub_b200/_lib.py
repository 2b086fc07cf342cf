"""ctypes binding of libheatnet_b200.so (the C ABI declared in include/heatnet_b200.h).

There is no CPU fallback: importing this module without the built library, or calling into it without an
sm_100 device, raises.  Build with `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
"""
import ctypes as C
import os

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HEATNET_B200_LIB", os.path.join(_DIR, "libheatnet_b200.so"))

HN_F32, HN_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2


class HnTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dtype", C.c_int32), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("c", C.c_int32), ("ld", C.c_int32)]


class HnEpilogue(C.Structure):
    _fields_ = [("scale", C.c_void_p), ("shift", C.c_void_p), ("residual", C.c_void_p), ("residual_ld", C.c_int32),
                ("act", C.c_int32), ("slope", C.c_float), ("slope_ptr", C.c_void_p), ("out_nchw", C.c_int32),
                ("stat_sum", C.c_void_p), ("stat_sqsum", C.c_void_p), ("per_image", C.c_int32), ("stat_groups", C.c_int32)]


class HnPackJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("cout", C.c_int32), ("cin", C.c_int32), ("r", C.c_int32), ("s", C.c_int32),
                ("rows_pad", C.c_int32), ("kpad", C.c_int32), ("kind", C.c_int32), ("dtype", C.c_int32), ("ty", C.c_int32), ("tx", C.c_int32),
                ("phiy", C.c_int32), ("phix", C.c_int32)]


class HnConv(C.Structure):
    _fields_ = [("cout", C.c_int32), ("r", C.c_int32), ("s", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("dil", C.c_int32)]


# name -> (restype, argtypes); every symbol declared in include/heatnet_b200.h
_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_T, _E, _CV = C.POINTER(HnTensor), C.POINTER(HnEpilogue), C.POINTER(HnConv)
SIGNATURES = {
    "hn_last_error": (C.c_char_p, []),
    "hn_version": (C.c_int, []),
    "hn_device_check": (C.c_int, []),
    "hn_prof_read": (C.c_int, [_P, C.c_int]),
    "hn_nchw_to_nhwc": (C.c_int, [_P, _T, _P]),
    "hn_nhwc_to_nchw": (C.c_int, [_T, _P, _P]),
    "hn_pack_weight": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "hn_pack_weight_scaled": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "hn_bn_fold": (C.c_int, [_P, _P, _P, _P, _P, _F, _P, _P, _I32, _P]),
    "hn_conv_cout_pad": (_I32, [_I32, _I32]),
    "hn_conv_kpad": (_I32, [_I32, _I32, _I32]),
    "hn_conv2d_workspace_bytes": (_I64, [_T, _CV]),
    "hn_conv2d_fwd": (C.c_int, [_T, _P, _CV, _E, _T, _P, _I64, _P]),
    "hn_upconv3x3_fwd": (C.c_int, [_T, _P, _CV, _E, _T, _P]),
    "hn_stem_pad": (C.c_int, [_T, _T, _P]),
    "hn_pack_stem_weight": (C.c_int, [_P, _P, _P, _I32, _I32, _P]),
    "hn_stem7x7s2_fwd": (C.c_int, [_T, _P, _I32, _E, _T, _P]),
    "hn_maxpool3x3s2_fwd": (C.c_int, [_T, _T, _P]),
    "hn_pyramid_pool_workspace_bytes": (_I64, [_T, C.POINTER(_I32), _I32]),
    "hn_pyramid_pool_fwd": (C.c_int, [_T, C.POINTER(_I32), _I32, _P, _P, _I64, _P]),
    "hn_bilinear_fwd": (C.c_int, [_T, _T, _P]),
    "hn_bilinear_sum_fwd": (C.c_int, [_T, _I32, _T, _P]),
    "hn_affine_act": (C.c_int, [_T, _E, _T, _P]),
    "hn_dropout2d_scale": (C.c_int, [_P, _I64, _F, _P, _P]),
    "hn_channel_stats": (C.c_int, [_T, _P, _P, _P]),
    "hn_bn_finalize": (C.c_int, [_P, _P, _I64, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _I32, _P]),
    "hn_bn_finalize_tracked": (C.c_int, [_P, _P, _I64, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _I32, _P]),
    "hn_bn_apply_train": (C.c_int, [_T, _P, _P, _I64, _P, _P, _F, _F, _P, _P, _P, _E, _T, _P, _P, _P, _P, _P]),
    "hn_act_bwd": (C.c_int, [_T, _T, _I32, _F, _T, _P]),
    "hn_bn_bwd": (C.c_int, [_T, _T, _T, _P, _P, _P, _I32, _F, _P, _P, _T, _T, _I32, _I32, _P, _P, _P, _I32, _P, _P, _I32, _P]),
    "hn_stem_pad_slack_bytes": (_I64, []),
    "hn_prepare_rgb_u8": (C.c_int, [_P, _P, _T, _P]),
    "hn_prepare_ir": (C.c_int, [_P, _I32, _I32, _I32, _P, _T, _P]),
    "hn_rect_drop": (C.c_int, [_T, _P, _P]),
    "hn_label_scale": (C.c_int, [_T, _P, _P, _I32, _F, _P, _P]),
    "hn_conv3x3_head_fwd": (C.c_int, [_T, _P, _CV, _E, _P, _P, _I32, _P, _P]),
    "hn_bn_batch_stats_scratch_bytes": (_I64, [_I32]),
    "hn_bn_batch_stats": (C.c_int, [_T, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
    "hn_accumulate": (C.c_int, [_T, _T, _I32, _P]),
    "hn_maxpool3x3s2_fwd_idx": (C.c_int, [_T, _T, _P, _P]),
    "hn_maxpool3x3s2_bwd": (C.c_int, [_T, _P, _T, _I32, _P]),
    "hn_bilinear_bwd_workspace_bytes": (_I64, [_T, _T]),
    "hn_bilinear_bwd": (C.c_int, [_T, _T, _I32, _P, _I64, _P]),
    "hn_pyramid_pool_bwd": (C.c_int, [_P, C.POINTER(_I32), _I32, _T, _I32, _P]),
    "hn_dilate": (C.c_int, [_T, _I32, _T, _P]),
    "hn_pack_weight_dgrad": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "hn_dgrad_s2_phase_kpad": (_I32, [_I32, _I32, _I32, _I32, _I32]),
    "hn_pack_weight_dgrad_phase": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "hn_conv2d_dgrad_s2_ok": (C.c_int, [_T, _CV, _T]),
    "hn_conv2d_dgrad_s2": (C.c_int, [_T, _P, _CV, _T, _I32, _P]),
    "hn_pack_chunk": (_I32, []),
    "hn_pack_job_init": (C.c_int, [C.POINTER(HnPackJob), _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32]),
    "hn_pack_weights_multi": (C.c_int, [_P, _P, _I32, _P]),
    "hn_conv2d_wgrad_workspace_bytes": (_I64, [_T, _CV]),
    "hn_conv2d_wgrad": (C.c_int, [_T, _T, _CV, _P, _I32, _P, _I64, _P]),
    "hn_unpack_wgrad": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "hn_vec_to_grad": (C.c_int, [_P, _P, _I32, _I32, _P]),
    "hn_confusion": (C.c_int, [_P, _P, _I64, _I64, _P, _I32, _P, _P, _P]),
    "hn_argmax_labels": (C.c_int, [_P, _I64, _I64, _I32, _P, _P, _P]),
    "hn_loss_scratch_bytes": (_I64, []),
    "hn_ce_loss_fwd_bwd": (C.c_int, [_P, _P, _I64, _I32, _I64, _I64, _F, _P, _P, _P, _P, _I64, _P]),
    "hn_critic_loss_fwd_bwd": (C.c_int, [_P, _P, _F, _I64, _I32, _F, _P, _P, _P, _I64, _P]),
    "hn_scale_by_scalar": (C.c_int, [_P, _I64, _P, _P]),
    "hn_optim_chunk": (_I32, []),
    "hn_grad_sqnorm": (C.c_int, [_P, _P, _I32, _P, _P, _P]),
    "hn_rmsprop_step": (C.c_int, [_P, _P, _I32, _F, _F, _F, _F, _F, _F, _F, _P, _P]),
    "hn_adam_step": (C.c_int, [_P, _P, _I32, _F, _F, _F, _F, _F, _I64, _F, _F, _P, _P]),
}

_lib = None


def load():
    """Load the shared library and bind every declared symbol.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the sm_100a CUDA library first "
                "(python -c 'import __graft_entry__ as g; g.build()').  heatnet_pub_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)        # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        if os.environ.get("HN_TIMELINE"):
            lib = _Timed(lib)
        _lib = lib
    return _lib


# ---- optional per-call device timeline (HN_TIMELINE=1): CUDA events around every enqueueing C-ABI call, used by
# bench.py --layer-table on the training workloads.  `timeline` is a list while recording, else None.
timeline = None
_NO_TIMING = {"hn_last_error", "hn_version", "hn_device_check", "hn_prof_read", "hn_conv_cout_pad", "hn_conv_kpad", "hn_optim_chunk",
              "hn_conv2d_workspace_bytes", "hn_conv2d_wgrad_workspace_bytes", "hn_pyramid_pool_workspace_bytes",
              "hn_bilinear_bwd_workspace_bytes", "hn_dgrad_s2_phase_kpad", "hn_conv2d_dgrad_s2_ok", "hn_pack_chunk", "hn_pack_job_init", "hn_loss_scratch_bytes", "hn_bn_batch_stats_scratch_bytes", "hn_stem_pad_slack_bytes"}


def _describe(name, args):
    try:
        if name in ("hn_conv2d_fwd", "hn_conv2d_wgrad", "hn_upconv3x3_fwd"):
            x, cv = args[0]._obj, args[2]._obj
            return f"{x.c}->{cv.cout} k{cv.r} s{cv.stride} d{cv.dil} @{x.h}x{x.w} n{x.n}"
        if name in ("hn_bilinear_bwd", "hn_bilinear_fwd"):
            a, b = args[0]._obj, args[1]._obj
            return f"{a.h}x{a.w}->{b.h}x{b.w} c{a.c}"
        if hasattr(args[0], "_obj") and isinstance(args[0]._obj, HnTensor):
            x = args[0]._obj
            return f"c{x.c} @{x.h}x{x.w} n{x.n} dt{x.dtype}"
    except Exception:
        pass
    return ""


class _Timed:
    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name in _NO_TIMING:
            return fn

        def call(*args):
            if timeline is None:
                return fn(*args)
            import torch
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = fn(*args)
            b.record()
            timeline.append((name, _describe(name, args), a, b))
            return rc

        setattr(self, name, call)
        return call


def check(rc: int):
    """Translate a nonzero C return code into RuntimeError(hn_last_error())."""
    if rc < 0:
        raise RuntimeError("libheatnet_b200: " + load().hn_last_error().decode())
    return rc


_device_ok = False


def require_device():
    """The kernels are sm_100a-only; fail loudly anywhere else."""
    global _device_ok
    if not _device_ok:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("heatnet_pub_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        check(load().hn_device_check())
        _device_ok = True
