"""ctypes loader for oracle/heatnet_oracle.c (plain-C primitives).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libheatnet_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_DIR, "heatnet_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-B", "libheatnet_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.hno_cross_entropy.restype = ctypes.c_double
    return _lib


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i64(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def conv2d(x, w, bias, stride, pad, dil):
    x = np.ascontiguousarray(x, np.float32); w = np.ascontiguousarray(w, np.float32)
    N, C, H, W = x.shape; K, _, R, S = w.shape
    Ho = (H + 2 * pad - dil * (R - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (S - 1) - 1) // stride + 1
    y = np.empty((N, K, Ho, Wo), np.float32)
    b = None if bias is None else _f(np.ascontiguousarray(bias, np.float32))
    lib().hno_conv2d(_f(x), _f(w), b, _f(y), N, C, H, W, K, R, S, stride, pad, dil)
    return y


def batchnorm(x, gamma, beta, rm, rv, training, momentum=0.1, eps=1e-5):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    N, C, H, W = x.shape
    lib().hno_batchnorm(_f(x), _f(y), _f(gamma), _f(beta), _f(rm), _f(rv), N, C, H, W, int(training),
                        ctypes.c_float(momentum), ctypes.c_float(eps))
    return y


def maxpool3x3s2(x):
    x = np.ascontiguousarray(x, np.float32)
    N, C, H, W = x.shape
    y = np.empty((N, C, (H - 1) // 2 + 1, (W - 1) // 2 + 1), np.float32)
    lib().hno_maxpool3x3s2(_f(x), _f(y), N, C, H, W)
    return y


def adaptive_avgpool(x, s):
    x = np.ascontiguousarray(x, np.float32)
    N, C, H, W = x.shape
    y = np.empty((N, C, s, s), np.float32)
    lib().hno_adaptive_avgpool(_f(x), _f(y), N, C, H, W, s)
    return y


def bilinear(x, Ho, Wo):
    x = np.ascontiguousarray(x, np.float32)
    N, C, H, W = x.shape
    y = np.empty((N, C, Ho, Wo), np.float32)
    lib().hno_bilinear(_f(x), _f(y), N, C, H, W, Ho, Wo)
    return y


def leaky(x, slope):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    lib().hno_leaky(_f(x), _f(y), ctypes.c_size_t(x.size), ctypes.c_float(slope))
    return y


def argmax(scores):
    scores = np.ascontiguousarray(scores, np.float32)
    N, K, H, W = scores.shape
    out = np.empty((N, H, W), np.int64)
    lib().hno_argmax(_f(scores), _i64(out), N, K, H, W)
    return out


def confusion(pred, target, K, conf=None):
    pred = np.ascontiguousarray(pred, np.int64).reshape(-1)
    target = np.ascontiguousarray(target, np.int64).reshape(-1)
    if conf is None:
        conf = np.zeros((K, K), np.int32)
    rc = lib().hno_confusion(_i64(pred), _i64(target), ctypes.c_size_t(pred.size), K,
                             conf.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    if rc != 0:
        raise AssertionError("values are not between 0 and k-1")
    return conf


def cross_entropy(logits, labels, ignore_index=-100):
    logits = np.ascontiguousarray(logits, np.float32)
    labels = np.ascontiguousarray(labels, np.int64)
    N, K, H, W = logits.shape
    return lib().hno_cross_entropy(_f(logits), _i64(labels), N, K, H, W, ignore_index)
