"""Confusion matrix / IoU metric with the histogram on the device.

Drop-in for the reference's `scripts/iou_eval.py` (`Metric`, `ConfusionMatrix`, `IoU`): same constructor
arguments, `.add/.value/.reset`, same int32 `conf` accumulator (a numpy array that `IoU.value()` mutates in
place for the ignored classes, as the reference does), same assertion messages.  CUDA tensors never leave the
GPU: `hn_confusion` histograms `pred + K*target` (optionally fusing the class argmax) and only the K x K
matrix is copied back, instead of the reference's 16 bytes/pixel D2H + single-thread np.bincount
(iou_eval.py:53-88).  Host inputs (numpy arrays / CPU tensors) are uploaded once and take the same kernel.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _require(condition, message):
    """The reference signals bad inputs with `assert cond, message` (iou_eval.py:58-79,146-151): same exception, same text."""
    if not condition:
        raise AssertionError(message)


class Metric(object):
    """Interface of the reference's metrics (iou_eval.py:5-17): reset / add / value, all no-ops here."""

    def reset(self): pass

    def add(self): pass

    def value(self): pass


def _device():
    _lib.require_device()
    return torch.device("cuda", torch.cuda.current_device())


def _as_cuda(x, dtype=None):
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not x.is_cuda:
        x = x.to(_device(), non_blocking=True)
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x.contiguous()


class ConfusionMatrix(Metric):
    """iou_eval.py:19-101.  `conf[t, p]` counts pixels of target class t predicted as p."""

    def __init__(self, num_classes, normalized=False):
        Metric.__init__(self)
        self.num_classes, self.normalized = num_classes, normalized
        self.conf = np.zeros((num_classes, num_classes), dtype=np.int32)      # int32 like the reference's accumulator (iou_eval.py:32)

    def reset(self):
        self.conf[...] = 0

    def _histogram(self, pred_labels, scores, n_images, hw, target):
        k = self.num_classes
        dev = target.device
        out = torch.zeros(k * k + 1, dtype=torch.int64, device=dev)      # K*K counts + flag word
        flags = out[k * k:].view(torch.int32)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            rc = _lib.load().hn_confusion(pred_labels.data_ptr() if pred_labels is not None else None,
                                          scores.data_ptr() if scores is not None else None,
                                          n_images, hw, target.data_ptr(), k, out.data_ptr(), flags.data_ptr(), stream)
        _lib.check(rc)
        host = out.cpu().numpy()                                          # the only D2H: (K*K+1) * 8 bytes
        flag = int(host[k * k]) & 0xFFFFFFFF
        _require(not flag & 1, 'predicted values are not between 0 and k-1')
        _require(not flag & 2, 'target values are not between 0 and k-1')
        return host[:k * k].reshape(k, k)

    def add(self, predicted, target):
        """predicted: N labels or N x K scores; target: N labels or N x K one-hot (iou_eval.py:38-88)."""
        k = self.num_classes
        _require(predicted.shape[0] == target.shape[0], 'number of targets and predicted outputs do not match')
        predicted, target = _as_cuda(predicted), _as_cuda(target)
        scores = None
        if predicted.dim() == 1:
            predicted = predicted.to(torch.int64)
        else:
            _require(predicted.shape[1] == k, 'number of predictions does not match size of confusion matrix')
            # (N, K) scores: one "image" per row with hw = 1 is the same NCHW layout the kernel expects
            scores, predicted = predicted.float().contiguous(), None
        if target.dim() != 1:
            _require(target.shape[1] == k, 'Onehot target does not match size of confusion matrix')
            _require(bool(((target >= 0) & (target <= 1)).all()), 'in one-hot encoding, target values should be 0 or 1')
            _require(bool((target.sum(1) == 1).all()), 'multi-label setting is not supported')
            target = target.argmax(1)
        target = target.to(torch.int64).contiguous()
        n = target.shape[0]
        if n == 0:
            return
        if scores is not None:
            conf = self._histogram(None, scores, n, 1, target)
        else:
            conf = self._histogram(predicted.contiguous(), None, 1, n, target)
        self.conf += conf.astype(np.int32)                                # int32 accumulator, wraps like numpy

    def add_scores_nchw(self, scores, target):
        """(N, K, H, W) FP32 scores + (N, H, W) labels: fused first-max argmax + histogram, one pass."""
        k = self.num_classes
        scores = _as_cuda(scores, torch.float32)
        target = _as_cuda(target, torch.int64)
        _require(scores.shape[1] == k, 'number of predictions does not match size of confusion matrix')
        n, _, h, w = scores.shape
        if n * h * w == 0:
            return
        self.conf += self._histogram(None, scores, n, h * w, target.view(-1)).astype(np.int32)

    def value(self):
        """The accumulator itself (NOT a copy: IoU.value() edits it in place, as in the reference), or its row-normalised FP32
        version (iou_eval.py:90-101)."""
        if not self.normalized:
            return self.conf
        as_float = self.conf.astype(np.float32)
        row_sums = as_float.sum(1).clip(min=1e-12)
        return as_float / row_sums[:, None]


class IoU(Metric):
    """iou_eval.py:103-182."""

    def __init__(self, num_classes, normalized=False, ignore_index=None):
        Metric.__init__(self)
        self.conf_metric = ConfusionMatrix(num_classes, normalized)
        self.ignore_index = self._as_index_tuple(ignore_index)

    @staticmethod
    def _as_index_tuple(ignore_index):
        """None, one class id, or an iterable of class ids (iou_eval.py:122-133)."""
        if ignore_index is None:
            return None
        if isinstance(ignore_index, int):
            return (ignore_index,)
        try:
            return tuple(ignore_index)
        except TypeError:
            raise ValueError("'ignore_index' must be an int or iterable")

    def reset(self):
        self.conf_metric.reset()

    def add(self, predicted, target):
        """predicted: (N, K, H, W) scores or (N, H, W) labels; target likewise (iou_eval.py:135-159)."""
        _require(predicted.size(0) == target.size(0), 'number of targets and predicted outputs do not match')
        _require(predicted.dim() in (3, 4), "predictions must be of dimension (N, H, W) or (N, K, H, W)")
        _require(target.dim() in (3, 4), "targets must be of dimension (N, H, W) or (N, K, H, W)")

        if target.dim() == 4:
            _, target = _as_cuda(target).max(1)
        if predicted.dim() == 4:
            # fused max(1) + bincount; the reference's ConfusionMatrix sees labels here, so any K is accepted
            # for the score tensor only when it matches num_classes -- otherwise fall back to explicit labels
            if predicted.size(1) == self.conf_metric.num_classes:
                self.conf_metric.add_scores_nchw(predicted, target)
                return
            _, predicted = _as_cuda(predicted).max(1)
        self.conf_metric.add(predicted.reshape(-1), target.reshape(-1))

    def value(self):
        """-> (per-class IoU float64 array, mean IoU ignoring NaN) (iou_eval.py:161-182), including the
        reference's in-place zeroing of the ignored rows/columns of the shared accumulator."""
        cm = self.conf_metric.value()
        if self.ignore_index is not None:
            ignored = list(self.ignore_index)
            # the reference zeroes ALL ignored rows and columns on every pass of its loop, in the matrix value() returned -- for
            # an un-normalised metric that is the shared accumulator, so the counts are gone for later add() calls as well
            cm[:, ignored] = 0
            cm[ignored, :] = 0
        tp = np.diag(cm)
        fp = cm.sum(axis=0) - tp
        fn = cm.sum(axis=1) - tp
        with np.errstate(divide='ignore', invalid='ignore'):
            iou = tp / (tp + fp + fn)                 # 0 / 0 -> NaN for classes that never occur
        return iou, np.nanmean(iou)
