"""CPU oracle for the `iou_eval` confusion-matrix / mIoU metric.  TEST INFRASTRUCTURE ONLY.

numpy restatement of /root/reference/scripts/iou_eval.py (ConfusionMatrix :19-101, IoU :103-182) and of
calculate_ious (models/confusion_maximization/utils.py:134-163).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.

Pinned against the reference itself by tests/golden/iou_golden.npz (tests/golden/make_golden.py).
The integer core (`bincount of pred + K*target`, int32 accumulator) is restated once more in plain C
(`hno_confusion` in oracle/heatnet_oracle.c).
"""
from __future__ import annotations

import numpy as np


def argmax_first(scores: np.ndarray, axis: int = 1) -> np.ndarray:
    """torch `max(1)` / np.argmax: ties resolve to the lowest index (iou_eval.py:154-157)."""
    return np.argmax(scores, axis=axis)


class ConfusionMatrixOracle:
    """iou_eval.py:29-101.  conf[t, p] += #{pred == p and target == t}; accumulator int32 (:32)."""

    def __init__(self, num_classes: int, normalized: bool = False):
        self.conf = np.zeros((num_classes, num_classes), dtype=np.int32)
        self.normalized = normalized
        self.num_classes = num_classes

    def reset(self):
        self.conf.fill(0)

    def add(self, predicted: np.ndarray, target: np.ndarray):
        k = self.num_classes
        predicted = np.asarray(predicted)
        target = np.asarray(target)
        assert predicted.shape[0] == target.shape[0], 'number of targets and predicted outputs do not match'
        if predicted.ndim != 1:
            assert predicted.shape[1] == k, 'number of predictions does not match size of confusion matrix'
            predicted = np.argmax(predicted, 1)
        else:
            assert predicted.max() < k and predicted.min() >= 0, 'predicted values are not between 0 and k-1'
        if target.ndim != 1:
            assert target.shape[1] == k, 'Onehot target does not match size of confusion matrix'
            assert (target >= 0).all() and (target <= 1).all(), 'in one-hot encoding, target values should be 0 or 1'
            assert (target.sum(1) == 1).all(), 'multi-label setting is not supported'
            target = np.argmax(target, 1)
        else:
            assert target.max() < k and target.min() >= 0, 'target values are not between 0 and k-1'
        x = predicted + k * target                                           # :82
        binc = np.bincount(x.astype(np.int32), minlength=k * k)             # :83-84
        assert binc.size == k * k
        self.conf += binc.reshape(k, k).astype(np.int32)                     # :86-88 (int32 wrap-around)

    def value(self):
        if self.normalized:
            conf = self.conf.astype(np.float32)
            return conf / conf.sum(1).clip(min=1e-12)[:, None]
        return self.conf                      # the accumulator itself, not a copy (:100-101)


class IoUOracle:
    """iou_eval.py:118-182 on numpy inputs ((N,H,W) labels or (N,K,H,W) scores)."""

    def __init__(self, num_classes: int, normalized: bool = False, ignore_index=None):
        self.conf_metric = ConfusionMatrixOracle(num_classes, normalized)
        if ignore_index is None:
            self.ignore_index = None
        elif isinstance(ignore_index, int):
            self.ignore_index = (ignore_index,)
        else:
            try:
                self.ignore_index = tuple(ignore_index)
            except TypeError:
                raise ValueError("'ignore_index' must be an int or iterable")

    def reset(self):
        self.conf_metric.reset()

    def add(self, predicted: np.ndarray, target: np.ndarray):
        predicted = np.asarray(predicted)
        target = np.asarray(target)
        assert predicted.shape[0] == target.shape[0], 'number of targets and predicted outputs do not match'
        assert predicted.ndim in (3, 4), "predictions must be of dimension (N, H, W) or (N, K, H, W)"
        assert target.ndim in (3, 4), "targets must be of dimension (N, H, W) or (N, K, H, W)"
        if predicted.ndim == 4:
            predicted = argmax_first(predicted, 1)
        if target.ndim == 4:
            target = argmax_first(target, 1)
        self.conf_metric.add(predicted.reshape(-1), target.reshape(-1))

    def value(self):
        conf_matrix = self.conf_metric.value()
        if self.ignore_index is not None:
            for _ in self.ignore_index:                       # quirk: whole tuple re-applied (:171-173)
                conf_matrix[:, self.ignore_index] = 0         # in place on the shared accumulator
                conf_matrix[self.ignore_index, :] = 0
        tp = np.diag(conf_matrix)
        fp = np.sum(conf_matrix, 0) - tp
        fn = np.sum(conf_matrix, 1) - tp
        with np.errstate(divide='ignore', invalid='ignore'):
            iou = tp / (tp + fp + fn)
        return iou, np.nanmean(iou)


def calculate_ious_oracle(pred: np.ndarray, target: np.ndarray, n_classes: int = 13) -> np.ndarray:
    """utils.calculate_ious (cm/utils.py:134-163): boolean-mask IoU, classes 12 (background) and 13
    (ignore) skipped, pixels with target == 13 excluded from the union."""
    pred = np.asarray(pred).reshape(-1)
    target = np.asarray(target).reshape(-1)
    ious = []
    keep = target != 13
    for cls in range(n_classes):
        if cls in (12, 13):
            continue
        p = pred == cls
        t = target == cls
        inter = int(np.sum(p[t]))
        union = int(np.sum(p[keep])) + int(np.sum(t[keep])) - inter
        ious.append(float('nan') if union == 0 else inter / max(union, 1))
    return np.array(ious)
