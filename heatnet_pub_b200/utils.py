"""Helpers the HeatNet wrapper needs from the reference's `models/confusion_maximization/utils.py`:
`weights_init_normal` (:126-132), `initModelRenamed/Partial/Full` (:59-90) and `calculate_ious` (:134-163,
here derived from the device confusion matrix instead of boolean-mask loops on the host)."""
import numpy as np
import torch

from . import iou_eval


def weights_init_normal(m):
    classname = m.__class__.__name__
    if classname.find('Conv') != -1:
        torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find('BatchNorm2d') != -1:
        torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
        torch.nn.init.constant_(m.bias.data, 0.0)


def initModelRenamed(model, weights_path, to_rename, rename):
    saved_model = torch.load(weights_path, map_location=lambda storage, loc: storage)
    if 'state_dict' in saved_model.keys():
        saved_model = saved_model['state_dict']
    model_dict = model.state_dict()
    weights_changed = {}
    for k, v in saved_model.items():
        k = k.replace(to_rename, rename)
        weights_changed[k] = v
    weights_changed = {k: v for k, v in weights_changed.items() if k in model_dict}
    print("Loaded dict with %d entries..." % len(weights_changed))
    assert (len(weights_changed) > 0)
    model_dict.update(weights_changed)
    model.load_state_dict(model_dict)


def initModelPartial(model, weights_path):
    model_dict = model.state_dict()
    pretrained_dict = torch.load(weights_path, map_location=lambda storage, loc: storage)['state_dict']
    pretrained_dict = {k: v for k, v in pretrained_dict.items() if k in model_dict}
    print('Updated : %d entries (initModelPartial)' % pretrained_dict.__len__())
    model_dict.update(pretrained_dict)
    model.load_state_dict(model_dict)


def initModelFull(model, weights_path):
    pretrained_dict = torch.load(weights_path, map_location=lambda storage, loc: storage)
    model.load_state_dict(pretrained_dict)


def calculate_ious(pred, target, n_classes=13):
    """Per-class IoU of utils.py:134-163 (classes 12 and 13 skipped; pixels whose target is 13 leave the
    union) computed from one 14 x 14 device histogram:
        inter = C[c,c];  union = sum_{t != 13} C[t,c] + sum_p C[c,p] - inter."""
    k = max(n_classes, 14)
    cm = iou_eval.ConfusionMatrix(k)
    cm.add(pred.reshape(-1), target.reshape(-1))
    conf = cm.conf.astype(np.int64)
    ious = []
    for cls in range(n_classes):
        if cls in (12, 13):
            continue
        inter = conf[cls, cls]
        union = conf[:13, cls].sum() + conf[13 + 1:, cls].sum() + conf[cls, :].sum() - inter
        ious.append(float('nan') if union == 0 else float(inter) / float(max(union, 1)))
    return np.array(ious)
