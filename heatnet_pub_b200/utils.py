"""Helpers the HeatNet wrapper needs from the reference's `models/confusion_maximization/utils.py`:
`weights_init_normal` (:126-132), `initModelRenamed/Partial/Full` (:59-90), `calculate_ious` (:134-163, here derived from the
device confusion matrix instead of boolean-mask loops on the host) and the validation loop of `validation_bdd_mf.py:259-379`
(`validate_step` / `validate_model`: batch duplication, BatchNorm in train mode, active Dropout2d, device argmax + histogram)."""
import numpy as np
import torch

from . import _lib
from . import engine as E
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


def ious_from_confusion(conf, n_classes=13):
    """utils.py:134-163 evaluated on a (>= 14 x 14) confusion matrix conf[target, pred]: classes 12 and 13 are skipped, pixels whose
    target is 13 leave the union:  inter = C[c,c];  union = sum_{t != 13} C[t,c] + sum_p C[c,p] - inter;  NaN for an empty union."""
    conf = np.asarray(conf).astype(np.int64)
    ious = []
    for cls in range(n_classes):
        if cls in (12, 13):
            continue
        inter = conf[cls, cls]
        union = conf[:13, cls].sum() + conf[13 + 1:, cls].sum() + conf[cls, :].sum() - inter
        ious.append(float('nan') if union == 0 else float(inter) / float(max(union, 1)))
    return np.array(ious)


def validate_step(model, inputs, label, conf=None):
    """One iteration of the reference's validate_model loop (cm/validation_bdd_mf.py:281-335) without the host round trips:
    every modality of the (batch-1) sample is duplicated along the batch axis (:297-299), the model runs under no_grad in
    WHATEVER mode it is in -- the reference never calls .eval() (:263), so BatchNorm uses (and updates) batch statistics and
    Dropout2d is active -- the first image's logits are arg-maxed on the device and counted against `label` into a 14 x 14
    device confusion matrix (hn_confusion with the fused first-max argmax) instead of `.cpu()` + torch.argmax + boolean-mask
    loops.  -> (segmented[0:1] logits, its int64 argmax map, the ConfusionMatrix accumulated into)."""
    from . import iou_eval
    if conf is None:
        conf = iou_eval.ConfusionMatrix(14)
    doubled = [torch.cat([t, t], dim=0) for t in inputs]
    with torch.no_grad():
        segmented, _, _ = model(*doubled)
    segmented = segmented[0:1, ...]
    scores = segmented.float().contiguous()
    n, k, h, w = scores.shape
    pred = torch.empty((n, h, w), dtype=torch.int64, device=scores.device)          # torch.argmax(segmented, 1): first maximum
    _lib.check(_lib.load().hn_argmax_labels(scores.data_ptr(), n, h * w, k, None, pred.data_ptr(), E._stream()))
    E._count()
    conf.add(pred.view(-1), label.to(device=scores.device, dtype=torch.long).reshape(-1))
    return segmented, pred, conf


def validate_model(model, val_loader, modalities, mode="day", vis=False, save_dir=""):
    """cm/validation_bdd_mf.py:259-379 minus visualisation / wandb logging (outside the hot path): -> per-class IoUs
    (12 entries: classes 12 and 13 are skipped), computed from ONE device confusion matrix accumulated over the loader."""
    from . import iou_eval
    print('Evaluating {}.'.format(mode))
    conf = iou_eval.ConfusionMatrix(14)
    for i, batch in enumerate(val_loader):
        print('Validating ... %d of %d ...' % (i, len(val_loader)))
        rgb_im, ir_im = batch['rgb'].cuda(), batch['ir'].cuda()
        label = batch['label'].cuda().to(torch.long)
        if 'rgb' in modalities and 'ir' in modalities:
            in_night = [rgb_im, ir_im]
        elif 'rgb' in modalities:
            in_night = [rgb_im]
        elif 'ir' in modalities:
            in_night = [ir_im]
        else:
            print('No known modality selected....')
            raise SystemExit
        validate_step(model, in_night, label, conf)
    return ious_from_confusion(conf.conf)


def calculate_ious(pred, target, n_classes=13):
    """Per-class IoU of utils.py:134-163 (classes 12 and 13 skipped; pixels whose target is 13 leave the
    union) computed from one 14 x 14 device histogram:
        inter = C[c,c];  union = sum_{t != 13} C[t,c] + sum_p C[c,p] - inter."""
    k = max(n_classes, 14)
    cm = iou_eval.ConfusionMatrix(k)
    cm.add(pred.reshape(-1), target.reshape(-1))
    return ious_from_confusion(cm.conf, n_classes)
