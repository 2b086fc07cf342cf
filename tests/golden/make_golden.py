"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference on the CPU.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py [--out tests/golden]

The reference ships no tests or golden vectors, so these files are what pins
oracle/ and, through it, the CUDA path.  Reference modules are imported with the three shims of
SURVEY.md section 8c (torchvision symbol, no .cuda() without a GPU, pretrained=False).  Weights come
from oracle.heatnet_oracle.recipe_fill applied to the reference modules' own state_dict, so any
implementation can rebuild them from the key names alone; BN running statistics produced by the
reference's train-mode forwards are stored in the fixture.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def load_reference():
    sys.path[:0] = [f"{REF}/models/confusion_maximization", f"{REF}/scripts"]
    import torchvision.models.resnet as _r
    _r.load_state_dict_from_url = torch.hub.load_state_dict_from_url      # shim 1 (critic_resnet.py:3)
    if not torch.cuda.is_available():
        nn.Module.cuda = lambda self, *a, **k: self                        # shim 2 (build_net.py:27)
    from models.pspnet import PSPNet
    from models import conf_segnet
    import discriminator_model
    import iou_eval
    import utils as cm_utils
    return PSPNet, conf_segnet, discriminator_model, iou_eval, cm_utils


def summarize(t):
    t = t.detach().double()
    return np.array([t.mean().item(), t.std().item(), t.abs().max().item(), t.sum().item()])


def sub(t, cstep=8):
    return t.detach()[:, ::cstep].contiguous().numpy()


def bn_stats(module):
    return {k: v.clone().numpy() for k, v in module.state_dict().items()
            if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    from oracle import heatnet_oracle as O
    PSPNet, conf_segnet, discriminator_model, iou_eval, cm_utils = load_reference()

    # ------------------------------------------------------------------ iou_eval
    rng = np.random.RandomState(1203412412 % (2 ** 31))
    K = 14
    out = {}
    pred = rng.randint(0, K, size=(6, 24, 40)).astype(np.int64)
    tgt = rng.randint(0, K, size=(6, 24, 40)).astype(np.int64)
    m = iou_eval.IoU(K, False, [12, 13])
    m.add(torch.from_numpy(pred[:3]), torch.from_numpy(tgt[:3]))
    out["conf_after_add1"] = m.conf_metric.conf.copy()
    iou1, miou1 = m.value()
    out["iou1"], out["miou1"] = iou1, np.array(miou1)
    out["conf_after_value1"] = m.conf_metric.conf.copy()          # ignore rows/cols zeroed in place
    m.add(torch.from_numpy(pred[3:]), torch.from_numpy(tgt[3:]))
    out["conf_after_add2"] = m.conf_metric.conf.copy()
    iou2, miou2 = m.value()
    out["iou2"], out["miou2"] = iou2, np.array(miou2)
    out["pred"], out["tgt"] = pred, tgt
    # scores path with ties (first max wins) and a class that never occurs (NaN IoU)
    scores = rng.randint(-3, 4, size=(2, K, 16, 20)).astype(np.float32)
    scores[:, 5] = -10.0
    tgt_s = rng.randint(0, K, size=(2, 16, 20)).astype(np.int64)
    tgt_s[tgt_s == 5] = 4
    m2 = iou_eval.IoU(K, False, None)
    m2.add(torch.from_numpy(scores), torch.from_numpy(tgt_s))
    out["scores"], out["tgt_s"] = scores, tgt_s
    out["conf_scores"] = m2.conf_metric.conf.copy()
    i3, mi3 = m2.value()
    out["iou3"], out["miou3"] = i3, np.array(mi3)
    m3 = iou_eval.IoU(K, True, 13)
    m3.add(torch.from_numpy(pred), torch.from_numpy(tgt))
    out["conf_normalized"] = m3.conf_metric.value().copy()
    # calculate_ious (cm/utils.py:134-163)
    out["calc_ious"] = cm_utils.calculate_ious(torch.from_numpy(pred), torch.from_numpy(tgt), 13)
    np.savez_compressed(os.path.join(args.out, "iou_golden.npz"), **out)
    print("iou_golden: miou1=%.6f miou2=%.6f miou3=%.6f" % (miou1, miou2, mi3))

    # ------------------------------------------------------------------ PSPNet late fusion, eval + train
    H, W = 64, 96
    net = PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50',
                 in_channels=4, pretrained=False, late_fusion=True)
    sd = net.state_dict()
    ref_keys = list(sd.keys())
    assert ref_keys == list(O.pspnet_state_dict(True, 4).keys()), "oracle key order differs from reference"
    O.recipe_fill(sd, seed=0)
    net.load_state_dict(sd)
    net.drop_1.p = 0.0           # parity runs disable Dropout2d (torch RNG cannot be matched), BN stays in train mode
    net.drop_2.p = 0.0
    rgb, ir = O.synthetic_inputs(2, H, W)
    net.train()
    with torch.no_grad():
        logits_tr, taps_tr, _ = net(rgb, ir)
    out = {"rgb": rgb.numpy(), "ir": ir.numpy(), "logits_train": logits_tr.numpy()}
    for i, t in enumerate(taps_tr[1:], 1):
        out[f"tap{i}_train_sub"] = sub(t)
        out[f"tap{i}_train_stats"] = summarize(t)
    for k, v in bn_stats(net).items():
        out["bn_after_train/" + k] = v
    net.eval()
    with torch.no_grad():
        logits_ev, taps_ev, _ = net(rgb[:1], ir[:1])
    out["logits_eval"] = logits_ev.numpy()
    for i, t in enumerate(taps_ev[1:], 1):
        out[f"tap{i}_eval_sub"] = sub(t)
        out[f"tap{i}_eval_stats"] = summarize(t)
    out["state_dict_keys"] = np.array(ref_keys)
    out["state_dict_shapes"] = np.array([str(tuple(v.shape)) for v in sd.values()])
    np.savez_compressed(os.path.join(args.out, "pspnet_late_golden.npz"), **out)
    print("pspnet late: logits_train std %.4f absmax %.4f | logits_eval std %.4f absmax %.4f | x5 eval std %.4f" % (
        logits_tr.std(), logits_tr.abs().max(), logits_ev.std(), logits_ev.abs().max(), taps_ev[1].std()))

    # ------------------------------------------------------------------ PSPNet early fusion (4 ch) + top-level signature
    net_e = PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50',
                   in_channels=4, pretrained=False, late_fusion=False)
    sd_e = net_e.state_dict()
    assert list(sd_e.keys()) == list(O.pspnet_state_dict(False, 4).keys())
    O.recipe_fill(sd_e, seed=1)
    net_e.load_state_dict(sd_e)
    net_e.eval()
    with torch.no_grad():
        le, te, _ = net_e(rgb[:1], ir[:1])
    out = {"logits_eval": le.numpy(), "state_dict_keys": np.array(list(sd_e.keys()))}
    for i, t in enumerate(te[1:], 1):
        out[f"tap{i}_eval_stats"] = summarize(t)
    np.savez_compressed(os.path.join(args.out, "pspnet_early_golden.npz"), **out)
    print("pspnet early: logits std %.4f" % le.std())

    # ------------------------------------------------------------------ FCDiscriminator
    crit = discriminator_model.FCDiscriminator(13)
    sd_c = crit.state_dict()
    assert list(sd_c.keys()) == list(O.critic_state_dict(13).keys())
    O.recipe_fill(sd_c, seed=2)
    crit.load_state_dict(sd_c)
    g = torch.Generator().manual_seed(7)
    xc = torch.randn(2, 13, 64, 96, generator=g)
    with torch.no_grad():
        yc = crit(xc)
    np.savez_compressed(os.path.join(args.out, "critic_golden.npz"), x=xc.numpy(), y=yc.numpy())
    print("critic: out", tuple(yc.shape), "std %.4f" % yc.std())

    # ------------------------------------------------------------------ conf_segnet: one train_critic and one train_seg step
    full = conf_segnet.conv_segnet(pretrained=False, disc_arch='cyclegan', num_critics=6, no_conf=False,
                                   modalities='ir_rgb', arch='pspnet', late_fusion=True)
    sd_f = full.state_dict()
    assert list(sd_f.keys()) == list(O.conf_segnet_state_dict(True, 6).keys())
    O.recipe_fill(sd_f, seed=3)
    full.load_state_dict(sd_f)
    full.trgb_segnet.drop_1.p = 0.0
    full.trgb_segnet.drop_2.p = 0.0
    full.train()
    HS = 256
    rgb_d, ir_d = O.synthetic_inputs(1, HS, HS, seed=11)
    rgb_n, ir_n = O.synthetic_inputs(1, HS, HS, seed=12)
    label = torch.randint(0, 13, (1, HS, HS), generator=torch.Generator().manual_seed(13))
    mse = nn.MSELoss()
    ce = nn.CrossEntropyLoss()
    out = {"label": label.numpy()}
    for phase in ("train_critic", "train_seg"):
        full.setPhase(phase)
        full.zero_grad()
        for p in full.parameters():
            p.grad = None
        o = full([rgb_d, ir_d], [rgb_n, ir_n])
        # cm/train_trgb_segnet_conf.py:437-446
        total_critics = sum(mse(c, torch.full_like(c, 1)) for c in o['critics_a']) + \
            sum(mse(c, torch.full_like(c, 0)) for c in o['critics_b'])
        if phase == "train_seg":
            seg_loss = ce(o['pred_label_a'], label)                       # :452
            conf = 0
            w0 = torch.ones_like(o['critics_a'][0])
            for c in o['critics_a']:                                      # :529-540
                conf = conf + torch.mean(torch.nn.functional.interpolate(w0, size=c.shape[2:], mode="bilinear") * mse(c, torch.full_like(c, 1)))
            for c in o['critics_b']:
                conf = conf + torch.mean(torch.nn.functional.interpolate(w0, size=c.shape[2:], mode="bilinear") * mse(c, torch.full_like(c, 1)))
            total = seg_loss + 0.1 * conf
            out["seg_loss"] = np.array(seg_loss.item())
            out["conf_loss"] = np.array(conf.item())
        else:
            total = total_critics
        total.backward()
        out[phase + "/total_critics"] = np.array(total_critics.item())
        out[phase + "/total"] = np.array(total.item())
        names, gn, gs = [], [], []
        for k, p in full.named_parameters():
            if p.grad is not None:
                names.append(k)
                gn.append(p.grad.double().norm().item())
                gs.append(p.grad.double().sum().item())
        out[phase + "/grad_names"] = np.array(names)
        out[phase + "/grad_norm"] = np.array(gn)
        out[phase + "/grad_sum"] = np.array(gs)
        out[phase + "/critic_out_shapes"] = np.array([str(tuple(c.shape)) for c in o['critics_a']])
        out[phase + "/logits_a_stats"] = summarize(o['pred_label_a'])
        print(phase, "total %.6f" % total.item(), "n grads", len(names))
    np.savez_compressed(os.path.join(args.out, "conf_segnet_golden.npz"), **out)


if __name__ == "__main__":
    main()
