"""BF16 training parity with both numbers on the table: this implementation's deviation from the FP32 reference result
AND stock PyTorch's own BF16-autocast deviation on the same graph, inputs and weights (the oracle run on the GPU by
torch / cuDNN, TF32 off for the FP32 side).  Batch-statistic BatchNorm re-normalises 79 times per forward, which
amplifies BF16 rounding for ANY implementation, so the north star's 2e-2 is only reachable where stock torch reaches it
too; everywhere else the gate is "no worse than 1.25x stock torch's BF16 error" (SURVEY.md appendix D)."""
import os

import numpy as np
import pytest
import torch

from oracle import heatnet_oracle as O

pytestmark = pytest.mark.gpu


def _cuda_oracle_step(sd_cpu, phase, day, night, label, autocast):
    """One conf_segnet step of the oracle graph with torch autograd on the GPU -> (losses dict, {name: grad}, logits_a)."""
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        sd = {k: v.clone().cuda() for k, v in sd_cpu.items()}
        live = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k
                and (k.startswith("critics.") == (phase == "train_critic"))]
        for k in live:
            sd[k].requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = O.conf_segnet_forward(sd, day, night, training=True, dropout=False)
            out = {k: ([t.float() for t in v] if isinstance(v, list) else (v.float() if torch.is_tensor(v) else v)) for k, v in out.items()}
        if phase == "train_seg":
            total, seg_loss, conf = O.train_seg_loss(out, label)
            losses = {"total": total.item(), "seg_loss": seg_loss.item(), "conf_loss": conf.item()}
        else:
            total = O.total_critics_loss(out)
            losses = {"total": total.item()}
        total.backward()
        return losses, {k: sd[k].grad.float() for k in live if sd[k].grad is not None}, out["pred_label_a"].detach()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def _product_step(sd_cpu, phase, day, night, label, precision):
    import contextlib
    import io
    from heatnet_pub_b200 import conf_segnet, losses
    with contextlib.redirect_stdout(io.StringIO()):
        m = conf_segnet.conv_segnet(pretrained=False, disc_arch='cyclegan', num_critics=6, no_conf=False, modalities='ir_rgb',
                                    arch='pspnet', late_fusion=True)
        m.load_state_dict(sd_cpu)
        m = m.cuda().train()
        m.setPhase(phase)
    m.trgb_segnet.set_precision(precision)
    for c in m.critics:
        c.precision = precision
    m.trgb_segnet.drop_1.p = m.trgb_segnet.drop_2.p = 0.0
    mse, ce = losses.MSELoss(), losses.CrossEntropyLoss()
    o = m(list(day), list(night))
    if phase == "train_seg":
        seg_loss = ce(o['pred_label_a'], label)
        conf = sum(mse(c, 1.0) for c in o['critics_a']) + sum(mse(c, 1.0) for c in o['critics_b'])
        total = seg_loss + 0.1 * conf
        ls = {"total": total.item(), "seg_loss": seg_loss.item(), "conf_loss": conf.item()}
    else:
        total = sum(mse(c, 1.0) for c in o['critics_a']) + sum(mse(c, 0.0) for c in o['critics_b'])
        ls = {"total": total.item()}
    total.backward()
    return ls, {k: p.grad.float() for k, p in m.named_parameters() if p.grad is not None}, o['pred_label_a'].detach()


def _rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _summary(errs):
    v = np.sort(np.array(list(errs)))
    return {"median": float(np.median(v)), "p90": float(v[int(0.9 * (len(v) - 1))]), "max": float(v[-1])}


@pytest.mark.timeout(900)
@pytest.mark.parametrize("phase", ["train_critic", "train_seg"])
def test_bf16_conf_segnet_step_vs_reference_golden(golden_dir, phase):
    """The BF16 step against the REFERENCE's own FP32 gradients (tests/golden/conf_segnet_golden.npz: grad norm of every live
    tensor, losses), next to stock torch's BF16 autocast of the same step against the same golden."""
    g = np.load(os.path.join(golden_dir, "conf_segnet_golden.npz"))
    sd = O.recipe_fill(O.conf_segnet_state_dict(True, 6), seed=3)
    day = [t.cuda() for t in O.synthetic_inputs(1, 256, 256, seed=11)]
    night = [t.cuda() for t in O.synthetic_inputs(1, 256, 256, seed=12)]
    label = torch.from_numpy(g["label"]).cuda()
    names = [str(n) for n in g[phase + "/grad_names"]]
    want = g[phase + "/grad_norm"]
    ls, grads, _ = _product_step(sd, phase, day, night, label, "bf16")
    fl, fgrads, _ = _cuda_oracle_step(sd, phase, day, night, label, autocast=True)
    assert list(grads) == names
    gn = np.array([grads[k].double().norm().item() for k in names])
    fn = np.array([fgrads[k].double().norm().item() for k in names])
    scale = np.maximum(want, 1e-5 * want.max())       # conv biases in front of a BN: exactly-zero true gradient
    ours, floor = _summary(np.abs(gn - want) / scale), _summary(np.abs(fn - want) / scale)
    key = "total_critics" if phase == "train_critic" else "total"
    loss_err = abs(ls["total"] - float(g[phase + "/" + key])) / abs(float(g[phase + "/" + key]))
    floor_loss_err = abs(fl["total"] - float(g[phase + "/" + key])) / abs(float(g[phase + "/" + key]))
    print(f"[bf16 {phase} vs reference golden, {len(names)} tensors] grad-norm rel err ours {ours} | stock torch bf16 autocast {floor}; "
          f"loss rel err ours {loss_err:.3e} | torch {floor_loss_err:.3e}")
    assert loss_err < max(2e-2, 1.25 * floor_loss_err)
    for q in ("median", "p90"):
        assert ours[q] < max(2e-2, 1.25 * floor[q]), (q, ours, floor)
    # the worst tensor is one whose true gradient norm is ~0 (both BF16 implementations are > 100 % off there): bounded loosely
    assert ours["max"] < max(2e-2, 2.0 * floor["max"]), (ours, floor)


@pytest.mark.timeout(900)
def test_bf16_train_seg_step_config3_shape_vs_fp32_oracle():
    """BASELINE config 3 shape (320x640, 4 day+night pairs per GPU): train-mode logits, losses and every live gradient
    tensor of the BF16 step against the FP32 oracle (torch on the GPU, TF32 off); stock torch's BF16 autocast of the same
    step is measured against the same FP32 result.  A larger batch x map makes the batch statistics far better conditioned
    than the 64x96 / 256x256 fixtures, so this is the tightest train-mode bound on offer."""
    sd = O.recipe_fill(O.conf_segnet_state_dict(True, 6), seed=3)
    B, H, W = 4, 320, 640
    day = [t.cuda() for t in O.synthetic_inputs(B, H, W, seed=21)]
    night = [t.cuda() for t in O.synthetic_inputs(B, H, W, seed=22)]
    label = torch.randint(0, 13, (B, H, W), generator=torch.Generator().manual_seed(23)).cuda()
    rl, rgrads, rlogits = _cuda_oracle_step(sd, "train_seg", day, night, label, autocast=False)
    fl, fgrads, flogits = _cuda_oracle_step(sd, "train_seg", day, night, label, autocast=True)
    ls, grads, logits = _product_step(sd, "train_seg", day, night, label, "bf16")
    assert set(grads) == set(rgrads)
    rel = lambda a, b: ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
    lo, lf = rel(logits, rlogits), rel(flogits, rlogits)
    ours = _summary(_rel_l2(grads[k], rgrads[k]) for k in rgrads if rgrads[k].abs().max() > 0)
    floor = _summary(_rel_l2(fgrads[k], rgrads[k]) for k in rgrads if rgrads[k].abs().max() > 0)
    le = {k: abs(ls[k] - rl[k]) / abs(rl[k]) for k in rl}
    lfe = {k: abs(fl[k] - rl[k]) / abs(rl[k]) for k in rl}
    print(f"[bf16 train_seg 320x640 B=4 vs FP32 oracle] train-mode logits rel err ours {lo:.3e} | stock torch bf16 autocast {lf:.3e}; "
          f"per-tensor gradient rel-L2 err ours {ours} | torch {floor}; loss rel err ours {le} | torch {lfe}")
    assert lo < max(2e-2, 1.25 * lf)
    for k in rl:
        assert le[k] < max(2e-2, 1.25 * lfe[k]), k
    for q in ("median", "p90", "max"):
        assert ours[q] < max(3e-2, 1.25 * floor[q]), (q, ours, floor)
