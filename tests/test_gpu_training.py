"""Training path on a B200: forward + backward through torch.autograd against (a) the reference's own
gradients stored in tests/golden/conf_segnet_golden.npz (both phases of the adversarial step) and (b) autograd
of the CPU oracle on the same inputs and weights."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import heatnet_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _oracle_grads(sd, loss_fn):
    keys = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    for k in keys:
        sd[k].requires_grad_(True)
        sd[k].grad = None
    loss = loss_fn(sd)
    loss.backward()
    return loss.item(), {k: sd[k].grad for k in keys if sd[k].grad is not None}


@pytest.mark.parametrize("precision,gtol", [("fp32", 2e-3), ("bf16", None)])
def test_pspnet_backward_matches_oracle_autograd(precision, gtol):
    """Late-fusion PSPNet, train-mode BN, dropout off, loss = CE(logits) + mean-square of every tap: every
    parameter gradient against torch.autograd of the oracle."""
    from heatnet_pub_b200 import pspnet
    sd = O.recipe_fill(O.pspnet_state_dict(True, 4), seed=0)
    rgb, ir = O.synthetic_inputs(2, 64, 96)
    label = torch.randint(0, 13, (2, 64, 96), generator=torch.Generator().manual_seed(5))

    def loss_of(logits, taps):
        return F.cross_entropy(logits, label.to(logits.device)) + sum((t.float() ** 2).mean() for t in taps[1:])

    ref_loss, ref_g = _oracle_grads({k: v.clone() for k, v in sd.items()},
                                    lambda s: loss_of(*O.pspnet_forward(s, rgb, ir, late_fusion=True, training=True, dropout=False)[:2]))
    net = pspnet.PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50', in_channels=4,
                        pretrained=False, late_fusion=True)
    net.load_state_dict(sd)
    net = net.cuda().train().set_precision(precision)
    net.drop_1.p = net.drop_2.p = 0.0
    logits, taps, _ = net(rgb.cuda(), ir.cuda())
    assert logits.requires_grad
    loss = loss_of(logits, taps)
    loss.backward()
    got = {k: p.grad for k, p in net.named_parameters()}
    assert set(k for k, g in got.items() if g is not None) == set(ref_g)
    if precision == "fp32":
        assert abs(loss.item() - ref_loss) < 1e-4 * abs(ref_loss)
    worst = 0.0
    for k, g in ref_g.items():
        e = rel(got[k].cpu(), g)
        worst = max(worst, e)
        if gtol is not None:
            assert e < gtol, (k, e)
    print(f"[{precision}] loss {loss.item():.6f} (oracle {ref_loss:.6f}); worst per-tensor gradient rel err {worst:.3e}")
    if gtol is None:
        # BF16 + batch-stat BN: bounded by the oracle's own BF16-autocast gradient noise
        sd2 = {k: v.clone() for k, v in sd.items()}
        with torch.autocast("cpu", dtype=torch.bfloat16):
            _, bf_g = _oracle_grads(sd2, lambda s: loss_of(*O.pspnet_forward(s, rgb, ir, late_fusion=True, training=True, dropout=False)[:2]))
        floor = max(rel(bf_g[k].float(), g) for k, g in ref_g.items())
        print(f"[bf16] oracle BF16-autocast worst gradient rel err {floor:.3e}")
        assert worst < max(1.5 * floor, 0.1)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 5e-2)])
def test_critic_backward(precision, tol):
    from heatnet_pub_b200 import discriminator_model
    sd = O.recipe_fill(O.critic_state_dict(64), seed=2)
    x = torch.randn(2, 64, 64, 96, generator=torch.Generator().manual_seed(1))
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    yr = O.fc_discriminator(xr, ref_sd, "")
    F.mse_loss(yr, torch.ones_like(yr)).backward()
    crit = discriminator_model.FCDiscriminator(64)
    crit.load_state_dict(sd)
    crit = crit.cuda()
    crit.precision = precision
    xg = x.cuda().requires_grad_(True)
    y = crit(xg)
    assert y.shape == yr.shape and y.requires_grad
    F.mse_loss(y, torch.ones_like(y)).backward()
    assert rel(y.detach().cpu(), yr.detach()) < (1e-4 if precision == "fp32" else 2e-2)
    for k, p in crit.named_parameters():
        assert rel(p.grad.cpu(), ref_sd[k].grad) < tol, k
    assert rel(xg.grad.cpu(), xr.grad) < tol
    # frozen critic (train_seg phase): no parameter gradients, input gradient still flows
    for p in crit.parameters():
        p.requires_grad = False
        p.grad = None
    xg2 = x.cuda().requires_grad_(True)
    F.mse_loss(crit(xg2), torch.ones_like(y)).backward()
    assert all(p.grad is None for p in crit.parameters())
    assert rel(xg2.grad.cpu(), xr.grad) < tol


@pytest.mark.timeout(900)
def test_conf_segnet_step_matches_reference_golden_fp32(golden_dir):
    """Both phases of cm/train_trgb_segnet_conf.py:437-568 in the FP32 parity mode: losses and the gradient norm of
    every live parameter tensor against the reference's autograd (golden), 251 seg / 60 critic tensors."""
    from heatnet_pub_b200 import conf_segnet
    g = np.load(os.path.join(golden_dir, "conf_segnet_golden.npz"))
    m = conf_segnet.conv_segnet(pretrained=False, disc_arch='cyclegan', num_critics=6, no_conf=False, modalities='ir_rgb',
                                arch='pspnet', late_fusion=True)
    m.load_state_dict(O.recipe_fill(O.conf_segnet_state_dict(True, 6), seed=3))
    m = m.cuda().train()
    m.trgb_segnet.set_precision("fp32")
    for c in m.critics:
        c.precision = "fp32"
    m.trgb_segnet.drop_1.p = m.trgb_segnet.drop_2.p = 0.0
    rgb_d, ir_d = O.synthetic_inputs(1, 256, 256, seed=11)
    rgb_n, ir_n = O.synthetic_inputs(1, 256, 256, seed=12)
    label = torch.from_numpy(g["label"]).cuda()
    mse, ce = nn.MSELoss(), nn.CrossEntropyLoss()
    for phase in ("train_critic", "train_seg"):
        m.setPhase(phase)
        for p in m.parameters():
            p.grad = None
        o = m([rgb_d.cuda(), ir_d.cuda()], [rgb_n.cuda(), ir_n.cuda()])
        assert [str(tuple(c.shape)) for c in o['critics_a']] == [str(s) for s in g[phase + "/critic_out_shapes"]]
        total_critics = sum(mse(c, torch.full_like(c, 1)) for c in o['critics_a']) + sum(mse(c, torch.full_like(c, 0)) for c in o['critics_b'])
        assert abs(total_critics.item() - g[phase + "/total_critics"]) < 5e-4 * abs(g[phase + "/total_critics"])
        if phase == "train_seg":
            seg_loss = ce(o['pred_label_a'], label)
            conf = sum(mse(c, torch.full_like(c, 1)) for c in o['critics_a']) + sum(mse(c, torch.full_like(c, 1)) for c in o['critics_b'])
            total = seg_loss + 0.1 * conf
            assert abs(seg_loss.item() - g["seg_loss"]) < 2e-4 * abs(g["seg_loss"])
            assert abs(conf.item() - g["conf_loss"]) < 5e-4 * abs(g["conf_loss"])
        else:
            assert not o['pred_label_a'].requires_grad          # frozen seg net: no graph through it
            total = total_critics
        total.backward()
        names = [str(n) for n in g[phase + "/grad_names"]]
        live = [k for k, p in m.named_parameters() if p.grad is not None]
        assert live == names
        gn = np.array([dict(m.named_parameters())[k].grad.double().norm().item() for k in names])
        err = np.abs(gn - g[phase + "/grad_norm"]) / np.maximum(g[phase + "/grad_norm"], 1e-12)
        print(f"{phase}: {len(names)} gradient tensors, worst grad-norm rel err {err.max():.3e}")
        np.testing.assert_allclose(gn, g[phase + "/grad_norm"], rtol=3e-3, atol=1e-7)
