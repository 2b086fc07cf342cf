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


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_pspnet_backward_matches_oracle_autograd(precision):
    """Late-fusion PSPNet, train-mode BN, dropout off, loss = CE(logits) + mean-square of every tap: every
    parameter gradient against torch.autograd of the oracle evaluated in FP64.

    Gradients through 79 batch-statistic BN layers on a small map are ill-conditioned: the oracle's own FP32
    gradients differ from its FP64 ones by up to 1e-1 on individual tensors (measured: feats.conv1.weight 6.6e-3,
    feats.layer3.3.conv1.weight 1.1e-1), and three conv biases that sit in front of a BN have an exactly-zero true
    gradient.  Each tensor is therefore bounded by the oracle's own FP32 (resp. BF16-autocast) deviation from FP64,
    with errors measured against max(|ref|, 1e-4 * largest gradient entry of the whole net)."""
    from heatnet_pub_b200 import pspnet
    sd = O.recipe_fill(O.pspnet_state_dict(True, 4), seed=0)
    rgb, ir = O.synthetic_inputs(2, 64, 96)
    label = torch.randint(0, 13, (2, 64, 96), generator=torch.Generator().manual_seed(5))

    def loss_of(logits, taps):
        return F.cross_entropy(logits.float(), label.to(logits.device)) + sum((t.float() ** 2).mean() for t in taps[1:])

    def oracle(dtype, autocast=False):
        s = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            return _oracle_grads(s, lambda q: loss_of(*O.pspnet_forward(q, rgb.to(dtype), ir.to(dtype), late_fusion=True, training=True,
                                                                        dropout=False)[:2]))

    ref_loss, ref_g = oracle(torch.float64)
    _, floor_g = oracle(torch.float32, autocast=(precision == "bf16"))
    gmax = max(g.abs().max().item() for g in ref_g.values())

    def err(a, b):
        return ((a.double() - b.double()).abs().max() / max(b.abs().max().item(), 1e-4 * gmax)).item()

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
    worst, worst_floor, bad = 0.0, 0.0, []
    for k, g in ref_g.items():
        e, f = err(got[k].cpu(), g), err(floor_g[k].float(), g)
        worst, worst_floor = max(worst, e), max(worst_floor, f)
        bad.append((k, e, f))
    # per tensor: within 3x its own oracle noise, or within the worst oracle noise over all tensors (the
    # ill-conditioning is shared by the whole backward chain), never worse than that
    bad = [(k, e, f) for k, e, f in bad if e > max(3.0 * f, worst_floor, 2e-3 if precision == "fp32" else 5e-2)]
    print(f"[{precision}] loss {loss.item():.6f} (FP64 oracle {ref_loss:.6f}); worst gradient err {worst:.3e}; "
          f"oracle {'BF16-autocast' if precision == 'bf16' else 'FP32'} worst {worst_floor:.3e}")
    assert not bad, bad[:5]


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 3e-2)])
def test_critic_backward(precision, tol):
    """FCDiscriminator forward + backward (MSE vs 1): FP32 max-abs relative 1e-3; BF16 relative L2 3e-2 per tensor."""
    metric = rel if precision == "fp32" else rel_l2
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
    floor = {}
    if precision == "bf16":      # the oracle's own BF16-autocast gradients vs its FP32 ones = the noise floor
        bsd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        xb = x.clone().requires_grad_(True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            yb = O.fc_discriminator(xb, bsd, "")
            F.mse_loss(yb.float(), torch.ones_like(yr)).backward()
        floor = {k: rel_l2(bsd[k].grad.float(), ref_sd[k].grad) for k in bsd}
        floor["x"] = rel_l2(xb.grad.float(), xr.grad)
        print("critic bf16 noise floor (oracle autocast):", {k: round(v, 4) for k, v in floor.items()})
    for k, p in crit.named_parameters():
        assert metric(p.grad.cpu(), ref_sd[k].grad) < max(tol, 1.5 * floor.get(k, 0.0)), k
    assert metric(xg.grad.cpu(), xr.grad) < max(tol, 1.5 * floor.get("x", 0.0))
    # frozen critic (train_seg phase): no parameter gradients, input gradient still flows
    for p in crit.parameters():
        p.requires_grad = False
        p.grad = None
    xg2 = x.cuda().requires_grad_(True)
    F.mse_loss(crit(xg2), torch.ones_like(y)).backward()
    assert all(p.grad is None for p in crit.parameters())
    assert metric(xg2.grad.cpu(), xr.grad) < max(tol, 1.5 * floor.get("x", 0.0))


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
        # conv biases in front of a BN have an exactly-zero true gradient (both sides hold rounding noise there):
        # absolute slack of 1e-5 of the largest gradient norm
        np.testing.assert_allclose(gn, g[phase + "/grad_norm"], rtol=3e-3, atol=1e-5 * g[phase + "/grad_norm"].max())


def test_graphed_train_step_matches_eager():
    """graphs.GraphedStep: the captured conv_segnet train_seg step (fwd + fused losses + bwd + fused RMSprop) replays to the
    same parameters, BN running statistics and loss as the eager loop (FP32 reds in wgrad are unordered: tolerance, not bits),
    and bumps the version counters so a later eval forward sees the updated weights."""
    import contextlib
    import io
    from heatnet_pub_b200 import conf_segnet, graphs, losses, optim

    def make():
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            m = conf_segnet.conv_segnet(pretrained=False, disc_arch='cyclegan', num_critics=6, no_conf=False, modalities='ir_rgb',
                                        arch='pspnet', late_fusion=True)
            m = m.cuda().train()
            m.setPhase('train_seg')
        for mod in m.modules():
            if isinstance(mod, nn.Dropout2d):
                mod.p = 0.0                       # identical arithmetic in both runs (the RNG stream positions differ)
        return m

    g = torch.Generator().manual_seed(3)
    mk = lambda c: (torch.rand(1, c, 256, 256, generator=g) * 2 - 1).cuda()
    batches = [(mk(3), mk(1), mk(3), mk(1), torch.randint(0, 13, (1, 256, 256), generator=g).cuda()) for _ in range(2)]
    mse, ce = losses.MSELoss(), losses.CrossEntropyLoss()

    def runner(model, opt):
        def train_step(rgb_d, ir_d, rgb_n, ir_n, label):
            for p in model.parameters():
                p.grad = None
            o = model([rgb_d, ir_d], [rgb_n, ir_n])
            conf = sum(mse(c, 1.0) for c in o['critics_a']) + sum(mse(c, 1.0) for c in o['critics_b'])
            total = ce(o['pred_label_a'], label) + 0.1 * conf
            total.backward()
            opt.step()
            return total
        return train_step

    def eager_run():
        m = make()
        opt = optim.RMSprop(m.parameters(), lr=1e-5)
        st = runner(m, opt)
        for _ in range(3):
            st(*batches[0])
        ls = [float(st(*bt).detach()) for bt in (batches[1], batches[0])]
        torch.cuda.synchronize()
        return m, ls

    a, la = eager_run()
    a2, la2 = eager_run()          # the eager loop against itself: FP32 reds in wgrad are unordered and RMSprop's first steps are
    #                                sign-like (zero-initialised BN biases move by +-10 lr whatever |g|), so runs decorrelate on tiny gradients
    b = make()
    opt_b = optim.RMSprop(b.parameters(), lr=1e-5)
    gstep = graphs.GraphedStep(runner(b, opt_b), batches[0], module=b, warmup=3)
    v0 = b.trgb_segnet.final[0].weight._version
    lb = [float(gstep(*bt).detach()) for bt in (batches[1], batches[0])]
    assert b.trgb_segnet.final[0].weight._version > v0
    torch.cuda.synchronize()
    for x, y, z in zip(la, lb, la2):
        assert abs(x - y) < max(2e-3 * abs(x), 3 * abs(x - z)), (la, lb, la2)
    sa, sa2, sb = a.state_dict(), a2.state_dict(), b.state_dict()
    worst = (0.0, None)
    for k in sa:
        if sa[k].is_floating_point():
            noise = rel_l2(sa2[k].cpu(), sa[k].cpu())
            dev = rel_l2(sb[k].cpu(), sa[k].cpu())
            assert dev < max(2e-3, 3 * noise), (k, dev, noise)
            worst = max(worst, (dev, k, noise))
        else:
            assert torch.equal(sa[k], sb[k]), k            # num_batches_tracked
    print("graph vs eager: worst deviation", worst)


@pytest.mark.parametrize("phase", ["train_seg", "train_critic"])
def test_paired_forward_equals_two_consecutive_calls(phase, monkeypatch):
    """conv_segnet's batched [day; night] pass (PSPNet.forward_pair: shared convolution launches, per-domain BatchNorm statistics via
    statistics groups in the conv epilogue / apply / backward kernels) against the two consecutive seg-net calls of
    cm/models/conf_segnet.py:114-115: every output, every parameter gradient and every BN running statistic.  Both runs are BF16;
    they differ only by summation order (statistics atomics, wgrad over 2B instead of B + B), bounded at BF16 resolution."""
    import contextlib
    import io
    from heatnet_pub_b200 import conf_segnet, losses
    sd = O.recipe_fill(O.conf_segnet_state_dict(True, 6), seed=3)
    B, H, W = 2, 256, 256
    day = [t.cuda() for t in O.synthetic_inputs(B, H, W, seed=31)]
    night = [t.cuda() for t in O.synthetic_inputs(B, H, W, seed=32)]
    label = torch.randint(0, 13, (B, H, W), generator=torch.Generator().manual_seed(33)).cuda()
    mse, ce = losses.MSELoss(), losses.CrossEntropyLoss()

    def run(pair):
        monkeypatch.setattr(conf_segnet, "PAIR_FORWARD", pair)
        with contextlib.redirect_stdout(io.StringIO()):
            m = conf_segnet.conv_segnet(pretrained=False, disc_arch='cyclegan', num_critics=6, no_conf=False, modalities='ir_rgb',
                                        arch='pspnet', late_fusion=True)
            m.load_state_dict(sd)
            m = m.cuda().train()
            m.setPhase(phase)
        m.trgb_segnet.drop_1.p = m.trgb_segnet.drop_2.p = 0.0
        o = m(list(day), list(night))
        if phase == "train_seg":
            total = ce(o['pred_label_a'], label) + 0.1 * (sum(mse(c, 1.0) for c in o['critics_a']) + sum(mse(c, 1.0) for c in o['critics_b']))
        else:
            total = sum(mse(c, 1.0) for c in o['critics_a']) + sum(mse(c, 0.0) for c in o['critics_b'])
        total.backward()
        outs = {k: ([t.detach().float().clone() for t in v] if isinstance(v, list) else v.detach().float().clone())
                for k, v in o.items() if v is not None}
        return m, total.item(), outs, {k: p.grad.float().clone() for k, p in m.named_parameters() if p.grad is not None}

    ms, ls, os_, gs = run(False)
    mp, lp, op, gp = run(True)
    assert abs(lp - ls) < 5e-3 * abs(ls), (lp, ls)
    for k in os_:
        a, b = os_[k], op[k]
        for x, y in (zip(a, b) if isinstance(a, list) else [(a, b)]):
            assert x.shape == y.shape, k
            assert rel(y.cpu(), x.cpu()) < 3e-2, k
    assert set(gs) == set(gp)
    worst = max((rel_l2(gp[k].cpu(), gs[k].cpu()), k) for k in gs if gs[k].abs().max() > 0)
    print(f"[{phase}] paired vs sequential: loss {lp:.6f} / {ls:.6f}; worst gradient rel-L2 deviation {worst}")
    assert worst[0] < 6e-2, worst
    ss, sp = ms.state_dict(), mp.state_dict()
    for k in ss:
        if k.endswith("num_batches_tracked"):
            assert int(ss[k]) == int(sp[k]) == 2, k                    # two forward calls' worth of updates, in order
        elif "running_" in k:
            assert rel(sp[k].cpu(), ss[k].cpu()) < 2e-3, k


def test_projected_psp_bottleneck_trains_like_the_concat_formulation(monkeypatch):
    """BF16 training runs the PSP bottleneck as W_f x5 + b + sum_s up(W_s prior_s) with one hand-written backward
    (engine.record_projected_bottleneck) instead of the 10240-channel concat of cm/models/pspnet.py:21-24.  The block alone, driven
    through the tape with the SAME incoming gradient for both formulations (inside the full step the two forward results differ in
    the last BF16 bit and the train-mode BatchNorm layers downstream turn that into O(10 %) gradient noise -- the whole-step accuracy
    is bounded against the FP32 oracle in test_gpu_bf16_training.py): output, input gradient, and every parameter gradient --
    bottleneck.weight's five column blocks come from five weight-gradient GEMMs into one sink -- against the literal formulation
    AND against torch autograd of the reference module in FP64 on the BF16-rounded operands."""
    from heatnet_pub_b200 import engine as E, pspnet
    g = torch.Generator().manual_seed(5)
    N, H, W, Cf, Co = 3, 20, 24, 256, 128
    ref = nn.Module()
    ref.stages = nn.ModuleList([nn.Sequential(nn.AdaptiveAvgPool2d((s, s)), nn.Conv2d(Cf, Cf, 1, bias=False)) for s in (1, 2, 3, 6)])
    ref.bottleneck = nn.Conv2d(Cf * 5, Co, 1)
    x = torch.randn(N, Cf, H, W, generator=g)
    dout = torch.randn(N, Co, H, W, generator=g)

    def run(projected):
        monkeypatch.setattr(pspnet, "PSP_PROJECTED_TRAIN", projected)
        m = pspnet.PSPModule(Cf, Co).cuda().train()
        m.precision = "bf16"
        m.load_state_dict(ref.state_dict())
        tape = E.Tape()
        E.current_tape = tape
        try:
            if projected:
                feats = E.from_nchw(x.cuda(), torch.bfloat16)
                tape.require(feats)
                y = m._run(feats)
            else:           # the literal formulation reads the features from their channel slice of the concat buffer
                cat = m.alloc_cat(N, H, W, torch.bfloat16, "cuda")
                feats = E.from_nchw(x.cuda(), torch.bfloat16, m.feats_slice(cat))
                tape.require(feats)
                y = m._run_cat(cat)
        finally:
            E.current_tape = None
        grads = E.Grads()
        gy, _ = grads.target(y)
        E.from_nchw(dout.cuda(), torch.bfloat16, gy)
        grads.mark(y)
        tape.backward(grads)
        pg = {k: grads.params[p].float().cpu().clone() for k, p in m.named_parameters()}
        return y.nchw().float().cpu(), grads.get(feats).nchw().float().cpu(), pg

    yc, dxc, gc = run(False)
    yp, dxp, gp = run(True)
    # FP64 autograd of the reference module on BF16-rounded operands
    r64 = ref.double()
    with torch.no_grad():
        for p in r64.parameters():
            p.copy_(p.float().bfloat16().double())
    xr = x.bfloat16().double().requires_grad_(True)
    priors = [F.interpolate(st(xr), size=(H, W), mode="bilinear", align_corners=False) for st in r64.stages]
    yr = F.relu(r64.bottleneck(torch.cat(priors + [xr], 1)))
    yr.backward(dout.bfloat16().double())
    gr = {k: p.grad.float() for k, p in r64.named_parameters()}
    assert rel(yp, yr.detach().float()) < 2e-2 and rel(yc, yr.detach().float()) < 2e-2
    errs = {"dx": (rel_l2(dxp, xr.grad.float()), rel_l2(dxc, xr.grad.float()))}
    for k in gr:
        assert gp[k].shape == gr[k].shape, k
        errs[k] = (rel_l2(gp[k], gr[k]), rel_l2(gc[k], gr[k]))
    print("rel-L2 error vs FP64 autograd (projected, concat):", {k: (round(a, 4), round(b, 4)) for k, (a, b) in errs.items()})
    for k, (a, b) in errs.items():
        assert a < max(1.5e-2, 1.5 * b), (k, a, b)          # as accurate as the literal formulation on the same engine
    blocks = gp["bottleneck.weight"].view(Co, 5, Cf)
    assert all(blocks[:, i].abs().max() > 0 for i in range(5))
