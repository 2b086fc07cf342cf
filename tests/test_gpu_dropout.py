"""nn.Dropout2d on the training / validation path (cm/models/pspnet.py:49,55,64-73): the reference never disables
dropout while training and its validate_model runs without .eval() (validation_bdd_mf.py:263), so the masked path is
the one every real step takes.  torch's RNG stream cannot be matched by another kernel: parity runs inject the same
four (B, C) keep-masks into the product (`PSPNet._injected_dropout_masks`) and into the oracle
(`pspnet_forward(dropout_masks=...)`); the product's own generator (keyed Philox) is checked bit for bit against
oracle/random_oracle.py."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import heatnet_oracle as O
from oracle import random_oracle as R

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2


def rel(a, b):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _masks(batch, seed=7):
    """Four seeded keep-masks for drop_1 (1024 ch, p=.3) and the three drop_2 calls (256, 64, 64 ch, p=.15)."""
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand(batch, c, generator=g) >= p).float() for c, p in ((1024, O.DROP_P1), (256, O.DROP_P2), (64, O.DROP_P2), (64, O.DROP_P2))]


def _net(sd, precision):
    from heatnet_pub_b200 import pspnet
    net = pspnet.PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50', in_channels=4,
                        pretrained=False, late_fusion=True)
    net.load_state_dict(sd)
    return net.cuda().train().set_precision(precision)


def test_mask_kernel_matches_philox_oracle():
    """hn_dropout2d_scale: bit-exact against the numpy restatement of Philox4x32-10, the call counter advances on the
    device, the same seed reproduces the sequence, p = 1 gives zeros (torch: all channels dropped, no NaN)."""
    from heatnet_pub_b200 import engine as E, _lib
    _lib.require_device()
    lib = _lib.load()
    seed = 0x1234_5678_9ABC
    st = E.dropout_seed(seed)
    for call, (n, p) in enumerate([(16 * 1024, 0.3), (4096, 0.15), (5, 0.15), (1000, 0.0), (1000, 1.0), (3000, 0.5)]):
        out = torch.full((n,), -1.0, device="cuda")
        _lib.check(lib.hn_dropout2d_scale(st.data_ptr(), n, p, out.data_ptr(), E._stream()))
        want = R.dropout2d_scale(seed, call, n, p)
        assert np.array_equal(out.cpu().numpy(), want), (call, n, p)
        assert int(st[1].item()) == call + 1
        if 0.0 < p < 1.0 and n >= 3000:
            assert abs((out > 0).float().mean().item() - (1.0 - p)) < 0.03
    assert torch.count_nonzero(torch.from_numpy(R.dropout2d_scale(seed, 4, 1000, 1.0))) == 0
    st = E.dropout_seed(seed)            # reseeding restarts the sequence
    out = torch.empty(4096, device="cuda")
    _lib.check(lib.hn_dropout2d_scale(st.data_ptr(), 4096, 0.3, out.data_ptr(), E._stream()))
    assert np.array_equal(out.cpu().numpy(), R.dropout2d_scale(seed, 0, 4096, 0.3))
    with pytest.raises(RuntimeError, match="dropout probability has to be between 0 and 1"):
        _lib.check(lib.hn_dropout2d_scale(st.data_ptr(), 16, 1.5, out.data_ptr(), E._stream()))


@pytest.mark.parametrize("p", [0.0, 0.15, 1.0])
def test_dropout2d_op_semantics(p):
    """engine.dropout2d against F.dropout2d's definition with an injected mask, including p = 1 (zeros, not NaN) and the
    (n, c) indexing of the per-image scale."""
    from heatnet_pub_b200 import engine as E
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 40, 5, 7, generator=g)
    mask = (torch.rand(3, 40, generator=g) >= 0.4).float() if p > 0.0 else torch.ones(3, 40)     # p = 0: nn.Dropout2d is the identity
    for dtype in (torch.float32, torch.bfloat16):
        a = E.from_nchw(x.cuda(), dtype)
        y = E.dropout2d(a, p, mask).nchw().float().cpu()
        want = O.dropout2d(x.to(dtype).float(), p, True, mask)
        assert torch.isfinite(y).all()
        if p == 1.0:
            assert y.abs().max().item() == 0.0 and want.abs().max().item() == 0.0
        else:
            assert rel(y, want) < (1e-6 if dtype == torch.float32 else 8e-3)
    with pytest.raises(ValueError):
        E.dropout2d(a, 1.5, mask)


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_train_forward_with_injected_masks_matches_oracle(precision, tol):
    """Train-mode forward (batch-stat BN + the four Dropout2d calls) with injected masks: logits and the five taps
    against the oracle given the same masks.  FP32: 1e-4.  BF16 + batch-stat BN is bounded by the oracle's own
    BF16-autocast error on the same inputs (see test_gpu_network: 79 re-normalisations amplify BF16 rounding for any
    implementation), both numbers printed."""
    sd = O.recipe_fill(O.pspnet_state_dict(True, 4), seed=0)
    rgb, ir = O.synthetic_inputs(2, 64, 96)
    masks = _masks(2)
    with torch.no_grad():
        ref, ref_taps, _ = O.pspnet_forward({k: v.clone() for k, v in sd.items()}, rgb, ir, late_fusion=True, training=True, dropout_masks=masks)
        ref_nodrop, _, _ = O.pspnet_forward({k: v.clone() for k, v in sd.items()}, rgb, ir, late_fusion=True, training=True, dropout=False)
    assert rel(ref_nodrop, ref) > 0.05, "the masks must matter for this test to mean anything"
    net = _net(sd, precision)
    net._injected_dropout_masks = [m.clone() for m in masks]
    with torch.no_grad():
        logits, taps, _ = net(rgb.cuda(), ir.cuda())
    assert net._injected_dropout_masks == [], "all four masks are consumed in order: drop_1, drop_2 x3"
    bound = tol
    if precision == "bf16":
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            lo, _, _ = O.pspnet_forward({k: v.clone() for k, v in sd.items()}, rgb, ir, late_fusion=True, training=True, dropout_masks=masks)
        floor = rel(lo.float(), ref)
        bound = max(tol, 1.25 * floor)
        print(f"[bf16 train + dropout] logits rel err {rel(logits.cpu(), ref):.3f}; torch CPU bf16-autocast floor {floor:.3f}")
    assert rel(logits.cpu(), ref) < bound
    for i in range(1, 6):        # the encoder taps sit in front of the dropouts
        assert rel(taps[i].float().cpu(), ref_taps[i]) < bound


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_backward_through_injected_masks_matches_oracle_autograd(precision):
    """Parameter gradients of CE(logits) through the masked decoder (the hn_affine_act(per_image=1) adjoint, three
    independent drop_2 masks, the 1/(1-p) factor) against autograd of the oracle in FP64 given the same masks; per tensor
    bounded by the oracle's own FP32 (resp. BF16-autocast) deviation from FP64, as in test_gpu_training."""
    sd = O.recipe_fill(O.pspnet_state_dict(True, 4), seed=0)
    rgb, ir = O.synthetic_inputs(2, 64, 96)
    label = torch.randint(0, 13, (2, 64, 96), generator=torch.Generator().manual_seed(5))
    masks = _masks(2, seed=11)

    def oracle(dtype, autocast=False):
        s = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        keys = [k for k, v in s.items() if v.is_floating_point() and "running_" not in k]
        for k in keys:
            s[k].requires_grad_(True)
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            logits = O.pspnet_forward(s, rgb.to(dtype), ir.to(dtype), late_fusion=True, training=True, dropout_masks=[m.to(dtype) for m in masks])[0]
        loss = F.cross_entropy(logits.float(), label)
        loss.backward()
        return loss.item(), {k: s[k].grad for k in keys if s[k].grad is not None}

    ref_loss, ref_g = oracle(torch.float64)
    _, floor_g = oracle(torch.float32, autocast=(precision == "bf16"))
    gmax = max(g.abs().max().item() for g in ref_g.values())

    def err(a, b):
        return ((a.double() - b.double()).abs().max() / max(b.abs().max().item(), 1e-4 * gmax)).item()

    net = _net(sd, precision)
    net._injected_dropout_masks = [m.clone() for m in masks]
    logits, _, _ = net(rgb.cuda(), ir.cuda())
    loss = F.cross_entropy(logits, label.cuda())
    loss.backward()
    got = {k: p.grad for k, p in net.named_parameters()}
    assert set(k for k, g in got.items() if g is not None) == set(ref_g)
    if precision == "fp32":
        assert abs(loss.item() - ref_loss) < 1e-4 * abs(ref_loss)
    rows = [(k, err(got[k].cpu(), g), err(floor_g[k].float(), g)) for k, g in ref_g.items()]
    worst, worst_floor = max(r[1] for r in rows), max(r[2] for r in rows)
    bad = [r for r in rows if r[1] > max(3.0 * r[2], worst_floor, 2e-3 if precision == "fp32" else 5e-2)]
    print(f"[{precision} + dropout] loss {loss.item():.6f} (FP64 oracle {ref_loss:.6f}); worst gradient err {worst:.3e}; oracle floor {worst_floor:.3e}")
    assert not bad, bad[:5]
    # decoder tensors directly behind a mask carry its zero pattern: channels dropped by drop_2 before up_3 leave
    # exactly-zero filter slices in up_3.conv.0.weight's input dimension only when dropped in EVERY image
    dead = (masks[2].sum(0) == 0).nonzero().flatten().tolist()
    for c in dead:
        assert got["up_3.conv.0.weight"][:, c].abs().max().item() == 0.0


def test_graph_replays_draw_new_masks_and_seed_reproduces():
    """A captured train-mode forward draws a fresh Dropout2d mask on every replay (the generator state lives on the device
    and the mask kernel advances it), and reseeding reproduces the sequence exactly."""
    from heatnet_pub_b200 import engine as E, graphs
    sd = O.recipe_fill(O.pspnet_state_dict(True, 4), seed=0)
    net = _net(sd, "bf16")
    rgb, ir = O.synthetic_inputs(2, 64, 96)
    rgb, ir = rgb.cuda(), ir.cuda()

    def fwd(a, b):
        with torch.no_grad():
            return net(a, b)[0]

    E.dropout_seed(1234)
    gstep = graphs.GraphedStep(fwd, [rgb, ir], module=net, warmup=1)
    calls0 = int(E._dropout_state[torch.cuda.current_device()][1].item())
    a = gstep(rgb, ir).clone()
    b = gstep(rgb, ir).clone()
    assert int(E._dropout_state[torch.cuda.current_device()][1].item()) == calls0 + 8      # 4 masks per forward, 2 replays
    # same input, same BN batch statistics: the only difference between two replays is the dropout mask
    assert rel(a.cpu(), b.cpu()) > 5e-2
    # eager with the generator rewound to where the first replay started gives the first replay's logits
    E._dropout_state[torch.cuda.current_device()][1] = calls0
    c = fwd(rgb, ir)
    assert rel(c.cpu(), a.cpu()) < 1e-3
