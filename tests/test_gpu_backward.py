"""Backward kernels on a B200 against torch.autograd on the CPU FP32 oracle primitives (the reference obtains
these gradients from torch.autograd).  FP32 kernels: 1e-4 relative; BF16 tensor-core dgrad/wgrad: 2e-2 against
the FP32 oracle and 2e-3 against an FP32 evaluation on the same BF16-rounded operands."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2


@pytest.fixture(scope="module")
def E():
    from heatnet_pub_b200 import _lib, engine
    _lib.require_device()
    return engine


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def to_act(E, t, dtype):
    return E.from_nchw(t.cuda(), dtype)


def back(act):
    return act.nchw().float().cpu()


GRAD_CASES = [
    # cin, cout, k, stride, pad, dil, h, w
    (64, 64, 1, 1, 0, 1, 20, 24),
    (64, 256, 1, 1, 0, 1, 12, 24),
    (128, 128, 3, 1, 1, 1, 10, 12),
    (256, 256, 3, 1, 2, 2, 10, 13),
    (512, 512, 3, 1, 4, 4, 11, 12),
    (128, 128, 3, 2, 1, 1, 21, 24),      # stride 2, odd height
    (256, 512, 1, 2, 0, 1, 21, 24),      # 1x1 stride 2
    (2048, 1024, 1, 1, 0, 1, 6, 9),
    (1024, 256, 3, 1, 1, 1, 12, 16),
    (64, 13, 1, 1, 0, 1, 16, 24),        # final
    (13, 64, 4, 2, 1, 1, 32, 48),        # critic conv1 on logits
    (128, 64, 4, 2, 1, 1, 16, 24),       # critic conv1 on x1
    (512, 1, 4, 2, 1, 1, 4, 6),          # critic classifier
    (3, 64, 7, 2, 3, 1, 33, 40),         # RGB stem (wgrad only matters; dgrad checked too)
    (256, 512, 4, 2, 1, 1, 40, 80),      # stride-2 wgrad through the element-strided X map, several pixel blocks
    (64, 128, 4, 2, 1, 1, 33, 47),       # odd input, ragged blocks
    (256, 384, 3, 1, 1, 1, 9, 11),       # wgrad on CTA pairs with an odd number of Cout tiles (filler tile beyond Cout)
    (128, 256, 3, 1, 1, 1, 10, 12),      # wgrad on CTA pairs: 18 column blocks = four items of 4 and a last one of 2
    (512, 320, 1, 1, 0, 1, 7, 9),        # wgrad on CTA pairs: Cout not a multiple of the 128-row tile
]


def _case(case, seed=0):
    cin, cout, k, stride, pad, dil, h, w = case
    g = torch.Generator().manual_seed(seed)
    conv = nn.Conv2d(cin, cout, k, stride, pad, dil, bias=False)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (k * k * cin)) ** 0.5)
    x = torch.randn(2, cin, h, w, generator=g)
    ho, wo = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1, (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    dy = torch.randn(2, cout, ho, wo, generator=g)
    return conv, x, dy


def _ref_grads(conv, x, dy, round_bf16=False):
    w = conv.weight.detach().clone()
    if round_bf16:
        x, w, dy = x.bfloat16().float(), w.bfloat16().float(), dy.bfloat16().float()
    x = x.clone().requires_grad_(True)
    w.requires_grad_(True)
    y = F.conv2d(x, w, None, conv.stride, conv.padding, conv.dilation)
    y.backward(dy)
    return x.grad, w.grad


@pytest.mark.parametrize("case", GRAD_CASES)
def test_conv_dgrad_wgrad_fp32(E, case):
    conv, x, dy = _case(case)
    dx_ref, dw_ref = _ref_grads(conv, x, dy)
    convg = copy.deepcopy(conv).cuda()
    dya = E.from_nchw(dy.cuda(), torch.float32) if case[1] % 8 == 0 else _padded_act(E, dy, torch.float32)
    dx = E.conv2d_dgrad(dya, convg, x.shape[2], x.shape[3])
    assert rel(back(dx), dx_ref) < FP32_TOL
    dw = E.conv2d_wgrad(to_act(E, x, torch.float32), dya, convg)
    assert rel(dw.cpu(), dw_ref) < FP32_TOL
    # second contribution accumulated into existing storage (a gradient-arena slot; the 1x1 layers accumulate in the kernel)
    prev = torch.randn(dw.shape, generator=torch.Generator().manual_seed(2)).cuda()
    acc = E.conv2d_wgrad(to_act(E, x, torch.float32), dya, convg, out=prev.clone(), accumulate=True)
    assert rel((acc - prev).cpu(), dw_ref) < 2 * FP32_TOL


def _padded_act(E, t, dtype):
    """NHWC act whose channel stride is padded to a multiple of 8 (what the training path allocates for C=13 / C=1)."""
    n, c, h, w = t.shape
    a = E.new_act(n, h, w, c, dtype, "cuda", ld=(c + 7) // 8 * 8)
    a.buf.zero_()
    E.from_nchw(t.cuda(), dtype, a)
    return a


@pytest.mark.parametrize("case", GRAD_CASES)
def test_conv_dgrad_wgrad_bf16_tensor_core(E, case):
    conv, x, dy = _case(case)
    dx_ref, dw_ref = _ref_grads(conv, x, dy)
    dx_r, dw_r = _ref_grads(conv, x, dy, round_bf16=True)
    convg = copy.deepcopy(conv).cuda()
    dya = _padded_act(E, dy, torch.bfloat16)
    dx = E.conv2d_dgrad(dya, convg, x.shape[2], x.shape[3])
    assert rel(back(dx), dx_r) < 1.2e-2          # BF16 output rounding of dX
    assert rel(back(dx), dx_ref) < BF16_TOL
    dw = E.conv2d_wgrad(to_act(E, x, torch.bfloat16), dya, convg)
    assert dw.dtype == torch.float32
    assert rel(dw.cpu(), dw_r) < 2e-3            # FP32 accumulation of BF16 products
    assert rel(dw.cpu(), dw_ref) < BF16_TOL
    prev = torch.randn(dw.shape, generator=torch.Generator().manual_seed(2)).cuda()
    acc = E.conv2d_wgrad(to_act(E, x, torch.bfloat16), dya, convg, out=prev.clone(), accumulate=True)
    assert rel((acc - prev).cpu(), dw_r) < 4e-3


S2_CASES = [
    # cin, cout, k, stride, pad, dil, h, w        stride-2 layers of the step: critics (4x4 p1), layer2.0 conv2 (3x3 p1) / downsample (1x1)
    (64, 64, 4, 2, 1, 1, 32, 48),
    (128, 256, 4, 2, 1, 1, 17, 31),      # odd input size: the four phase lattices differ in extent
    (2048, 64, 4, 2, 1, 1, 8, 12),
    (13, 64, 4, 2, 1, 1, 32, 48),        # 13-channel dX (critic on the logits): 16-wide Cout tile, channel stride 16
    (128, 128, 3, 2, 1, 1, 20, 25),
    (256, 512, 1, 2, 0, 1, 21, 24),      # three empty phases: zero-filled, or untouched when accumulating
    (64, 64, 7, 2, 3, 1, 20, 24),        # up to 4 taps per axis and phase
]


@pytest.mark.parametrize("case", S2_CASES)
def test_dgrad_stride2_parity_phases(E, case, monkeypatch):
    """hn_conv2d_dgrad_s2 (four parity phases onto strided sub-lattices of dX) against autograd, against the zero-insertion
    form it replaces, and in accumulate mode."""
    conv, x, dy = _case(case)
    dx_ref, _ = _ref_grads(conv, x, dy)
    dx_r, _ = _ref_grads(conv, x, dy, round_bf16=True)
    convg = copy.deepcopy(conv).cuda()
    dya = _padded_act(E, dy, torch.bfloat16)
    assert E.DGRAD_S2_PHASES
    l0 = E.launch_count
    dx = E.conv2d_dgrad(dya, convg, x.shape[2], x.shape[3])
    assert E.launch_count - l0 <= 8, "phase path expected (no dilate / im2col launches)"
    assert rel(back(dx), dx_r) < 1.2e-2 and rel(back(dx), dx_ref) < BF16_TOL
    monkeypatch.setattr(E, "DGRAD_S2_PHASES", False)
    dx_old = E.conv2d_dgrad(dya, convg, x.shape[2], x.shape[3])
    monkeypatch.setattr(E, "DGRAD_S2_PHASES", True)
    assert rel(back(dx), back(dx_old)) < 1.2e-2        # same products, different summation order / one BF16 rounding
    prev = torch.randn(x.shape, generator=torch.Generator().manual_seed(9))
    out = _padded_act(E, prev, torch.bfloat16)
    E.conv2d_dgrad(dya, convg, x.shape[2], x.shape[3], out=out, accumulate=True)
    assert rel(back(out), dx_r + prev.bfloat16().float()) < 1.5e-2


def test_dgrad_accumulates_through_epilogue(E):
    conv, x, dy = _case((128, 128, 3, 1, 1, 1, 10, 12))
    dx_ref, _ = _ref_grads(conv, x, dy)
    convg = copy.deepcopy(conv).cuda()
    prev = torch.randn(x.shape, generator=torch.Generator().manual_seed(9))
    out = to_act(E, prev, torch.float32)
    E.conv2d_dgrad(to_act(E, dy, torch.float32), convg, x.shape[2], x.shape[3], out=out, accumulate=True)
    assert rel(back(out), dx_ref + prev) < FP32_TOL


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("act", ["relu", "prelu", "none"])
def test_bn_train_backward(E, dtype, tol, act):
    """conv_raw -> BN(train) -> (+res) -> act: gradients wrt the pre-BN tensor, gamma, beta, residual, PReLU slope."""
    g = torch.Generator().manual_seed(3)
    N, Cc, H, W = 3, 64, 9, 14
    raw = (torch.randn(N, Cc, H, W, generator=g) * 1.5 + 0.3).requires_grad_(True)
    res = torch.randn(N, Cc, H, W, generator=g).requires_grad_(True) if act == "relu" else None
    gamma = (torch.rand(Cc, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(Cc, generator=g) * 0.1).requires_grad_(True)
    slope = torch.tensor([0.3], requires_grad=True)
    z = F.batch_norm(raw, None, None, gamma, beta, True, 0.1, 1e-5)
    if res is not None:
        z = z + res
    out = F.relu(z) if act == "relu" else (F.prelu(z, slope) if act == "prelu" else z)
    dout = torch.randn(N, Cc, H, W, generator=g)
    out.backward(dout)
    mean = raw.detach().mean((0, 2, 3))
    invstd = 1.0 / torch.sqrt(raw.detach().var((0, 2, 3), unbiased=False) + 1e-5)
    code = {"relu": E.ACT_RELU, "prelu": E.ACT_LEAKY, "none": E.ACT_NONE}[act]
    dres = E.new_act(N, H, W, Cc, dtype, "cuda") if res is not None else None
    def sinks(prefill=None):
        """(tensor, accumulate) for dbeta, dgamma, dslope; with `prefill` the kernel adds to existing contents"""
        ts = [torch.full((Cc,), prefill or 0.0, device="cuda"), torch.full((Cc,), prefill or 0.0, device="cuda"),
              torch.full((1,), prefill or 0.0, device="cuda") if act == "prelu" else None]
        return [(t, prefill is not None) for t in ts]

    sk = sinks()
    draw = E.bn_bwd(to_act(E, dout, dtype), to_act(E, out.detach(), dtype), to_act(E, raw.detach(), torch.float32), mean.cuda(),
                    invstd.cuda(), gamma.detach().cuda(), code, slope_ptr=slope.detach().cuda() if act == "prelu" else None,
                    dres=dres, sinks=sk)
    pg = [t for t, _ in sk]
    # second contribution into the same parameter-gradient storage (the net runs on the day and on the night batch)
    sk_acc = sinks(prefill=2.0)
    E.bn_bwd(to_act(E, dout, dtype), to_act(E, out.detach(), dtype), to_act(E, raw.detach(), torch.float32), mean.cuda(),
             invstd.cuda(), gamma.detach().cuda(), code, slope_ptr=slope.detach().cuda() if act == "prelu" else None, sinks=sk_acc)
    assert torch.allclose(sk_acc[0][0], pg[0] + 2.0, rtol=1e-5, atol=1e-5) and torch.allclose(sk_acc[1][0], pg[1] + 2.0, rtol=1e-5, atol=1e-5)
    assert rel(back(draw), raw.grad) < tol
    assert rel(pg[0].cpu(), beta.grad) < tol and rel(pg[1].cpu(), gamma.grad) < tol
    if res is not None:
        assert rel(back(dres), res.grad) < tol
    if res is None:
        # without a residual the saved output need not be read: z is recomputed from raw with the forward's scale / shift
        fscale = (gamma.detach() * invstd).cuda()
        fshift = (beta.detach() - mean * gamma.detach() * invstd).cuda()
        sk2 = sinks()
        draw2 = E.bn_bwd(to_act(E, dout, dtype), to_act(E, out.detach(), dtype), to_act(E, raw.detach(), torch.float32), mean.cuda(),
                         invstd.cuda(), gamma.detach().cuda(), code, slope_ptr=slope.detach().cuda() if act == "prelu" else None,
                         sinks=sk2, fwd_scale=fscale, fwd_shift=fshift)
        pg2 = [t for t, _ in sk2]
        assert rel(back(draw2), raw.grad) < tol
        assert rel(pg2[0].cpu(), beta.grad) < tol and rel(pg2[1].cpu(), gamma.grad) < tol
        if act == "prelu":
            assert abs(pg2[2].item() - slope.grad.item()) < tol * max(1.0, abs(slope.grad.item())) * 5
    if act == "prelu":
        assert abs(pg[2].item() - slope.grad.item()) < tol * max(1.0, abs(slope.grad.item())) * 5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_act_bwd_and_bias_grad(E, dtype):
    g = torch.Generator().manual_seed(4)
    z = torch.randn(2, 64, 7, 9, generator=g).to(dtype).float().requires_grad_(True)
    out = F.leaky_relu(z, 0.2)
    dout = torch.randn(2, 64, 7, 9, generator=g).to(dtype).float()
    out.backward(dout)
    dz = E.act_bwd(to_act(E, dout, dtype), to_act(E, out.detach(), dtype), E.ACT_LEAKY, 0.2)
    assert rel(back(dz), z.grad) < (1e-6 if dtype == torch.float32 else 1e-2)
    sums = E.channel_sums(dz)
    assert rel(sums[0].cpu(), back(dz).sum((0, 2, 3))) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("hw", [(17, 24), (16, 16), (7, 9)])
def test_maxpool_backward(E, dtype, hw):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, *hw, generator=g).to(dtype).float().requires_grad_(True)
    y = F.max_pool2d(x, 3, 2, 1)
    dy = torch.randn(y.shape, generator=g).to(dtype).float()
    y.backward(dy)
    ya, idx = E.maxpool3x3s2_idx(to_act(E, x.detach(), dtype))
    assert torch.equal(back(ya), y.detach())
    dx = E.new_act(2, hw[0], hw[1], 64, dtype, "cuda")
    E.maxpool3x3s2_bwd(to_act(E, dy, dtype), idx, dx, False)
    assert rel(back(dx), x.grad) < (1e-6 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("sizes", [((6, 6), (40, 80)), ((1, 1), (8, 12)), ((8, 12), (16, 24)), ((41, 60), (82, 120)), ((3, 3), (11, 30)), ((2, 3), (64, 96)),
                                   ((1, 1), (2, 2)), ((5, 7), (10, 14)), ((1, 9), (2, 18)), ((7, 1), (14, 2)),      # exact 2x, odd sizes (2x2 micro-tiles)
                                   ((2, 2), (40, 80)), ((6, 8), (41, 83)), ((3, 12), (30, 120)),                     # separable two-pass / gather fallback (W > 8)
                                   ((33, 70), (66, 140)), ((48, 49), (96, 98))])                                    # exact 2x, large: walk kernel, ragged last segment
def test_bilinear_backward(E, dtype, tol, sizes):
    """All three adjoint algorithms (exact-2x micro-tiles, separable two-pass, generic gather) against autograd of
    F.interpolate (= the reference's F.upsample / nn.Upsample), plus the accumulate flag."""
    (h, w), (ho, wo) = sizes
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 64, h, w, generator=g, requires_grad=True)
    y = F.interpolate(x, size=(ho, wo), mode="bilinear", align_corners=False)
    dy = torch.randn(y.shape, generator=g).to(dtype).float()
    y.backward(dy)
    dx = E.new_act(2, h, w, 64, dtype, "cuda")
    E.bilinear_bwd(to_act(E, dy, dtype), dx, False)
    assert rel(back(dx), x.grad) < tol
    E.bilinear_bwd(to_act(E, dy, dtype), dx, True)
    assert rel(back(dx), 2 * x.grad) < 2 * tol


def test_bilinear_backward_into_channel_slice(E):
    """dx is a channel slice of a wider buffer (zero-copy concat): neighbours must stay untouched."""
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 64, 6, 8, generator=g, requires_grad=True)
    y = F.interpolate(x, size=(12, 16), mode="bilinear", align_corners=False)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    wide = E.Act(torch.full((2, 6, 8, 192), 7.0, device="cuda"))
    E.bilinear_bwd(to_act(E, dy, torch.float32), wide.slice(64, 64), False)
    assert rel(back(wide.slice(64, 64)), x.grad) < 1e-5
    assert torch.all(wide.buf[..., :64] == 7.0) and torch.all(wide.buf[..., 128:] == 7.0)


@pytest.mark.parametrize("hw", [(2, 3), (10, 20), (1, 1), (5, 9)])
def test_bilinear_backward_single_channel_x32(E, hw):
    """The critics' nn.Upsample(scale_factor=32) adjoint on single-channel FP32 maps (cm/discriminator_model.py:46,61)."""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 1, *hw, generator=g, requires_grad=True)
    y = nn.Upsample(scale_factor=32, mode='bilinear')(x)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    dx = E.new_act(2, hw[0], hw[1], 1, torch.float32, "cuda")
    E.bilinear_bwd(to_act(E, dy, torch.float32), dx, False)
    assert rel(back(dx), x.grad) < 1e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("hw", [(8, 12), (40, 80), (11, 30), (5, 7)])
def test_pyramid_pool_backward(E, dtype, tol, hw):
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 128, *hw, generator=g, requires_grad=True)
    sizes = (1, 2, 3, 6)
    outs = [F.adaptive_avg_pool2d(x, s) for s in sizes]
    douts = [torch.randn(o.shape, generator=g).to(dtype).float() for o in outs]
    torch.autograd.backward(outs, douts)
    dpool = torch.cat([d.permute(0, 2, 3, 1).reshape(-1) for d in douts]).to(dtype).cuda()
    dx = E.new_act(2, hw[0], hw[1], 128, dtype, "cuda")
    E.pyramid_pool_bwd(dpool, sizes, dx, False)
    assert rel(back(dx), x.grad) < tol
    # accumulate flag
    E.pyramid_pool_bwd(dpool, sizes, dx, True)
    assert rel(back(dx), 2 * x.grad) < tol * 2
