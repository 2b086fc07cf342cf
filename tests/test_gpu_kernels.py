"""Per-kernel parity on a B200: every C-ABI entry point against the oracle (plain-C / torch-functional CPU
restatements) on seeded inputs.  Integer work is bit-exact; FP32 kernels within 1e-4 relative (the north
star's FP32 bound, written next to each check); BF16 tensor-core convs within 2e-2 of the FP32 oracle and
within 1e-3 of an FP32 evaluation on the same BF16-rounded operands."""
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
    from heatnet_pub_b200 import engine
    from heatnet_pub_b200 import _lib
    _lib.require_device()
    return engine


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def to_act(E, t, dtype):
    return E.from_nchw(t.cuda(), dtype)


def back(act):
    return act.nchw().float().cpu()


# ------------------------------------------------------------------------------------------------ layout
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("c,coff,cbuf", [(13, 0, 16), (13, 2, 20), (16, 4, 24), (9, 3, 13), (12, 4, 16)])
def test_nchw_into_narrow_channel_slice_keeps_the_neighbours(E, dtype, c, coff, cbuf):
    """9..16-channel NCHW tensors (the logits and their gradient) enter the NHWC engine through a one-thread-per-pixel kernel whose
    vector stores must stay inside the C valid channels: the other channels of the buffer keep their contents."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, c, 11, 7, generator=g)
    buf = E.new_act(2, 11, 7, cbuf, dtype, "cuda")
    buf.buf.fill_(7.0)
    E.from_nchw(x.cuda(), dtype, buf.slice(coff, c))
    full = buf.nchw().float().cpu()
    assert torch.equal(full[:, coff:coff + c], x.to(dtype).float())
    assert torch.all(full[:, :coff] == 7.0) and torch.all(full[:, coff + c:] == 7.0)


@pytest.mark.parametrize("shape", [(2, 3, 17, 23), (1, 1, 8, 8), (2, 13, 9, 31), (1, 64, 5, 7), (2, 100, 33, 4), (2, 16, 7, 9), (1, 9, 4, 5)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layout_roundtrip(E, shape, dtype):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(shape, generator=g)
    a = to_act(E, x, dtype)
    want = x.to(dtype).float()
    assert torch.equal(back(a), want)
    assert torch.equal(E.to_nchw_f32(a).cpu(), want)


# ------------------------------------------------------------------------------------------------ convolution
CONV_CASES = [
    # cin, cout, k, stride, pad, dil, h, w, bias        (layer classes of SURVEY.md appendix A)
    (64, 64, 1, 1, 0, 1, 20, 24, False),      # layer1 conv1
    (64, 256, 1, 1, 0, 1, 20, 24, False),     # layer1 conv3
    (128, 128, 3, 1, 1, 1, 10, 12, False),    # 3x3 d1
    (256, 256, 3, 1, 2, 2, 10, 13, False),    # layer3 3x3 d2 (ragged width)
    (512, 512, 3, 1, 4, 4, 11, 12, False),    # layer4 3x3 d4 (ragged height)
    (128, 128, 3, 2, 1, 1, 21, 24, False),    # layer2.0 conv2 stride 2, odd height
    (256, 512, 1, 2, 0, 1, 21, 24, False),    # layer2.0 downsample stride 2
    (3, 64, 7, 2, 3, 1, 33, 40, False),       # RGB stem
    (1, 64, 7, 2, 3, 1, 33, 40, False),       # IR stem
    (4, 64, 7, 2, 3, 1, 32, 40, False),       # early-fusion stem
    (2048, 1024, 1, 1, 0, 1, 6, 9, True),     # deep 1x1 with bias (PSP bottleneck class)
    (1024, 256, 3, 1, 1, 1, 12, 16, True),    # up_1 conv
    (64, 13, 1, 1, 0, 1, 16, 24, True),       # final
    (13, 64, 4, 2, 1, 1, 32, 48, True),       # critic conv1 on logits
    (128, 64, 4, 2, 1, 1, 16, 24, True),      # critic conv1 on x1
    (512, 1, 4, 2, 1, 1, 4, 6, True),         # critic classifier
    (256, 512, 4, 2, 1, 1, 40, 80, True),     # critic conv4 class at a realistic size: strided tensor map, several pixel tiles
    (64, 128, 4, 2, 1, 1, 33, 47, True),      # odd input, ragged output tiles
    (13, 64, 1, 1, 0, 1, 16, 24, True),       # Cin <= 16 1x1 (the classifier's dgrad shape): CUDA-core small-K kernel, padded pixel stride
    (3, 100, 1, 1, 0, 1, 9, 11, True),        # small-K kernel, dense 3-channel pixels, Cout not a multiple of 8
    (16, 520, 1, 1, 0, 1, 5, 7, False),       # small-K kernel, Cin = 16, many output chunks
]


def _conv_case(case, seed=0):
    cin, cout, k, stride, pad, dil, h, w, bias = case
    g = torch.Generator().manual_seed(seed)
    conv = nn.Conv2d(cin, cout, k, stride, pad, dil, bias=bias)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (k * k * cin)) ** 0.5)
        if bias:
            conv.bias.copy_(torch.randn(cout, generator=g) * 0.1)
    x = torch.randn(2, cin, h, w, generator=g)
    return conv, x


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fp32_matches_oracle(E, case):
    conv, x = _conv_case(case)
    ref = F.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation)     # oracle primitive (CPU FP32)
    convg = conv.cuda()
    scale, shift = E.folded_affine(convg, None)
    y = E.conv2d(to_act(E, x, torch.float32), convg, scale, shift)
    assert rel(back(y), ref) < FP32_TOL


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_bf16_tensor_core_matches_oracle(E, case):
    conv, x = _conv_case(case)
    ref = F.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation)
    # same operands rounded to BF16, evaluated in FP32 on the CPU: isolates kernel bugs from rounding
    ref_r = F.conv2d(x.bfloat16().float(), conv.weight.detach().bfloat16().float(), conv.bias, conv.stride, conv.padding,
                     conv.dilation)
    convg = conv.cuda()
    scale, shift = E.folded_affine(convg, None)
    y32 = E.conv2d(to_act(E, x, torch.bfloat16), convg, scale, shift, out_dtype=torch.float32)
    assert rel(back(y32), ref_r) < 1e-3
    y16 = E.conv2d(to_act(E, x, torch.bfloat16), convg, scale, shift)
    assert y16.dtype == torch.bfloat16
    assert rel(back(y16), ref) < BF16_TOL


HALO_CASES = [
    # cin, cout, dil, h, w      3x3 stride-1 layers with narrow outputs: one halo patch per k-block (hn_conv_halo.cu)
    (64, 64, 1, 33, 40),       # ragged both ways, weights resident in shared memory
    (64, 64, 1, 16, 8),        # exactly one tile
    (256, 64, 1, 20, 24),      # 4 k-blocks, streamed weights
    (128, 128, 1, 18, 30),     # two Cout tiles possible
    (64, 64, 2, 21, 17),       # dilation 2
    (192, 64, 1, 7, 5),        # smaller than a tile
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv3x3_halo_patch_path(E, case):
    cin, cout, dil, h, w = case
    g = torch.Generator().manual_seed(11)
    conv = nn.Conv2d(cin, cout, 3, 1, dil, dil, bias=True)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (9 * cin)) ** 0.5)
        conv.bias.copy_(torch.randn(cout, generator=g) * 0.1)
    x = torch.randn(3, cin, h, w, generator=g)
    res = torch.randn(3, cout, h, w, generator=g)
    ref_r = F.relu(F.conv2d(x.bfloat16().float(), conv.weight.detach().bfloat16().float(), conv.bias, 1, dil, dil) + res.bfloat16().float())
    import copy
    convg = copy.deepcopy(conv).cuda()
    scale, shift = E.folded_affine(convg, None)
    y32 = E.conv2d(to_act(E, x, torch.bfloat16), convg, scale, shift, act=E.ACT_RELU, out_dtype=torch.float32)
    ref_nores = F.relu(F.conv2d(x.bfloat16().float(), conv.weight.detach().bfloat16().float(), conv.bias, 1, dil, dil))
    assert rel(back(y32), ref_nores) < 1e-3
    y16 = E.conv2d(to_act(E, x, torch.bfloat16), convg, scale, shift, residual=to_act(E, res, torch.bfloat16), act=E.ACT_RELU)
    assert rel(back(y16), ref_r) < 1e-2


PAIR_CASES = [
    # cin, cout, k, pad, dil, n, h, w          two-CTA (cta_group::2) launches: the M tiles are handed out in pairs
    (64, 64, 3, 1, 1, 1, 48, 8),       # halo kernel, 3 tiles: the last pair has a filler tile beyond the image
    (64, 64, 3, 1, 1, 3, 16, 8),       # halo kernel, one tile per image, odd image count: filler tile = image index n
    (256, 64, 3, 1, 1, 2, 40, 24),     # halo kernel, streamed weights (4 k-blocks x 3 ring stages), 9 tiles per image
    (128, 128, 3, 1, 1, 1, 48, 24),    # halo kernel, N = 128
    (64, 64, 3, 4, 4, 1, 24, 16),      # generic kernel (dilation 4), N = 64, 3 tiles
    (128, 128, 3, 4, 4, 3, 8, 16),     # generic kernel, N = 128, odd tile count
    (256, 256, 3, 4, 4, 1, 24, 16),    # generic kernel, N = 256 (the pair halves the staged weight rows), 3 tiles
    (1024, 256, 1, 0, 1, 1, 9, 43),    # flat 1x1 with a long reduction (16 k-blocks): pairs over a flattened, ragged pixel axis
    (192, 100, 3, 4, 4, 1, 24, 16),    # Cout not a multiple of the tile: the pair's second weight half is partly padding
    (64, 512, 1, 0, 1, 3, 80, 80),     # single CTAs, N = 256, 300 tiles on 148 CTAs: the residual prefetch crosses tile boundaries
    (576, 256, 1, 0, 1, 3, 80, 80),    # pairs, N = 256, 75 tile pairs on 74 clusters: one cluster runs two tiles (prefetch across them)
    (128, 320, 1, 0, 1, 2, 40, 41),    # N = 64 tail-free but Cout = 320 -> 256-wide tiles with a half-empty second Cout tile (chunks switched off)
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_cta_pairs_with_odd_tile_counts(E, case):
    """Layers that run on CTA pairs (hn_tc_ptx.cuh "CTA pairs": N <= 128 tiles, or >= 9 k-blocks on the generic kernel) with tile counts
    that leave the last pair a filler tile: output, residual add and the fused batch statistics must ignore it."""
    cin, cout, k, pad, dil, n, h, w = case
    g = torch.Generator().manual_seed(21)
    conv = nn.Conv2d(cin, cout, k, 1, pad, dil, bias=True)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (k * k * cin)) ** 0.5)
        conv.bias.copy_(torch.randn(cout, generator=g) * 0.1)
    x = torch.randn(n, cin, h, w, generator=g)
    res = torch.randn(n, cout, h, w, generator=g)
    pre = F.conv2d(x.bfloat16().float(), conv.weight.detach().bfloat16().float(), conv.bias, 1, pad, dil)
    convg = conv.cuda()
    scale, shift = E.folded_affine(convg, None)
    xa = to_act(E, x, torch.bfloat16)
    ld = (cout + 7) // 8 * 8
    # BF16 output + residual + ReLU through the TMA epilogue
    y = E.conv2d(xa, convg, scale, shift, residual=to_act(E, res, torch.bfloat16), act=E.ACT_RELU, out=E.new_act(n, h, w, cout, torch.bfloat16, "cuda", ld=ld))
    assert rel(back(y), F.relu(pre + res.bfloat16().float())) < 1e-2
    # FP32 output with fused statistics: exactly the sums of what was stored
    sums = torch.zeros((2, cout), dtype=torch.float64, device="cuda")
    y32 = E.conv2d(xa, convg, scale, shift, None, E.ACT_NONE, out=E.new_act(n, h, w, cout, torch.float32, "cuda", ld=ld), stats=sums)
    assert rel(back(y32), pre) < 1e-3
    stored = y32.nchw().double()
    want = torch.stack([stored.sum((0, 2, 3)), (stored * stored).sum((0, 2, 3))])
    assert ((sums - want).abs() / want.abs().clamp_min(1e-3 * want.abs().max())).max().item() < 2e-5
    # an unaligned output view (pixel stride not a multiple of 16 bytes) takes the per-thread store path: the filler tile must not write
    if cout % 8 == 0:
        guard = E.new_act(n, h, w, cout + 3, torch.bfloat16, "cuda")
        guard.buf.fill_(7.0)
        yv = E.conv2d(xa, convg, scale, shift, out=guard.slice(1, cout))
        assert rel(back(yv), pre) < 1e-2
        full = guard.nchw().float()
        assert torch.all(full[:, 0] == 7.0) and torch.all(full[:, cout + 1:] == 7.0)


@pytest.mark.parametrize("case", [(64, 64, 9, 11), (64, 64, 16, 8), (256, 64, 10, 12), (1024, 256, 6, 9), (128, 64, 1, 1)])
def test_upsample2x_conv3x3_fused(E, case):
    """PSPUpsample's bilinear 2x + 3x3 conv in one kernel (the upsampled tensor only ever exists as shared-memory halo
    patches) == F.interpolate(align_corners=False) -> conv2d on the same BF16-rounded intermediate."""
    cin, cout, h, w = case
    g = torch.Generator().manual_seed(12)
    conv = nn.Conv2d(cin, cout, 3, 1, 1, bias=True)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (9 * cin)) ** 0.5)
        conv.bias.copy_(torch.randn(cout, generator=g) * 0.1)
    x = torch.randn(2, cin, h, w, generator=g)
    up = F.interpolate(x.bfloat16().float(), size=(2 * h, 2 * w), mode="bilinear", align_corners=False).bfloat16().float()
    ref = F.leaky_relu(F.conv2d(up, conv.weight.detach().bfloat16().float(), conv.bias, 1, 1), 0.25)
    import copy
    convg = copy.deepcopy(conv).cuda()
    scale, shift = E.folded_affine(convg, None)
    xa = to_act(E, x, torch.bfloat16)
    y = E.upconv3x3(xa, convg, scale, shift, E.ACT_LEAKY, slope=0.25, out_dtype=torch.float32)
    assert (y.n, y.h, y.w, y.c) == (2, 2 * h, 2 * w, cout)
    assert rel(back(y), ref) < 1e-3
    # and identical (up to FP32 summation order) to the unfused path
    y2 = E.conv2d(E.bilinear(xa, 2 * h, 2 * w), convg, scale, shift, act=E.ACT_LEAKY, slope=0.25, out_dtype=torch.float32)
    assert rel(back(y), back(y2)) < 1e-5


@pytest.mark.parametrize("case", [(64, 13, 16, 8), (64, 13, 21, 30), (128, 16, 9, 11), (64, 1, 5, 3), (64, 13, 80, 72)])
def test_conv3x3_bn_prelu_classifier_fused(E, case):
    """up_3.conv -> BN(eval) -> PReLU -> final 1x1 (cm/models/pspnet.py:72-75) in one kernel: NCHW FP32 logits equal
    conv -> BN -> PReLU -> 1x1 conv of the oracle on the BF16-rounded operands (the classifier is a second tcgen05 MMA on
    the BF16 activation tile with BF16 weights, FP32 accumulation -- the numerics of the two-launch path)."""
    cin, ncls, h, w = case
    g = torch.Generator().manual_seed(13)
    conv = nn.Conv2d(cin, 64, 3, 1, 1, bias=True)
    bn = nn.BatchNorm2d(64)
    prelu = nn.PReLU()
    head = nn.Conv2d(64, ncls, 1)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (9 * cin)) ** 0.5)
        bn.weight.copy_(torch.rand(64, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(64, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(64, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(64, generator=g) + 0.5)
        prelu.weight.fill_(0.3)
        head.weight.copy_(torch.randn(head.weight.shape, generator=g) * 0.125)
        head.bias.copy_(torch.randn(ncls, generator=g) * 0.1)
    bn.eval()
    x = torch.randn(3, cin, h, w, generator=g)
    with torch.no_grad():
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        wfold = (conv.weight * scale[:, None, None, None]).bfloat16().float()
        z = F.conv2d(x.bfloat16().float(), wfold, None, 1, 1) + (bn.bias + (conv.bias - bn.running_mean) * scale)[None, :, None, None]
        ref = F.conv2d(F.prelu(z, prelu.weight).bfloat16().float(), head.weight.bfloat16().float(), head.bias)
        ref_exact = head(F.prelu(z, prelu.weight))
    import copy
    convg, bng, prelug, headg = (copy.deepcopy(m).cuda() for m in (conv, bn, prelu, head))
    xa = to_act(E, x, torch.bfloat16)
    assert E.conv3x3_head_ok(xa, convg, bng, headg)
    out = E.conv3x3_head(xa, convg, bng, headg, E.ACT_LEAKY, slope_ptr=prelug.weight)
    assert out.shape == (3, ncls, h, w) and out.dtype == torch.float32 and out.is_contiguous()
    # 3e-3: an activation that sits on a BF16 rounding boundary may round the other way than in the CPU evaluation
    assert rel(out.cpu(), ref) < 3e-3
    assert rel(out.cpu(), ref_exact) < BF16_TOL
    # and the unfused pair of launches agrees (same roundings, different summation order)
    y = E.conv_bn_act(xa, convg, bng, E.ACT_LEAKY, slope_ptr=prelug.weight)
    lo = E.new_act(3, h, w, ncls, torch.float32, "cuda", ld=(ncls + 7) // 8 * 8)
    E.conv_bn_act(y, headg, None, out=lo)
    assert rel(out.cpu(), back(lo)) < 3e-3


@pytest.mark.parametrize("cin,h,w", [(3, 33, 40), (1, 33, 40), (4, 32, 40), (3, 65, 257), (1, 7, 9)])
def test_stem7x7_overlapping_window_path(E, cin, h, w):
    """7x7 s2 p3 stem -> BN(eval) -> ReLU without im2col (padded 4-channel image + strided tensor map) vs the oracle
    primitive on BF16-rounded operands; also the FP32-output variant the train-mode BN path uses."""
    g = torch.Generator().manual_seed(21)
    conv = nn.Conv2d(cin, 64, 7, 2, 3, bias=False)
    bn = nn.BatchNorm2d(64)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (49 * cin)) ** 0.5)
        bn.weight.copy_(torch.rand(64, generator=g) + 0.5); bn.bias.copy_(torch.randn(64, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(64, generator=g) * 0.1); bn.running_var.copy_(torch.rand(64, generator=g) + 0.5)
    bn.eval()
    x = torch.rand(2, cin, h, w, generator=g) * 2 - 1
    with torch.no_grad():
        ref = F.relu(bn(conv(x)))
        ref_raw = F.conv2d(x.bfloat16().float(), conv.weight.bfloat16().float(), None, 2, 3)
    import copy
    convg, bng = copy.deepcopy(conv).cuda(), copy.deepcopy(bn).cuda().eval()
    xa = to_act(E, x, torch.bfloat16)
    assert E.stem_ok(xa, convg)
    y = E.conv_bn_act(xa, convg, bng, E.ACT_RELU)
    assert (y.h, y.w) == (ref.shape[2], ref.shape[3])
    assert rel(back(y), ref) < BF16_TOL
    wp, shift = E.packed_stem_weight(convg, None)
    raw = E.stem_conv(xa, convg, wp, shift, out_dtype=torch.float32)
    assert rel(back(raw), ref_raw) < 1e-3


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
def test_conv_fused_epilogue_bn_residual_relu(E, dtype, tol):
    """conv -> BN(eval) -> += residual -> ReLU in one launch == cm/models/extractors.py:96-101."""
    g = torch.Generator().manual_seed(1)
    conv = nn.Conv2d(128, 512, 1, bias=False)
    bn = nn.BatchNorm2d(512)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.1)
        bn.weight.copy_(torch.rand(512, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(512, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(512, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(512, generator=g) + 0.5)
    bn.eval()
    x = torch.randn(2, 128, 9, 14, generator=g)
    res = torch.randn(2, 512, 9, 14, generator=g)
    with torch.no_grad():
        ref = F.relu(bn(conv(x)) + res)
    import copy
    convg, bng = copy.deepcopy(conv).cuda(), copy.deepcopy(bn).cuda()
    y = E.conv_bn_act(to_act(E, x, dtype), convg, bng, E.ACT_RELU, residual=to_act(E, res, dtype))
    assert rel(back(y), ref) < tol
    # PReLU slope read from device memory
    prelu = nn.PReLU().cuda()
    with torch.no_grad():
        prelu.weight.fill_(0.3)
        ref2 = F.prelu(bn(conv(x)), torch.tensor([0.3]))
    y2 = E.conv_bn_act(to_act(E, x, dtype), convg, bng, E.ACT_LEAKY, slope_ptr=prelu.weight)
    assert rel(back(y2), ref2) < tol


STAT_CASES = [
    # cin, cout, k, stride, pad, dil, h, w, bias     -- one case per Cout-tile width / kernel family of the BF16 engine
    (64, 32, 1, 1, 0, 1, 9, 14, True),         # BLOCK_N = 32 (half of a staging row)
    (64, 64, 1, 1, 0, 1, 20, 24, False),       # BLOCK_N = 64, alternating epilogue groups
    (64, 256, 1, 1, 0, 1, 20, 24, False),      # wide tile, two 64-channel chunks per warp
    (128, 512, 1, 1, 0, 1, 33, 40, False),     # several Cout tiles per pixel tile, ragged last pixel tile
    (256, 256, 3, 1, 2, 2, 10, 13, False),     # implicit 3x3 with ragged spatial tiles (rows beyond the image must not count)
    (64, 64, 3, 1, 1, 1, 24, 40, True),        # halo-patch kernel (3x3, narrow output)
    (128, 128, 3, 2, 1, 1, 21, 24, False),     # strided: im2col + flat path
    (3, 64, 7, 2, 3, 1, 33, 40, False),        # dense-row stem
    (256, 100, 1, 1, 0, 1, 12, 16, True),      # Cout not a multiple of the tile: padded channels stay out of the sums
]


@pytest.mark.parametrize("case", STAT_CASES)
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_conv_fused_batch_statistics(E, case, out_dtype):
    """hn_epilogue.stat_sum / stat_sqsum: the per-channel sum and sum of squares the conv epilogue accumulates equal those of
    the tensor it stored (FP32 output: of the FP32 values; BF16 output: of the ROUNDED values), over exactly the valid pixels."""
    cin, cout, k, stride, pad, dil, h, w, bias = case
    g = torch.Generator().manual_seed(11)
    conv = nn.Conv2d(cin, cout, k, stride, pad, dil, bias=bias)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (cin * k * k)) ** 0.5)
        if bias:
            conv.bias.copy_(torch.randn(cout, generator=g))
    conv = conv.cuda()
    x = torch.randn(3, cin, h, w, generator=g) + 0.5
    xa = to_act(E, x, torch.bfloat16)
    sums = torch.zeros((2, cout), dtype=torch.float64, device="cuda")
    scale, shift = E.folded_affine(conv, None)
    if E.stem_ok(xa, conv):
        wp, shift = E.packed_stem_weight(conv, None)
        y = E.stem_conv(xa, conv, wp, shift, out_dtype=out_dtype, stats=sums)
    else:
        ho, wo = E.conv_out_hw(h, w, conv)
        out = E.new_act(3, ho, wo, cout, out_dtype, "cuda", ld=(cout + 7) // 8 * 8)        # 16-byte aligned pixel stride
        y = E.conv2d(xa, conv, scale, shift, None, E.ACT_NONE, out=out, stats=sums)
    stored = y.nchw().double()                              # exactly what the kernel wrote
    want = torch.stack([stored.sum((0, 2, 3)), (stored * stored).sum((0, 2, 3))])
    err = ((sums - want).abs() / want.abs().clamp_min(1e-3 * want.abs().max())).max().item()
    assert err < 2e-5, err                                   # FP32 partial sums per 32-pixel block, FP64 across blocks
    with torch.no_grad():
        ref = F.conv2d(x.cuda().bfloat16().float(), conv.weight.bfloat16().float(), conv.bias, stride, pad, dil)
    assert rel(y.nchw().float(), ref) < 1e-2


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
def test_conv_into_channel_slice_is_the_concat(E, dtype, tol):
    """Two producers writing channel halves of one buffer == torch.cat(dim=1) (cm/models/extractors.py:192)."""
    g = torch.Generator().manual_seed(2)
    c1, c2 = nn.Conv2d(64, 128, 1, bias=False), nn.Conv2d(64, 128, 3, padding=1, bias=False)
    xa, xb = torch.randn(1, 64, 10, 12, generator=g), torch.randn(1, 64, 10, 12, generator=g)
    with torch.no_grad():
        ref = torch.cat([c1(xa), c2(xb)], 1)
    cat = E.new_act(1, 10, 12, 256, dtype, "cuda")
    E.conv2d(to_act(E, xa, dtype), c1.cuda(), out=cat.slice(0, 128))
    E.conv2d(to_act(E, xb, dtype), c2.cuda(), out=cat.slice(128, 128))
    assert rel(back(cat), ref) < tol
    # and a conv reading a channel slice
    c3 = nn.Conv2d(128, 64, 1, bias=False)
    with torch.no_grad():
        ref3 = c3(ref[:, 128:])
    y3 = E.conv2d(cat.slice(128, 128), c3.cuda())
    assert rel(back(y3), ref3) < (tol if dtype == torch.float32 else 3e-2)


def test_conv_kernel_larger_than_input_raises(E):
    conv = nn.Conv2d(64, 64, 4, 2, 1).cuda()
    with pytest.raises(RuntimeError, match="Kernel size can't be greater than actual input size"):
        E.conv2d(E.new_act(1, 1, 1, 64, torch.bfloat16, "cuda"), conv)


def test_weight_cache_follows_parameter_updates(E):
    conv = nn.Conv2d(64, 64, 1, bias=False).cuda()
    x = torch.randn(1, 64, 8, 8)
    y0 = back(E.conv2d(to_act(E, x, torch.float32), conv))
    with torch.no_grad():
        conv.weight.mul_(2.0)          # in-place update bumps the version counter -> re-pack
    y1 = back(E.conv2d(to_act(E, x, torch.float32), conv))
    assert rel(y1, 2 * y0) < 1e-6


# ------------------------------------------------------------------------------------------------ bandwidth kernels
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("hw", [(17, 24), (16, 16), (7, 9)])
def test_maxpool(E, dtype, hw):
    x = torch.randn(2, 64, *hw, generator=torch.Generator().manual_seed(3))
    ref = F.max_pool2d(x.to(dtype).float(), 3, 2, 1)
    assert torch.equal(back(E.maxpool3x3s2(to_act(E, x, dtype))), ref)      # max is exact in any dtype


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("hw", [(8, 12), (40, 80), (11, 30), (5, 7)])
def test_pyramid_pool(E, dtype, tol, hw):
    x = torch.randn(2, 128, *hw, generator=torch.Generator().manual_seed(4))
    outs = E.pyramid_pool(to_act(E, x, dtype), (1, 2, 3, 6))
    for s, o in zip((1, 2, 3, 6), outs):
        ref = F.adaptive_avg_pool2d(x.to(dtype).float(), s)                  # overlapping bins when hw % s != 0
        assert (o.n, o.h, o.w, o.c) == (2, s, s, 128)
        assert rel(back(o), ref) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("sizes", [((6, 6), (40, 80)), ((1, 1), (8, 12)), ((8, 12), (16, 24)), ((41, 60), (82, 120)), ((3, 3), (11, 30))])
def test_bilinear(E, dtype, tol, sizes):
    (h, w), (ho, wo) = sizes
    x = torch.randn(2, 64, h, w, generator=torch.Generator().manual_seed(5))
    ref = F.interpolate(x.to(dtype).float(), size=(ho, wo), mode="bilinear", align_corners=False)
    assert rel(back(E.bilinear(to_act(E, x, dtype), ho, wo)), ref) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("out_hw", [(40, 80), (41, 83), (7, 9)])      # segmented (>= 4x magnification) and generic kernels
def test_bilinear_sum_of_pyramid_priors(E, dtype, tol, out_hw):
    """sum_s upsample(prior_s) of the PSP head (cm/models/pspnet.py:23-24 after the 1x1 projection) in one pass."""
    g = torch.Generator().manual_seed(11)
    ho, wo = out_hw
    xs = [torch.randn(2, 136, s, s, generator=g) for s in (1, 2, 3, 6)]
    ref = sum(F.interpolate(x.to(dtype).float(), size=(ho, wo), mode="bilinear", align_corners=False) for x in xs)
    y = E.bilinear_sum([to_act(E, x, dtype) for x in xs], ho, wo)
    assert rel(back(y), ref) < tol


def test_bilinear_x32_single_channel_fp32(E):
    """nn.Upsample(scale_factor=32, mode='bilinear') on the critics' 1-channel map (cm/discriminator_model.py:47)."""
    x = torch.randn(2, 1, 2, 3, generator=torch.Generator().manual_seed(6))
    ref = nn.Upsample(scale_factor=32, mode='bilinear')(x)
    y = E.bilinear(to_act(E, x, torch.float32), 64, 96)
    assert rel(back(y), ref) < 1e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
def test_batchnorm_train_mode(E, dtype, tol):
    """conv -> BN(train) -> ReLU: batch statistics, running-stat update (momentum .1, unbiased var),
    num_batches_tracked -- appendix B.3."""
    g = torch.Generator().manual_seed(7)
    conv = nn.Conv2d(64, 128, 3, padding=1, bias=True)
    bn = nn.BatchNorm2d(128)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(128, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(128, generator=g) * 0.1)
    x = torch.randn(3, 64, 10, 14, generator=g) + 0.5
    import copy
    convg, bng = copy.deepcopy(conv).cuda(), copy.deepcopy(bn).cuda()
    bn.train()
    with torch.no_grad():
        ref = F.relu(bn(conv(x)))
    bng.train()
    y = E.conv_bn_act(to_act(E, x, dtype), convg, bng, E.ACT_RELU)
    assert rel(back(y), ref) < tol
    st = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel(bng.running_mean.cpu(), bn.running_mean) < st
    assert rel(bng.running_var.cpu(), bn.running_var) < st
    assert int(bng.num_batches_tracked) == 1
    # eval right after must pick the updated running statistics up (cache invalidation)
    bn.eval(); bng.eval()
    with torch.no_grad():
        ref_e = F.relu(bn(conv(x)))
    y_e = E.conv_bn_act(to_act(E, x, dtype), convg, bng, E.ACT_RELU)
    assert rel(back(y_e), ref_e) < tol


# ------------------------------------------------------------------------------------------------ confusion matrix
def test_confusion_bit_exact_labels_and_scores(golden_dir):
    import os
    from heatnet_pub_b200 import iou_eval
    from oracle.iou_oracle import IoUOracle
    g = np.load(os.path.join(golden_dir, "iou_golden.npz"))
    m = iou_eval.IoU(14, False, [12, 13])
    m.add(torch.from_numpy(g["pred"][:3]).cuda(), torch.from_numpy(g["tgt"][:3]).cuda())
    assert np.array_equal(m.conf_metric.conf, g["conf_after_add1"]) and m.conf_metric.conf.dtype == np.int32
    iou, miou = m.value()
    assert np.array_equal(iou, g["iou1"], equal_nan=True) and miou == g["miou1"]
    assert np.array_equal(m.conf_metric.conf, g["conf_after_value1"])
    m.add(torch.from_numpy(g["pred"][3:]), torch.from_numpy(g["tgt"][3:]))          # CPU tensors take the same kernel
    assert np.array_equal(m.conf_metric.conf, g["conf_after_add2"])
    iou, miou = m.value()
    assert np.array_equal(iou, g["iou2"], equal_nan=True) and miou == g["miou2"]
    m2 = iou_eval.IoU(14, False, None)
    m2.add(torch.from_numpy(g["scores"]).cuda(), torch.from_numpy(g["tgt_s"]).cuda())   # fused argmax, ties -> first
    assert np.array_equal(m2.conf_metric.conf, g["conf_scores"])
    iou, miou = m2.value()
    assert np.array_equal(iou, g["iou3"], equal_nan=True) and miou == g["miou3"]
    m3 = iou_eval.IoU(14, True, 13)
    m3.add(torch.from_numpy(g["pred"]).cuda(), torch.from_numpy(g["tgt"]).cuda())
    assert np.array_equal(m3.conf_metric.value(), g["conf_normalized"])
    from heatnet_pub_b200 import utils
    got = utils.calculate_ious(torch.from_numpy(g["pred"]).cuda(), torch.from_numpy(g["tgt"]).cuda(), 13)
    assert np.array_equal(got, g["calc_ious"], equal_nan=True)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 255, 4097, 320 * 640 * 3 + 1])
def test_confusion_ragged_sizes_vs_oracle(n):
    from heatnet_pub_b200 import iou_eval
    from oracle import c_oracle
    rng = np.random.RandomState(n)
    pred, tgt = rng.randint(0, 14, n).astype(np.int64), rng.randint(0, 14, n).astype(np.int64)
    cm = iou_eval.ConfusionMatrix(14)
    cm.add(torch.from_numpy(pred).cuda(), torch.from_numpy(tgt).cuda())
    want = c_oracle.confusion(pred, tgt, 14) if n else np.zeros((14, 14), np.int32)
    assert np.array_equal(cm.conf, want)


def test_confusion_skewed_labels_and_large_k():
    """Real label maps are dominated by a few classes: every lane of a warp hits the same bin."""
    from heatnet_pub_b200 import iou_eval
    from oracle import c_oracle
    n = 1 << 22
    pred, tgt = np.full(n, 3, np.int64), np.full(n, 3, np.int64)
    pred[::1000] = 7
    cm = iou_eval.ConfusionMatrix(14)
    cm.add(torch.from_numpy(pred).cuda(), torch.from_numpy(tgt).cuda())
    assert np.array_equal(cm.conf, c_oracle.confusion(pred, tgt, 14))
    rng = np.random.RandomState(1)
    pred, tgt = rng.randint(0, 32, 100000).astype(np.int64), rng.randint(0, 32, 100000).astype(np.int64)
    cm = iou_eval.ConfusionMatrix(32)
    cm.add(torch.from_numpy(pred).cuda(), torch.from_numpy(tgt).cuda())
    assert np.array_equal(cm.conf, c_oracle.confusion(pred, tgt, 32))


def test_confusion_range_asserts_like_reference():
    from heatnet_pub_b200 import iou_eval
    cm = iou_eval.ConfusionMatrix(14)
    ok = torch.zeros(10, dtype=torch.int64).cuda()
    bad = ok.clone(); bad[3] = 14
    with pytest.raises(AssertionError, match="predicted values are not between 0 and k-1"):
        cm.add(bad, ok)
    neg = ok.clone(); neg[5] = -1
    with pytest.raises(AssertionError, match="target values are not between 0 and k-1"):
        cm.add(ok, neg)
    with pytest.raises(AssertionError, match="number of targets and predicted outputs do not match"):
        cm.add(ok, ok[:5])
    big = torch.zeros(50001, dtype=torch.int64).cuda()          # the TMA-staged kernel (n >= 4096), odd length
    bad = big.clone(); bad[40000] = 14
    with pytest.raises(AssertionError, match="predicted values are not between 0 and k-1"):
        cm.add(bad, big)
    neg = big.clone(); neg[50000] = -(1 << 40)
    with pytest.raises(AssertionError, match="target values are not between 0 and k-1"):
        cm.add(big, neg)
    m = iou_eval.IoU(14)
    with pytest.raises(AssertionError, match="predictions must be of dimension"):
        m.add(torch.zeros(2, 3).cuda(), torch.zeros(2, 3).cuda())


def test_confusion_checksum_at_full_size():
    """Config 5 property at the reference's map size: the matrix sums to the pixel count, row sums are the
    target histogram and column sums the prediction histogram (10 x 500 maps of 320x640 would take the CPU
    oracle minutes; 40 maps are checked bit-exactly, the rest through the checksums)."""
    from heatnet_pub_b200 import iou_eval
    from oracle import c_oracle
    g = torch.Generator(device="cuda").manual_seed(1203412412)
    cm = iou_eval.ConfusionMatrix(14)
    tot_t, tot_p = torch.zeros(14, dtype=torch.int64), torch.zeros(14, dtype=torch.int64)
    n_maps = 0
    for chunk in range(4):
        pred = torch.randint(0, 14, (100, 320, 640), generator=g, device="cuda")
        tgt = torch.randint(0, 14, (100, 320, 640), generator=g, device="cuda")
        cm.add(pred.view(-1), tgt.view(-1))
        tot_t += torch.bincount(tgt.view(-1), minlength=14).cpu()
        tot_p += torch.bincount(pred.view(-1), minlength=14).cpu()
        n_maps += 100
        if chunk == 0:
            want = c_oracle.confusion(pred[:40].cpu().numpy(), tgt[:40].cpu().numpy(), 14)
            cm40 = iou_eval.ConfusionMatrix(14)
            cm40.add(pred[:40].reshape(-1), tgt[:40].reshape(-1))
            assert np.array_equal(cm40.conf, want)
    conf = cm.conf.astype(np.int64)
    assert conf.sum() == n_maps * 320 * 640
    assert np.array_equal(conf.sum(1), tot_t.numpy()) and np.array_equal(conf.sum(0), tot_p.numpy())
