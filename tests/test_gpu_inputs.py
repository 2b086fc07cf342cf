"""Device-side input pipeline (heatnet_pub_b200.inputs) against the loader restatement in oracle/inputs_oracle.py: FP32 outputs
bit-identical, BF16 = round-to-nearest of them, rectDropTensor incl. Python slice clamping, and the pipeline feeding the network
zero-copy gives the same logits as the reference-style NCHW FP32 tensors."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _frames(n, h, w, seed=0):
    rng = np.random.RandomState(seed)
    rgb = rng.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
    ir = rng.randint(20000, 27000, (n, h, w)).astype(np.uint16)          # below, inside and above the clip range
    ir[0, 0, :4] = [0, 21800, 25000, 65535]
    return rgb, ir


@pytest.mark.parametrize("shape", [(2, 33, 47), (1, 64, 96), (3, 7, 5)])
def test_prepare_rgb_ir_bit_exact(shape):
    from heatnet_pub_b200 import inputs as I
    from oracle import inputs_oracle as IO
    n, h, w = shape
    rgb, ir = _frames(n, h, w)
    want_rgb = torch.stack([IO.load_rgb(rgb[i]) for i in range(n)])
    want_ir = torch.stack([IO.load_ir(ir[i]) for i in range(n)])
    got_rgb = I.prepare_rgb(torch.from_numpy(rgb).cuda(), precision="fp32")
    got_ir = I.prepare_ir(torch.from_numpy(ir.astype(np.int32)).cuda(), precision="fp32")
    assert got_rgb.shape == (n, 3, h, w) and got_ir.shape == (n, 1, h, w)
    assert torch.equal(got_rgb.cpu(), want_rgb)                         # bit-exact: integer-indexed gathers
    assert torch.equal(got_ir.cpu(), want_ir)
    if hasattr(torch, "uint16"):
        got16 = I.prepare_ir(torch.from_numpy(ir).cuda(), precision="fp32")
        assert torch.equal(got16.cpu(), want_ir)
    b_rgb = I.prepare_rgb(torch.from_numpy(rgb).cuda(), precision="bf16")
    b_ir = I.prepare_ir(torch.from_numpy(ir.astype(np.int32)).cuda(), precision="bf16")
    assert torch.equal(b_rgb.float().cpu(), want_rgb.bfloat16().float()) and torch.equal(b_ir.float().cpu(), want_ir.bfloat16().float())
    # other statistics (db_stats in the reference) and clip ranges
    m, s = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    got = I.prepare_rgb(torch.from_numpy(rgb).cuda(), m, s, precision="fp32")
    assert torch.equal(got.cpu(), torch.stack([IO.load_rgb(rgb[i], m, s) for i in range(n)]))
    got = I.prepare_ir(torch.from_numpy(ir.astype(np.int32)).cuda(), 20800, 27000, precision="fp32")
    assert torch.equal(got.cpu(), torch.stack([IO.load_ir(ir[i], 20800, 27000) for i in range(n)]))


def test_rect_drop_like_reference():
    from heatnet_pub_b200 import inputs as I
    from oracle import inputs_oracle as IO
    rgb, ir = _frames(4, 40, 56, seed=1)
    params = torch.tensor([[5, 7, 10, 20], [30, 40, 100, 100], [0, 0, 0, 5], [39, 55, 1, 1]])     # inside, clipped, empty, last pixel
    for frames, prep, chans in ((rgb, I.prepare_rgb, 3), (ir.astype(np.int32), I.prepare_ir, 1)):
        t = prep(torch.from_numpy(frames).cuda(), precision="fp32")
        want = IO.rect_drop_tensor(t.cpu().clone(), params)
        out = I.rectDropTensor(t, params)
        assert out is t and torch.equal(t.cpu(), want)
    plain = torch.randn(4, 3, 40, 56).cuda()                        # a foreign NCHW tensor takes the indexing path
    want = IO.rect_drop_tensor(plain.cpu().clone(), params)
    assert torch.equal(I.rectDropTensor(plain, params).cpu(), want)


def test_pipeline_feeds_the_network_zero_copy():
    from heatnet_pub_b200 import engine as E, inputs as I, pspnet
    from oracle import heatnet_oracle as O
    from oracle import inputs_oracle as IO
    rgb, ir = _frames(2, 64, 96, seed=2)
    sd = O.recipe_fill(O.pspnet_state_dict(True, 4), seed=5)
    net = pspnet.PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50', in_channels=4,
                        pretrained=False, late_fusion=True)
    net.load_state_dict(sd)
    net = net.cuda().eval().set_precision("bf16")
    ref_rgb = torch.stack([IO.load_rgb(rgb[i]) for i in range(2)]).cuda()
    ref_ir = torch.stack([IO.load_ir(ir[i]) for i in range(2)]).cuda()
    with torch.no_grad():
        want, _, _ = net(ref_rgb, ref_ir)                                # the reference's way: normalised NCHW FP32 tensors
        a = I.prepare_rgb(torch.from_numpy(rgb).cuda())
        b = I.prepare_ir(torch.from_numpy(ir.astype(np.int32)).cuda())
        assert E.act_from_view(a) is not None and E.act_from_view(b) is not None      # consumed without a layout conversion
        l0 = E.launch_count
        got, _, _ = net(a, b)
        assert torch.equal(got, want)


@pytest.mark.parametrize("source", ["plain_nchw", "pipeline_view"])
def test_smart_augment_and_ir_scale_bit_exact(source):
    """smartAugment / ir_scale_aug (cm/train_trgb_segnet_conf.py:101-110,404-406) as one gather-multiply launch: bit-identical to
    the reference's per-class torch.where loop given the same `random` stream, incl. classes absent from the label map (their
    factors are never drawn) and the order the factors are consumed in."""
    import random
    from heatnet_pub_b200 import inputs as I
    from oracle import inputs_oracle as IO
    n, h, w = 3, 37, 53
    g = torch.Generator().manual_seed(4)
    label = torch.randint(0, 13, (n, h, w), generator=g)
    label[label == 5] = 7                                  # class 5 absent; 12 present only in image 0
    label[label == 12] = 0
    label[0, :3, :3] = 12
    if source == "plain_nchw":
        ir = (torch.rand(n, 1, h, w, generator=g) * 2 - 1)
        dev = ir.cuda()
    else:
        counts = torch.randint(21000, 26000, (n, h, w), generator=g, dtype=torch.int32)
        dev = I.prepare_ir(counts.cuda(), precision="fp32")
        ir = dev.cpu().contiguous()
    want = IO.smart_augment(ir.clone(), label, random.Random(99))
    got = I.smartAugment(dev, label.cuda(), rng=random.Random(99))
    assert got is dev
    assert torch.equal(got.cpu().contiguous(), want)
    # ir_scale_aug: one in-place launch, same FP32 product as `scale * ir_day`
    scale = random.Random(5).uniform(0.1, 1)
    assert torch.equal(I.ir_scale_aug(dev, scale).cpu().contiguous(), IO.ir_scale_aug(want, scale))
    # BF16 pipeline views round once per multiply
    if source == "pipeline_view":
        devb = I.prepare_ir(counts.cuda(), precision="bf16")
        base = devb.float().cpu().contiguous()
        gotb = I.smartAugment(devb, label.cuda(), rng=random.Random(99)).float().cpu().contiguous()
        wantb = IO.smart_augment(base.clone(), label, random.Random(99)).bfloat16().float()
        assert torch.equal(gotb, wantb)
