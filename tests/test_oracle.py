"""Pins the oracle (oracle/*.py, oracle/heatnet_oracle.c) against the golden fixtures the REFERENCE
produced (tests/golden/make_golden.py), and the torch-functional restatement against the plain-C one.
CPU only."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import c_oracle, heatnet_oracle as O
from oracle.iou_oracle import ConfusionMatrixOracle, IoUOracle, calculate_ious_oracle


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


# ---------------------------------------------------------------------------------- iou_eval
def test_iou_oracle_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "iou_golden.npz")
    m = IoUOracle(14, False, [12, 13])
    m.add(g["pred"][:3], g["tgt"][:3])
    assert np.array_equal(m.conf_metric.conf, g["conf_after_add1"])
    assert m.conf_metric.conf.dtype == np.int32
    iou, miou = m.value()
    assert np.array_equal(iou, g["iou1"], equal_nan=True) and miou == g["miou1"]
    assert np.array_equal(m.conf_metric.conf, g["conf_after_value1"])      # in-place mutation quirk
    m.add(g["pred"][3:], g["tgt"][3:])
    assert np.array_equal(m.conf_metric.conf, g["conf_after_add2"])
    iou, miou = m.value()
    assert np.array_equal(iou, g["iou2"], equal_nan=True) and miou == g["miou2"]


def test_iou_oracle_scores_path_and_normalized(golden_dir):
    g = _load(golden_dir, "iou_golden.npz")
    m = IoUOracle(14, False, None)
    m.add(g["scores"], g["tgt_s"])
    assert np.array_equal(m.conf_metric.conf, g["conf_scores"])
    iou, miou = m.value()
    assert np.array_equal(iou, g["iou3"], equal_nan=True) and miou == g["miou3"]
    assert np.isnan(iou).any()
    m3 = IoUOracle(14, True, 13)
    m3.add(g["pred"], g["tgt"])
    assert np.array_equal(m3.conf_metric.value(), g["conf_normalized"])
    assert np.array_equal(calculate_ious_oracle(g["pred"], g["tgt"], 13), g["calc_ious"], equal_nan=True)


def test_c_confusion_and_argmax_match_numpy_oracle(golden_dir):
    g = _load(golden_dir, "iou_golden.npz")
    conf = c_oracle.confusion(g["pred"][:3], g["tgt"][:3], 14)
    assert np.array_equal(conf, g["conf_after_add1"])
    am = c_oracle.argmax(g["scores"])
    assert np.array_equal(am, np.argmax(g["scores"], 1))
    conf2 = c_oracle.confusion(am, g["tgt_s"], 14)
    assert np.array_equal(conf2, g["conf_scores"])
    with pytest.raises(AssertionError):
        c_oracle.confusion(np.array([14]), np.array([0]), 14)
    with pytest.raises(AssertionError):
        ConfusionMatrixOracle(14).add(np.array([3]), np.array([-1]))


def test_int32_accumulator_wraps_like_numpy():
    m = ConfusionMatrixOracle(2)
    m.conf[1, 1] = np.int32(2 ** 31 - 1)
    m.add(np.array([1]), np.array([1]))
    assert m.conf[1, 1] == np.int32(-2 ** 31)
    conf = np.zeros((2, 2), np.int32)
    conf[1, 1] = 2 ** 31 - 1
    c_oracle.confusion(np.array([1]), np.array([1]), 2, conf)
    assert conf[1, 1] == np.int32(-2 ** 31)


# ---------------------------------------------------------------------------------- primitives: torch port == plain C
def test_c_primitives_match_torch_semantics():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 13, 17, generator=g)
    for (k, s, p, d) in [(1, 1, 0, 1), (3, 1, 1, 1), (3, 1, 2, 2), (3, 1, 4, 4), (3, 2, 1, 1), (7, 2, 3, 1), (4, 2, 1, 1), (1, 2, 0, 1)]:
        w = torch.randn(6, 5, k, k, generator=g)
        b = torch.randn(6, generator=g)
        ref = F.conv2d(x, w, b, s, p, d).numpy()
        got = c_oracle.conv2d(x.numpy(), w.numpy(), b.numpy(), s, p, d)
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-4)
    np.testing.assert_array_equal(c_oracle.maxpool3x3s2(x.numpy()), F.max_pool2d(x, 3, 2, 1).numpy())
    for s in (1, 2, 3, 6):
        np.testing.assert_allclose(c_oracle.adaptive_avgpool(x.numpy(), s), F.adaptive_avg_pool2d(x, s).numpy(), rtol=1e-5, atol=1e-6)
    for size in [(26, 34), (13, 17), (40, 80), (5, 3)]:
        np.testing.assert_allclose(c_oracle.bilinear(x.numpy(), *size), O.upsample_bilinear(x, size).numpy(), rtol=1e-5, atol=1e-5)
    x6 = torch.randn(1, 4, 2, 3, generator=g)
    np.testing.assert_allclose(c_oracle.bilinear(x6.numpy(), 64, 96),
                               torch.nn.Upsample(scale_factor=32, mode='bilinear')(x6).numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(c_oracle.leaky(x.numpy(), 0.2), F.leaky_relu(x, 0.2).numpy())
    np.testing.assert_array_equal(c_oracle.leaky(x.numpy(), 0.25), F.prelu(x, torch.tensor([0.25])).numpy())
    # batch norm, train + eval
    gamma, beta = torch.rand(5, generator=g) + 0.5, torch.randn(5, generator=g)
    rm, rv = torch.randn(5, generator=g), torch.rand(5, generator=g) + 0.5
    rm_c, rv_c = rm.clone().numpy(), rv.clone().numpy()
    rm_t, rv_t = rm.clone(), rv.clone()
    y_t = F.batch_norm(x, rm_t, rv_t, gamma, beta, True, 0.1, 1e-5)
    y_c = c_oracle.batchnorm(x.numpy(), gamma.numpy(), beta.numpy(), rm_c, rv_c, True)
    np.testing.assert_allclose(y_c, y_t.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rm_c, rm_t.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rv_c, rv_t.numpy(), rtol=1e-5, atol=1e-6)
    y_t = F.batch_norm(x, rm_t, rv_t, gamma, beta, False, 0.1, 1e-5)
    y_c = c_oracle.batchnorm(x.numpy(), gamma.numpy(), beta.numpy(), rm_c, rv_c, False)
    np.testing.assert_allclose(y_c, y_t.numpy(), rtol=1e-4, atol=1e-5)
    # cross entropy with and without ignore_index
    logits = torch.randn(2, 13, 6, 7, generator=g)
    labels = torch.randint(0, 13, (2, 6, 7), generator=g)
    assert abs(c_oracle.cross_entropy(logits.numpy(), labels.numpy()) - F.cross_entropy(logits, labels).item()) < 1e-5
    labels14 = torch.randint(0, 14, (2, 6, 7), generator=g)
    logits14 = torch.randn(2, 14, 6, 7, generator=g)
    assert abs(c_oracle.cross_entropy(logits14.numpy(), labels14.numpy(), 13)
               - F.cross_entropy(logits14, labels14, ignore_index=13).item()) < 1e-5


# ---------------------------------------------------------------------------------- network restatement == reference
def _late_sd():
    return O.recipe_fill(O.pspnet_state_dict(True, 4), seed=0)


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_pspnet_late_oracle_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "pspnet_late_golden.npz")
    sd = _late_sd()
    assert list(sd.keys()) == [str(k) for k in g["state_dict_keys"]]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g["state_dict_shapes"]]
    assert len(sd) == 488
    rgb, ir = O.synthetic_inputs(2, 64, 96)
    assert np.array_equal(rgb.numpy(), g["rgb"]) and np.array_equal(ir.numpy(), g["ir"])
    with torch.no_grad():
        logits, taps, none = O.pspnet_forward(sd, rgb, ir, late_fusion=True, training=True, dropout=False)
    assert none is None and len(taps) == 6 and taps[0] is logits
    assert _rel(logits.numpy(), g["logits_train"]) < 2e-5
    for i in range(1, 6):
        assert _rel(taps[i][:, ::8].numpy(), g[f"tap{i}_train_sub"]) < 2e-5
    for k in g.files:
        if k.startswith("bn_after_train/"):
            np.testing.assert_allclose(sd[k.split("/", 1)[1]].numpy(), g[k], rtol=1e-4, atol=1e-5)
    with torch.no_grad():
        logits_e, taps_e, _ = O.pspnet_forward(sd, rgb[:1], ir[:1], late_fusion=True, training=False)
    assert _rel(logits_e.numpy(), g["logits_eval"]) < 2e-5
    for i in range(1, 6):
        assert _rel(taps_e[i][:, ::8].numpy(), g[f"tap{i}_eval_sub"]) < 2e-5
    assert (logits_e.argmax(1).numpy() == g["logits_eval"].argmax(1)).mean() > 0.9999


def test_pspnet_early_oracle_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "pspnet_early_golden.npz")
    sd = O.recipe_fill(O.pspnet_state_dict(False, 4), seed=1)
    assert len(sd) == 350
    assert list(sd.keys()) == [str(k) for k in g["state_dict_keys"]]
    rgb, ir = O.synthetic_inputs(2, 64, 96)
    with torch.no_grad():
        logits, _, _ = O.pspnet_forward(sd, rgb[:1], ir[:1], late_fusion=False, training=False)
    assert _rel(logits.numpy(), g["logits_eval"]) < 2e-5


def test_critic_oracle_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "critic_golden.npz")
    sd = O.recipe_fill(O.critic_state_dict(13), seed=2)
    with torch.no_grad():
        y = O.fc_discriminator(torch.from_numpy(g["x"]), sd, "")
    assert y.shape == (2, 1, 64, 96)
    assert _rel(y.numpy(), g["y"]) < 2e-5


@pytest.mark.timeout(600)
def test_conf_segnet_training_step_matches_reference_golden(golden_dir):
    """Both phases of the adversarial step (cm/train_trgb_segnet_conf.py:437-568): losses and the
    gradient of every live parameter tensor vs the reference's autograd."""
    g = _load(golden_dir, "conf_segnet_golden.npz")
    sd = O.recipe_fill(O.conf_segnet_state_dict(True, 6), seed=3)
    assert len(sd) == 548
    rgb_d, ir_d = O.synthetic_inputs(1, 256, 256, seed=11)
    rgb_n, ir_n = O.synthetic_inputs(1, 256, 256, seed=12)
    label = torch.from_numpy(g["label"])
    float_keys = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    for phase in ("train_critic", "train_seg"):
        live = [k for k in float_keys if k.startswith("critics.") == (phase == "train_critic")]
        for k in float_keys:
            sd[k].requires_grad_(k in live)
            sd[k].grad = None
        # train_critic: seg tensors carry requires_grad=False -> no graph through the seg net
        out = O.conf_segnet_forward(sd, [rgb_d, ir_d], [rgb_n, ir_n], training=True, dropout=False)
        total_critics = O.total_critics_loss(out)
        assert abs(total_critics.item() - g[phase + "/total_critics"]) < 2e-4 * abs(g[phase + "/total_critics"])
        assert [str(tuple(c.shape)) for c in out["critics_a"]] == [str(s) for s in g[phase + "/critic_out_shapes"]]
        if phase == "train_seg":
            total, seg_loss, conf = O.train_seg_loss(out, label)
            assert abs(seg_loss.item() - g["seg_loss"]) < 1e-4 * abs(g["seg_loss"])
            assert abs(conf.item() - g["conf_loss"]) < 2e-4 * abs(g["conf_loss"])
        else:
            total = total_critics
        total.backward()
        names = [str(n) for n in g[phase + "/grad_names"]]
        assert names == live
        gn = np.array([sd[k].grad.double().norm().item() for k in live])
        np.testing.assert_allclose(gn, g[phase + "/grad_norm"], rtol=2e-3, atol=1e-7)


def test_inputs_oracle_matches_torchvision():
    """oracle/inputs_oracle.py restates F.to_tensor / F.normalize; check it against torchvision itself (cm/thermal_loader.py:715-728)."""
    import numpy as np
    import torch
    tvf = pytest.importorskip("torchvision.transforms.functional")
    from oracle import inputs_oracle as IO
    rng = np.random.RandomState(0)
    rgb = rng.randint(0, 256, (37, 53, 3)).astype(np.uint8)
    want = tvf.normalize(tvf.to_tensor(rgb), mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5))
    assert torch.equal(IO.load_rgb(rgb), want)
    ir = rng.randint(20000, 27000, (37, 53)).astype(np.uint16)
    x = ir.astype(np.int64)
    x[x < 21800] = 21800
    x[x > 25000] = 25000
    x = (x - 21800) / (25000 - 21800)
    want_ir = tvf.normalize(tvf.to_tensor(x), mean=[0.5], std=[0.5]).float()
    assert torch.equal(IO.load_ir(ir), want_ir)
    assert want_ir.min() == -1.0 and want_ir.max() == 1.0


def test_philox_known_answers():
    """oracle/random_oracle.py against the known-answer vectors of the Random123 distribution (philox4x32, 10 rounds)."""
    from oracle import random_oracle as R
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(v) for v in R.philox4x32_10(np.array(ctr), key)) == want
    s = R.dropout2d_scale(99, 3, 20000, 0.3)
    assert set(np.unique(s).tolist()) == {0.0, float(np.float32(1.0) / (np.float32(1.0) - np.float32(0.3)))}
    assert abs((s > 0).mean() - 0.7) < 0.01
    assert not np.array_equal(s, R.dropout2d_scale(99, 4, 20000, 0.3))
    assert np.count_nonzero(R.dropout2d_scale(99, 0, 100, 1.0)) == 0
