"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the Python host
mirrors the reference's module API / state_dict layout, and the product path refuses to run without a GPU
(no CPU fallback).  No kernel is launched here."""
import os
import re

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import heatnet_oracle as O
from oracle.iou_oracle import IoUOracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()


def test_library_exports_every_declared_symbol():
    from heatnet_pub_b200 import _lib
    header = open(os.path.join(ROOT, "include", "heatnet_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(hn_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/heatnet_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes signatures out of sync with the header"
    assert lib.hn_version() >= 100
    assert lib.hn_conv_kpad(3, 7, 7) == 192 and lib.hn_conv_kpad(2048, 1, 1) == 2048
    assert lib.hn_conv_cout_pad(13, _lib.HN_BF16) == 16 and lib.hn_conv_cout_pad(13, _lib.HN_F32) == 64
    assert lib.hn_conv_cout_pad(2048, _lib.HN_BF16) == 2048


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirrors of the C structs: sizes and field offsets compared with what gcc computes from the header."""
    import ctypes as C
    import subprocess
    from heatnet_pub_b200 import _lib
    structs = {"hn_tensor": _lib.HnTensor, "hn_epilogue": _lib.HnEpilogue, "hn_conv": _lib.HnConv, "hn_pack_job": _lib.HnPackJob}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "heatnet_b200.h"', 'int main(void){']
    for cname, st in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    want = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, st in structs.items():
        assert C.sizeof(st) == int(want[cname]), cname
        for fname, _ in st._fields_:
            assert getattr(st, fname).offset == int(want[f"{cname}.{fname}"]), f"{cname}.{fname}"


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_no_cpu_fallback():
    from heatnet_pub_b200 import pspnet
    net = pspnet.PSPNet(backend='resnet50', pretrained=False, late_fusion=True, in_channels=4).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            net(torch.zeros(1, 3, 32, 32), torch.zeros(1, 1, 32, 32))
    from heatnet_pub_b200 import iou_eval
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        iou_eval.IoU(14).add(torch.zeros(1, 4, 4, dtype=torch.int64), torch.zeros(1, 4, 4, dtype=torch.int64))


def test_product_path_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "heatnet_pub_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no CPU fallback", ""), f"{fn} mentions the oracle"


def test_state_dict_layout_matches_reference(golden_dir):
    from heatnet_pub_b200 import pspnet
    g = np.load(os.path.join(golden_dir, "pspnet_late_golden.npz"))
    net = pspnet.PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50', in_channels=4,
                        pretrained=False, late_fusion=True)
    sd = net.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["state_dict_keys"]]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g["state_dict_shapes"]]
    assert len(sd) == 488
    ge = np.load(os.path.join(golden_dir, "pspnet_early_golden.npz"))
    net_e = pspnet.PSPNet(backend='resnet50', in_channels=4, pretrained=False, late_fusion=False)
    assert list(net_e.state_dict().keys()) == [str(k) for k in ge["state_dict_keys"]]
    # top-level models/pspnet.py signature: n_classes=18, identical keys to the early-fusion 3-channel net
    net_rgb = pspnet.PSPNetRGB(backend='resnet50', pretrained=False)
    assert net_rgb.final[0].out_channels == 18 and len(net_rgb.state_dict()) == 350
    # attributes the reference exposes
    for attr in ("feats", "psp", "drop_1", "up_1", "up_2", "up_3", "drop_2", "final"):
        assert hasattr(net, attr)
    assert isinstance(net.psp.stages[0][0], nn.AdaptiveAvgPool2d) and isinstance(net.up_1.conv[2], nn.PReLU)
    assert net.drop_1.p == 0.3 and net.drop_2.p == 0.15
    # load_state_dict of a reference-shaped state dict works both ways
    net.load_state_dict(O.recipe_fill(O.pspnet_state_dict(True, 4), seed=0))


def test_build_network_both_signatures(monkeypatch):
    from heatnet_pub_b200 import build_net, pspnet
    monkeypatch.setattr(nn.Module, "cuda", lambda self, *a, **k: self)     # the reference calls .cuda() unconditionally
    net = build_net.build_network(None, 'resnet50', in_channels=4, late_fusion=True)
    assert isinstance(net, pspnet.PSPNet) and net.feats.late_fusion and len(net.state_dict()) == 488
    assert set(build_net.models) == {'squeezenet', 'densenet', 'resnet18', 'resnet34', 'resnet50', 'resnet101', 'resnet152'}
    with pytest.raises(KeyError):
        build_net.build_network(None, 'vgg')


def test_conf_segnet_mirror(monkeypatch, capsys):
    from heatnet_pub_b200 import conf_segnet, discriminator_model
    monkeypatch.setattr(nn.Module, "cuda", lambda self, *a, **k: self)
    m = conf_segnet.conv_segnet(pretrained=False, disc_arch='cyclegan', num_critics=6, no_conf=False, modalities='ir_rgb',
                                arch='pspnet', late_fusion=True)
    sd = m.state_dict()
    assert list(sd.keys()) == list(O.conf_segnet_state_dict(True, 6).keys()) and len(sd) == 548
    assert [c.conv1.in_channels for c in m.critics] == [13, 2048, 1024, 1024, 512, 128]
    assert all(isinstance(c, discriminator_model.FCDiscriminator) for c in m.critics)
    assert m.phase == "train_seg"
    m.setPhase("train_critic")
    assert not any(p.requires_grad for p in m.trgb_segnet.parameters()) and all(p.requires_grad for p in m.critics.parameters())
    m.setPhase("train_seg")
    assert all(p.requires_grad for p in m.trgb_segnet.parameters()) and not any(p.requires_grad for p in m.critics.parameters())
    # weights_init_normal reached the BN layers: gamma ~ N(1, 0.02), beta = 0
    bn = m.trgb_segnet.feats.bn1
    assert abs(bn.weight.mean().item() - 1) < 0.02 and bn.bias.abs().max().item() == 0
    with pytest.raises(NotImplementedError):
        conf_segnet.conv_segnet(pretrained=False, disc_arch='cyclegan', arch='custom')


def test_iou_value_host_logic_matches_oracle():
    """IoU.value() (ignore-index zeroing in place, NaN classes, nanmean) is host arithmetic on the K x K matrix."""
    from heatnet_pub_b200 import iou_eval
    rng = np.random.RandomState(0)
    conf = rng.randint(0, 1000, (14, 14)).astype(np.int32)
    conf[5, :] = 0
    conf[:, 5] = 0
    m, mo = iou_eval.IoU(14, False, [12, 13]), IoUOracle(14, False, [12, 13])
    m.conf_metric.conf[:] = conf
    mo.conf_metric.conf[:] = conf
    (iou, miou), (iou_o, miou_o) = m.value(), mo.value()
    assert np.array_equal(iou, iou_o, equal_nan=True) and miou == miou_o and np.isnan(iou[5])
    assert np.array_equal(m.conf_metric.conf, mo.conf_metric.conf) and m.conf_metric.conf[12].sum() == 0
    with pytest.raises(ValueError):
        iou_eval.IoU(14, False, 1.5)


def test_input_pipeline_lookup_tables_are_the_loader_arithmetic():
    """heatnet_pub_b200.inputs evaluates the loaders' normalisation once per possible input value (CPU, no GPU needed): the tables
    must equal oracle/inputs_oracle.py (= torchvision's to_tensor / normalize, cm/thermal_loader.py:649-659,715-728) bit for bit."""
    import numpy as np
    import torch
    from heatnet_pub_b200 import inputs as I
    from oracle import inputs_oracle as IO
    for mean, std in (((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)), ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))):
        lut = I._rgb_lut(mean, std, "cpu")
        img = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 3, axis=2)          # (256, 1, 3): every byte value in every channel
        want = IO.load_rgb(img, mean, std).reshape(3, 256)
        assert torch.equal(lut, want)
    for lo, hi in ((21800, 25000), (20800, 27000)):
        lut = I._ir_lut(lo, hi, 0.5, 0.5, "cpu")
        counts = np.arange(lo, hi + 1, dtype=np.uint16)[None, :]
        assert torch.equal(lut, IO.load_ir(counts, lo, hi).reshape(-1))
        assert lut[0] == -1.0 and lut[-1] == 1.0


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) needs no GPU: exactly one JSON line on stdout with the
    contract's keys, whatever the libraries print elsewhere."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--height", "64",
                          "--width", "96"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "rgb_thermal_seg_images_per_sec" and d["value"] > 0
    from oracle import reference_loader as RL
    # the unmodified reference modules when baseline/_ref travelled with the snapshot, the oracle port otherwise
    assert d["cpu_baseline"]["kind"] == ("reference" if RL.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_install_matches_the_oracle_bit_for_bit():
    """baseline/_ref (file-level install of the reference's hot-path modules) imports cleanly next to this package, and the
    reference's own PSPNet.eval() forward equals the oracle restatement bit for bit on the same weights and frame."""
    from oracle import reference_loader as RL
    if not RL.available():
        pytest.skip("baseline/_ref not installed")
    PSPNet = RL.load_cm_pspnet()
    net = PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50', in_channels=4, pretrained=False, late_fusion=True)
    sd = O.recipe_fill(O.pspnet_state_dict(True, 4), seed=0)
    net.load_state_dict(sd)
    net.eval()
    rgb, ir = O.synthetic_inputs(1, 32, 64)
    import warnings
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = net(rgb, ir)[0]
        b = O.pspnet_forward(sd, rgb, ir, late_fusion=True, training=False)[0]
    assert torch.equal(a, b)
    assert "models" not in __import__("sys").modules or not getattr(__import__("sys").modules["models"], "__file__", "").startswith(RL.REF_DIR)


# ---------------------------------------------------------------------------------------------------- checkpoint I/O
def test_checkpoint_helpers_roundtrip(tmp_path, monkeypatch, capsys):
    """utils.initModelRenamed / Partial / Full (cm/utils.py:59-90) on the on-disk formats the trainers write: a DataParallel
    `'module.trgb_segnet.'`-prefixed {'state_dict': ...} checkpoint (cm/train_trgb_segnet_conf.py:643-655), the
    `<name>_<epoch>` snapshot convention of build_network (models/build_net.py:23-26, both signatures), and an
    optimizer + StepLR resume (cm/train_trgb_segnet_conf.py:276-283)."""
    from heatnet_pub_b200 import build_net, pspnet, utils
    monkeypatch.setattr(nn.Module, "cuda", lambda self, *a, **k: self)
    sd = O.recipe_fill(O.pspnet_state_dict(True, 4), seed=4)

    def fresh():
        return pspnet.PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50', in_channels=4,
                             pretrained=False, late_fusion=True)

    def same(net):
        got = net.state_dict()
        return all(torch.equal(got[k], v) for k, v in sd.items())

    # (1) DataParallel(conv_segnet) checkpoint -> seg net only (train_trgb_segnet_conf.py:231 night-supervision load)
    ck = {"state_dict": {"module.trgb_segnet." + k: v for k, v in sd.items()}, "epoch": 7, "best_iou": 0.5}
    ck["state_dict"]["module.critics.0.conv1.weight"] = torch.zeros(64, 13, 4, 4)          # foreign keys are ignored
    path = str(tmp_path / "checkpoint.pth.tar")
    torch.save(ck, path)
    net = fresh()
    utils.initModelRenamed(net, path, "module.trgb_segnet.", "")
    assert same(net) and "Loaded dict with 488 entries..." in capsys.readouterr().out
    # bare state_dict with the plain DataParallel prefix (conf_segnet.py:80)
    torch.save({"module." + k: v for k, v in sd.items()}, path)
    net = fresh()
    utils.initModelRenamed(net, path, "module.", "")
    assert same(net)
    with pytest.raises(AssertionError):                       # nothing matches: the reference asserts len > 0
        utils.initModelRenamed(fresh(), path, "module.", "other.")
    # (2) partial: only matching keys are taken, the rest keeps its initialisation
    part = {"state_dict": {k: v for k, v in sd.items() if k.startswith("feats.")} | {"not.a.key": torch.zeros(1)}}
    torch.save(part, path)
    net = fresh()
    before = net.final[0].weight.clone()
    utils.initModelPartial(net, path)
    got = net.state_dict()
    assert all(torch.equal(got[k], v) for k, v in part["state_dict"].items() if k in got) and torch.equal(net.final[0].weight, before)
    assert "Updated : %d entries" % len([k for k in part["state_dict"] if k in got]) in capsys.readouterr().out
    # (3) full
    torch.save(sd, path)
    net = fresh()
    utils.initModelFull(net, path)
    assert same(net)
    # (4) build_network snapshot convention, HeatNet signature (returns net) and top-level signature (returns (net, epoch))
    snap = str(tmp_path / "PSPNet_7")
    torch.save(sd, snap)
    net = build_net.build_network(snap, 'resnet50', in_channels=4, late_fusion=True)
    assert isinstance(net, pspnet.PSPNet) and same(net)
    sd_rgb = O.recipe_fill(O.pspnet_state_dict(False, 3, n_classes=18), seed=6)
    snap2 = str(tmp_path / "x_12")
    torch.save(sd_rgb, snap2)
    monkeypatch.setitem(build_net.models, 'resnet50',
                        lambda: pspnet.PSPNetRGB(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50', pretrained=False))
    net2, epoch = build_net.build_network(snap2, 'ResNet50')          # backend is lower-cased like the reference does
    assert epoch == 12 and isinstance(net2, pspnet.PSPNetRGB)
    assert all(torch.equal(net2.state_dict()[k], v) for k, v in sd_rgb.items())
    with pytest.raises(ValueError):                            # 'a_b_3' does not split into exactly (name, epoch): same failure as the reference
        build_net.build_network(str(tmp_path / "a_b_3"), 'resnet50')


def test_optimizer_and_scheduler_resume_roundtrip(tmp_path):
    """cm/train_trgb_segnet_conf.py:276-283,643-655: checkpoint = {'state_dict', 'optimizer', 'lr_scheduler', 'epoch',
    'best_iou'}; the fused RMSprop's state_dict loads into torch.optim.RMSprop and back (same keys), StepLR halves its lr."""
    from heatnet_pub_b200 import optim
    ps = [nn.Parameter(torch.randn(4, 3)), nn.Parameter(torch.randn(5))]
    opt = optim.RMSprop(ps, lr=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
    for p in ps:                                              # fabricate the state a few steps would leave (no GPU in this test)
        opt.state[p] = {"step": torch.tensor(3.0), "square_avg": torch.rand_like(p)}
    for _ in range(2):
        sched.step()
    assert opt.param_groups[0]["lr"] == pytest.approx(5e-4)
    path = str(tmp_path / "ck.pth.tar")
    torch.save({"optimizer": opt.state_dict(), "lr_scheduler": sched.state_dict(), "epoch": 2, "best_iou": 0.1}, path)
    ck = torch.load(path)
    ps2 = [nn.Parameter(p.detach().clone()) for p in ps]
    opt2 = optim.RMSprop(ps2, lr=1e-3)
    sched2 = torch.optim.lr_scheduler.StepLR(opt2, step_size=2, gamma=0.5)
    opt2.load_state_dict(ck["optimizer"])
    sched2.load_state_dict(ck["lr_scheduler"])
    assert opt2.param_groups[0]["lr"] == pytest.approx(5e-4) and sched2.last_epoch == 2
    for a, b in zip(ps, ps2):
        assert torch.equal(opt.state[a]["square_avg"], opt2.state[b]["square_avg"]) and float(opt2.state[b]["step"]) == 3.0
    ref = torch.optim.RMSprop([nn.Parameter(p.detach().clone()) for p in ps], lr=1e-3)       # and into stock torch
    ref.load_state_dict(ck["optimizer"])
    assert ref.param_groups[0]["lr"] == pytest.approx(5e-4)
