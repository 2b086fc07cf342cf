"""Drop-in proof: the REFERENCE's own caller code, unmodified, executed over this package.

`models/confusion_maximization/models/conf_segnet.py` (the adversarial wrapper the trainer instantiates,
cm/train_trgb_segnet_conf.py:208) is loaded from baseline/_ref (baseline/install_reference.py: a file-level copy of the
reference's hot-path modules, git-ignored, shipped with the snapshot) with its imports -- `discriminator_model`, `utils`,
`models.build_net` -- resolved to heatnet_pub_b200 through sys.modules.  Its `__init__`, `setPhase` and `forward` then drive
the B200 kernels, and the results are held against the golden vectors the reference produced with its own modules."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import heatnet_oracle as O
from oracle import reference_loader as RL

pytestmark = pytest.mark.gpu


@pytest.fixture()
def reference_conf_segnet():
    path = os.path.join(RL.CM_DIR, "models", "conf_segnet.py")
    if not os.path.exists(path):
        pytest.skip("baseline/_ref is not installed (python baseline/install_reference.py in the build container)")
    import heatnet_pub_b200 as pkg
    from heatnet_pub_b200 import build_net, discriminator_model, utils
    shim = types.ModuleType("models")
    shim.__path__ = []
    shim.build_net = build_net
    mods = {"models": shim, "models.build_net": build_net, "discriminator_model": discriminator_model, "utils": utils}
    for stub in ("trgb_segnet", "critic_resnet", "downscale_network", "input_adapter"):     # out-of-scope siblings: import targets only
        m = types.ModuleType("models." + stub)
        setattr(shim, stub, m)
        mods["models." + stub] = m
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        spec = importlib.util.spec_from_file_location("_ref_conf_segnet", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        yield mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.timeout(900)
def test_reference_conf_segnet_wrapper_runs_over_the_package(reference_conf_segnet, golden_dir, capsys):
    ref = reference_conf_segnet
    from heatnet_pub_b200 import discriminator_model, pspnet
    m = ref.conv_segnet(pretrained=False, disc_arch='cyclegan', num_critics=6, no_conf=False, modalities='ir_rgb', arch='pspnet',
                        late_fusion=True)
    assert type(m).__module__ == "_ref_conf_segnet"                                  # the reference's class ...
    assert isinstance(m.trgb_segnet, pspnet.PSPNet)                                   # ... built on this package's modules
    assert all(isinstance(c, discriminator_model.FCDiscriminator) for c in m.critics) and len(m.critics) == 6
    assert next(m.trgb_segnet.parameters()).is_cuda                                   # build_network(...).cuda(), build_net.py:27
    sd = O.recipe_fill(O.conf_segnet_state_dict(True, 6), seed=3)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.cuda().train()
    m.trgb_segnet.set_precision("fp32")
    for c in m.critics:
        c.precision = "fp32"
    m.trgb_segnet.drop_1.p = m.trgb_segnet.drop_2.p = 0.0
    g = np.load(os.path.join(golden_dir, "conf_segnet_golden.npz"))
    rgb_d, ir_d = O.synthetic_inputs(1, 256, 256, seed=11)
    rgb_n, ir_n = O.synthetic_inputs(1, 256, 256, seed=12)
    label = torch.from_numpy(g["label"]).cuda()
    mse, ce = nn.MSELoss(), nn.CrossEntropyLoss()
    for phase in ("train_critic", "train_seg"):
        m.setPhase(phase)                                                             # the reference's own phase switch
        assert all(p.requires_grad == (phase == "train_seg") for p in m.trgb_segnet.parameters())
        assert all(p.requires_grad == (phase == "train_critic") for c in m.critics for p in c.parameters())
        for p in m.parameters():
            p.grad = None
        o = m([rgb_d.cuda(), ir_d.cuda()], [rgb_n.cuda(), ir_n.cuda()])               # the reference's own forward
        assert set(o) == {"critics_a", "critics_b", "pred_label_a", "pred_label_b", "cert_a", "cert_b", "inter_f_b"}
        total_critics = sum(mse(c, torch.full_like(c, 1)) for c in o['critics_a']) + sum(mse(c, torch.full_like(c, 0)) for c in o['critics_b'])
        assert abs(total_critics.item() - g[phase + "/total_critics"]) < 5e-4 * abs(g[phase + "/total_critics"])
        if phase == "train_seg":
            conf = sum(mse(c, torch.full_like(c, 1)) for c in o['critics_a']) + sum(mse(c, torch.full_like(c, 1)) for c in o['critics_b'])
            total = ce(o['pred_label_a'], label) + 0.1 * conf
        else:
            total = total_critics
        total.backward()
        names = [str(n) for n in g[phase + "/grad_names"]]
        assert [k for k, p in m.named_parameters() if p.grad is not None] == names
        gn = np.array([dict(m.named_parameters())[k].grad.double().norm().item() for k in names])
        np.testing.assert_allclose(gn, g[phase + "/grad_norm"], rtol=3e-3, atol=1e-5 * g[phase + "/grad_norm"].max())
    out = capsys.readouterr().out
    assert "Using RGB" in out and "Creating 6 critics...." in out and "Switching to phase: train_seg" in out
