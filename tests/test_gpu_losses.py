"""Fused loss kernels (hn_ce_loss_fwd_bwd / hn_critic_loss_fwd_bwd) against the arithmetic the reference's criteria run:
torch.nn.CrossEntropyLoss / MSELoss / BCEWithLogitsLoss in FP64 on the CPU, same inputs.  Tolerances: loss 1e-6 relative,
gradients 1e-6 of the largest gradient entry (FP32 kernels; the reference's own FP32 CUDA kernels sit at the same distance
from FP64)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _close(got, ref, tol=1e-6):
    got, ref = got.detach().cpu().double(), ref.detach().double()
    scale = ref.abs().max().clamp_min(1e-30)
    return ((got - ref).abs().max() / scale).item() < tol


@pytest.mark.parametrize("shape", [(2, 13, 64, 96), (1, 13, 37, 53), (3, 14, 8, 8), (1, 19, 5, 7), (2, 1, 4, 4)])
@pytest.mark.parametrize("ignore", [-100, 13])
def test_cross_entropy_matches_torch(shape, ignore):
    from heatnet_pub_b200 import losses
    g = torch.Generator().manual_seed(11)
    n, k, h, w = shape
    logits = torch.randn(shape, generator=g) * 3
    labels = torch.randint(0, k, (n, h, w), generator=g)
    labels[0, 0, :3] = ignore                              # ignored pixels; legal for torch whether or not ignore is in [0,K)
    labels[-1, -1, -1] = ignore
    upstream = 0.37
    x64 = logits.double().requires_grad_(True)
    ref = nn.CrossEntropyLoss(ignore_index=ignore)(x64, labels)
    (ref * upstream).backward()

    x = logits.cuda().requires_grad_(True)
    crit = losses.CrossEntropyLoss(ignore_index=ignore)
    loss = crit(x, labels.cuda())
    (loss * upstream).backward()
    assert loss.shape == () and loss.dtype == torch.float32
    assert abs(loss.item() - ref.item()) <= 1e-6 * abs(ref.item()) + 1e-7
    assert _close(x.grad, x64.grad)
    crit.check_labels()


def test_cross_entropy_all_ignored_is_nan_and_zero_grad():
    from heatnet_pub_b200 import losses
    logits = torch.randn(1, 13, 4, 4).cuda().requires_grad_(True)
    labels = torch.full((1, 4, 4), 13, dtype=torch.int64).cuda()
    loss = losses.CrossEntropyLoss(ignore_index=13)(logits, labels)
    assert torch.isnan(loss).item()                      # torch: 0/0
    ref = nn.CrossEntropyLoss(ignore_index=13)(logits.detach().cpu(), labels.cpu())
    assert torch.isnan(ref).item()


def test_cross_entropy_out_of_range_label_is_flagged():
    from heatnet_pub_b200 import losses
    logits = torch.randn(1, 13, 4, 4).cuda()
    labels = torch.zeros((1, 4, 4), dtype=torch.int64)
    labels[0, 1, 1] = 40
    crit = losses.CrossEntropyLoss()
    crit(logits, labels.cuda())
    with pytest.raises(IndexError):
        crit.check_labels()


def test_cross_entropy_argument_errors():
    from heatnet_pub_b200 import losses
    with pytest.raises(NotImplementedError):
        losses.CrossEntropyLoss(weight=torch.ones(13))
    with pytest.raises(NotImplementedError):
        losses.CrossEntropyLoss(reduction='none')
    crit = losses.CrossEntropyLoss()
    with pytest.raises(ValueError):
        crit(torch.randn(1, 13, 4, 4).cuda(), torch.zeros((1, 5, 4), dtype=torch.int64).cuda())
    with pytest.raises(RuntimeError):
        crit(torch.randn(1, 13, 4, 4).cuda(), torch.zeros((1, 4, 4), dtype=torch.int32).cuda())


@pytest.mark.parametrize("n", [1, 3, 1000, 2 * 320 * 640 + 1])
@pytest.mark.parametrize("kind", ["mse", "bce"])
@pytest.mark.parametrize("target", [0.0, 1.0])
def test_critic_losses_match_torch(n, kind, target):
    from heatnet_pub_b200 import losses
    g = torch.Generator().manual_seed(n)
    c = torch.randn(n, generator=g) * 4
    c = c.view(1, 1, 1, n)
    ref_crit = nn.MSELoss() if kind == "mse" else nn.BCEWithLogitsLoss()
    crit = losses.MSELoss() if kind == "mse" else losses.BCEWithLogitsLoss()
    c64 = c.double().requires_grad_(True)
    ref = ref_crit(c64, torch.full_like(c64, target))
    (ref * 0.1).backward()
    for as_tensor in (False, True):                      # constant fast path and the reference's torch.full_like target tensor
        x = c.cuda().requires_grad_(True)
        loss = crit(x, torch.full_like(x, target) if as_tensor else target)
        (loss * 0.1).backward()
        assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item()) + 1e-7
        assert _close(x.grad, c64.grad, 2e-6)


def test_critic_loss_on_offset_views():
    """Unaligned storage offsets take the scalar tail path."""
    from heatnet_pub_b200 import losses
    base = torch.randn(1031).cuda()
    x = base[3:].clone().requires_grad_(True)
    loss = losses.MSELoss()(x.view(1, 1, 4, 257), 1.0)
    loss.backward()
    ref = ((base[3:].double() - 1) ** 2).mean()
    assert abs(loss.item() - ref.item()) < 1e-6 * ref.item()


def test_training_step_losses_match_reference_formulation():
    """The train_seg loss of cm/train_trgb_segnet_conf.py:452,529-546 assembled from the fused criteria equals the torch
    formulation on the same (random) logits and critic maps, including the gradients that flow back."""
    from heatnet_pub_b200 import losses
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(2, 13, 32, 64, generator=g)
    label = torch.randint(0, 13, (2, 32, 64), generator=g)
    critics = [torch.randn(2, 1, s, 2 * s, generator=g) for s in (32, 32, 64, 64, 32, 320)]

    def total(ce, mse, lg, cs, lab):
        conf = sum(mse(c, torch.full_like(c, 1)) for c in cs)
        return ce(lg, lab) + 0.1 * conf

    lg64 = logits.double().requires_grad_(True)
    cs64 = [c.double().requires_grad_(True) for c in critics]
    ref = total(nn.CrossEntropyLoss(), nn.MSELoss(), lg64, cs64, label)
    ref.backward()
    lg = logits.cuda().requires_grad_(True)
    cs = [c.cuda().requires_grad_(True) for c in critics]
    out = total(losses.CrossEntropyLoss(), losses.MSELoss(), lg, cs, label.cuda())
    out.backward()
    assert abs(out.item() - ref.item()) < 2e-6 * abs(ref.item())
    assert _close(lg.grad, lg64.grad)
    for a, b in zip(cs, cs64):
        assert _close(a.grad, b.grad, 2e-6)
