"""Fused multi-tensor optimiser kernels against torch.optim (the reference's optimisers: RMSprop at
cm/train_trgb_segnet_conf.py:270, Adam at scripts/main.py:159, clip_grad_norm at scripts/main.py:256-257) run in FP64 on the
CPU with the same parameters and gradient sequence.  Tolerance 2e-6 of the largest parameter entry after 5 steps."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(64, 3, 7, 7), (64,), (1,), (256, 64, 1, 1), (13, 64, 1, 1), (5,), (70001,), (16384,), (16385,)]


def _params(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) for s in SHAPES]


def _grads(step, seed=0):
    g = torch.Generator().manual_seed(1000 + step + seed)
    return [torch.randn(s, generator=g) * (10.0 if step == 2 else 0.1) for s in SHAPES]


def _run(opt_ref_cls, opt_cls, kwargs, steps=5, max_norm=None, grad_scale=1.0, skip=()):
    p_ref = [torch.nn.Parameter(p.double()) for p in _params()]
    p_gpu = [torch.nn.Parameter(p.cuda()) for p in _params()]
    ref = opt_ref_cls(p_ref, **kwargs)
    extra = {}
    if max_norm is not None:
        extra["max_norm"] = max_norm
    if grad_scale != 1.0:
        extra["grad_scale"] = grad_scale
    opt = opt_cls(p_gpu, **kwargs, **extra)
    for s in range(steps):
        for i, (a, b, g) in enumerate(zip(p_ref, p_gpu, _grads(s))):
            if i in skip and s < 2:                      # parameters without a gradient are skipped, state starts later
                a.grad, b.grad = None, None
                continue
            a.grad = g.double() * grad_scale
            b.grad = g.cuda()
        if max_norm is not None:
            torch.nn.utils.clip_grad_norm_([p for p in p_ref if p.grad is not None], max_norm)
        v0 = [p._version for p in p_gpu]
        ref.step()
        opt.step()
        assert all(p._version > v for p, v in zip(p_gpu, v0) if p.grad is not None)
    for a, b in zip(p_ref, p_gpu):
        err = (b.detach().cpu().double() - a.detach()).abs().max() / a.detach().abs().max()
        assert err.item() < 2e-6, err.item()
    return ref, opt


@pytest.mark.parametrize("kwargs", [dict(lr=1e-3), dict(lr=1e-2, alpha=0.9, eps=1e-6, weight_decay=1e-3), dict(lr=1e-3, momentum=0.9)])
def test_rmsprop_matches_torch(kwargs):
    from heatnet_pub_b200 import optim
    ref, opt = _run(torch.optim.RMSprop, optim.RMSprop, kwargs)
    sr, so = ref.state_dict(), opt.state_dict()
    assert set(sr["state"][0]) == set(so["state"][0])               # same per-parameter state keys as torch (checkpoint compat)
    assert set(sr["param_groups"][0]) >= {"lr", "alpha", "eps", "weight_decay", "momentum", "centered"} <= set(so["param_groups"][0])


@pytest.mark.parametrize("kwargs", [dict(lr=1e-3), dict(lr=1e-4, betas=(0.8, 0.99), eps=1e-6, weight_decay=1e-2)])
def test_adam_matches_torch(kwargs):
    from heatnet_pub_b200 import optim
    ref, opt = _run(torch.optim.Adam, optim.Adam, kwargs)
    assert {"step", "exp_avg", "exp_avg_sq"} == set(opt.state_dict()["state"][0])


def test_clip_and_grad_scale_fused():
    """clip_grad_norm_(params, 0.5) + the 1/world gradient average, applied inside the update kernel; step 2 has large
    gradients (clip active), the others small ones (clip inactive)."""
    from heatnet_pub_b200 import optim
    _run(torch.optim.Adam, optim.Adam, dict(lr=1e-3), max_norm=0.5, grad_scale=0.125)
    ref, opt = _run(torch.optim.RMSprop, optim.RMSprop, dict(lr=1e-3), max_norm=0.5)
    total = np.sqrt(sum(float((g.double() ** 2).sum()) for g in _grads(4)))
    assert abs(opt.total_grad_norm() - total) < 1e-9 * total


def test_parameters_without_grad_are_skipped():
    from heatnet_pub_b200 import optim
    _run(torch.optim.RMSprop, optim.RMSprop, dict(lr=1e-3), skip=(1, 3))
    _run(torch.optim.Adam, optim.Adam, dict(lr=1e-3), skip=(0, 6))        # distinct step counts -> separate bias corrections


def test_state_dict_roundtrip_and_lr_schedulers():
    from heatnet_pub_b200 import optim
    p = [torch.nn.Parameter(torch.randn(100).cuda())]
    opt = optim.RMSprop(p, lr=0.1)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)   # cm/train_trgb_segnet_conf.py:274
    p[0].grad = torch.ones(100).cuda()
    opt.step()
    sched.step()
    assert opt.param_groups[0]["lr"] == pytest.approx(0.05)
    assert optim.poly_lr_scheduler(opt, 0.1, 50, max_iter=100, power=0.9) == pytest.approx(0.1 * 0.5 ** 0.9)
    sd = opt.state_dict()
    opt2 = optim.RMSprop([torch.nn.Parameter(p[0].detach().clone())], lr=0.1)
    opt2.load_state_dict(sd)
    t = torch.optim.RMSprop([torch.nn.Parameter(p[0].detach().clone())], lr=0.1)
    t.load_state_dict(sd)                                                   # a torch optimizer accepts our checkpoint
    assert torch.equal(t.state_dict()["state"][0]["square_avg"], sd["state"][0]["square_avg"])


def test_argument_errors():
    from heatnet_pub_b200 import optim
    p = [torch.nn.Parameter(torch.randn(4).cuda())]
    with pytest.raises(NotImplementedError):
        optim.RMSprop(p, centered=True)
    with pytest.raises(NotImplementedError):
        optim.Adam(p, amsgrad=True)
    with pytest.raises(ValueError):
        optim.RMSprop(p, lr=-1.0)
    with pytest.raises(ValueError):
        optim.Adam(p, betas=(1.0, 0.9))


def test_captured_step_survives_interleaved_eager_steps():
    """A captured optimizer step keeps its own (immutable) slot tables: graph replay -> eager step on NEW gradient tensors ->
    graph replay again must equal the all-eager sequence.  (Round-1 shared one pinned staging buffer between eager steps and
    captures: the eager step overwrote the addresses the graph's H2D node re-reads.)"""
    from heatnet_pub_b200 import optim
    steps = 4

    def make():
        ps = [torch.nn.Parameter(p.cuda()) for p in _params(3)]
        return ps, optim.RMSprop(ps, lr=1e-2)

    # all-eager reference: gradients g0, g1, g2, g3
    ps_ref, opt_ref = make()
    for s in range(steps):
        for p, g in zip(ps_ref, _grads(s, 5)):
            p.grad = g.cuda()
        opt_ref.step()
    # graph for steps 0 and 2, eager steps 1 and 3 on fresh gradient tensors.  Two captured variants: gradients at static
    # addresses (table set built by the warm-up, reused by the capture) and gradients allocated inside the capture (graph pool:
    # the table set is built DURING the capture from a spare pinned pair).
    from heatnet_pub_b200 import graphs
    for in_pool in (False, True):
        ps, opt = make()

        def fn(*gs):
            for p, g in zip(ps, gs):
                p.grad = g * 1.0 if in_pool else g
            opt.step()
            return ps[0]

        zeros = [torch.zeros_like(p) for p in ps]       # warm-up with all-zero gradients: RMSprop leaves parameters and square_avg
        gstep = graphs.GraphedStep(fn, zeros, warmup=2)  # untouched (0 / (0 + eps)), so the captured sequence starts from scratch
        for s in range(steps):
            gs = [g.cuda() for g in _grads(s, 5)]
            if s % 2 == 0:
                gstep(*gs)
            else:
                for p, g in zip(ps, gs):
                    p.grad = g.clone()                     # new addresses: the eager path builds (and may evict) its own tables
                opt.step()
        torch.cuda.synchronize()
        for a, b in zip(ps_ref, ps):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), in_pool
