"""Fused multi-tensor optimisers for the trainers on this path.

`RMSprop` replaces `torch.optim.RMSprop(conf_segnet_model.parameters(), lr=opt.lr)`
(`models/confusion_maximization/train_trgb_segnet_conf.py:270`) and `Adam` replaces
`Adam(model.parameters(), lr=...)` (`scripts/main.py:159`).  Both subclass `torch.optim.Optimizer` with torch's
hyper-parameter names and per-parameter state keys (`step`, `square_avg`, `momentum_buffer`, `exp_avg`,
`exp_avg_sq`), so `lr_scheduler.StepLR`, `poly_lr_scheduler` (which writes `param_groups[i]['lr']`) and optimizer
checkpoints (`state_dict()` / `load_state_dict()`, `train_trgb_segnet_conf.py:281,652`) interoperate.

One kernel launch updates every parameter of a group (hn_rmsprop_step / hn_adam_step), with two things folded in:
the gradient average of the data-parallel all-reduce (`grad_scale`) and `clip_grad_norm` (`scripts/main.py:256-257`,
`max_norm=`; the clip coefficient comes from one extra read of the gradients, hn_grad_sqnorm, and is applied inside
the update instead of rewriting the gradients).  Parameters whose `.grad` is None are skipped, as torch does
(`conv_segnet.setPhase` leaves the frozen half of the model without gradients).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import engine as E


class _FusedOptimizer(torch.optim.Optimizer):
    _state_names = ()

    def __init__(self, params, defaults, max_norm=None, grad_scale=1.0):
        super().__init__(params, defaults)
        self.max_norm = max_norm
        self.grad_scale = float(grad_scale)
        self._tables = {}
        self._spares = {}
        self.last_sqnorm = None

    # ---- device tables: slots (param, grad, state1, state2, numel) and the block -> (slot, chunk) map
    EAGER_TABLES = 2          # eager-built table sets kept per param group (gradient tensors outside an arena move every step)

    def _table(self, gi, live, states):
        """Table set for this exact set of addresses.  A set is IMMUTABLE once built: its pinned staging buffers are written
        once and never reused for other contents, because a CUDA graph that captured the H2D copy re-reads them on every replay
        (and its kernels read the device tables).  Sets built during a capture are therefore kept for good (the graph needs
        them as long as it lives); sets built eagerly are kept in a small LRU.  With a gradient arena
        (parallel.GradientReducer) the addresses never change and one set serves eager steps and captures alike."""
        key = tuple((p.data_ptr(), p.grad.data_ptr(), *(s.data_ptr() if s is not None else 0 for s in st)) for p, st in zip(live, states))
        cache = self._tables.setdefault(gi, {})
        hit = cache.get(key)
        if hit is not None:
            if torch.cuda.is_current_stream_capturing():
                hit["captured"] = True                 # a graph now points at these tables: never evict
            elif not hit["captured"]:
                cache[key] = cache.pop(key)            # LRU order
            return hit["slots_d"], hit["bmap_d"], hit["partial"], hit["nblocks"]
        chunk = _lib.load().hn_optim_chunk()
        slots = np.zeros((len(live), 5), dtype=np.int64)
        blocks = []
        for i, (p, st) in enumerate(zip(live, states)):
            slots[i] = (p.data_ptr(), p.grad.data_ptr(), st[0].data_ptr(), st[1].data_ptr() if st[1] is not None else 0, p.numel())
            nchunk = (p.numel() + chunk - 1) // chunk
            blocks.append(np.stack([np.full(nchunk, i, dtype=np.int32), np.arange(nchunk, dtype=np.int32)], axis=1))
        bmap = np.concatenate(blocks, axis=0) if blocks else np.zeros((0, 2), dtype=np.int32)
        dev = live[0].device
        capturing = torch.cuda.is_current_stream_capturing()
        # the tables travel through PINNED staging buffers with asynchronous copies: no host sync per rebuild, and a rebuild
        # inside a capture is a memcpy node
        # (pinned memory cannot be allocated while a stream is capturing: every eager build leaves one spare pair of this
        # shape behind for a capture to take over -- graphs.GraphedStep always runs the step eagerly first)
        shape_key = (len(live), max(bmap.shape[0], 1))
        spares = self._spares.setdefault((gi, shape_key), [])
        if capturing and spares:
            slots_h, bmap_h = spares.pop()
        else:
            slots_h = torch.empty((shape_key[0], 5), dtype=torch.int64).pin_memory()
            bmap_h = torch.empty((shape_key[1], 2), dtype=torch.int32).pin_memory()
        if not capturing and not spares:
            spares.append((torch.empty((shape_key[0], 5), dtype=torch.int64).pin_memory(), torch.empty((shape_key[1], 2), dtype=torch.int32).pin_memory()))
        slots_h.numpy()[...] = slots
        if bmap.shape[0]:
            bmap_h.numpy()[...] = bmap
        ent = {"slots_h": slots_h, "bmap_h": bmap_h, "captured": capturing, "nblocks": bmap.shape[0],
               "slots_d": torch.empty(slots_h.shape, dtype=torch.int64, device=dev), "bmap_d": torch.empty(bmap_h.shape, dtype=torch.int32, device=dev),
               "partial": torch.empty(bmap.shape[0] + 1, dtype=torch.float64, device=dev)}
        ent["slots_d"].copy_(slots_h, non_blocking=True)
        ent["bmap_d"].copy_(bmap_h, non_blocking=True)
        cache[key] = ent
        eager = [k for k, v in cache.items() if not v["captured"]]
        for k in eager[:max(0, len(eager) - self.EAGER_TABLES)]:
            del cache[k]
        return ent["slots_d"], ent["bmap_d"], ent["partial"], ent["nblocks"]

    def _live(self, group):
        live = []
        for p in group['params']:
            if p.grad is None:
                continue
            if p.grad.is_sparse:
                raise RuntimeError(f"{type(self).__name__} does not support sparse gradients")
            if p.dtype != torch.float32 or not p.is_contiguous() or not p.is_cuda:
                raise RuntimeError("heatnet_pub_b200.optim: parameters must be dense FP32 CUDA tensors (the FP32 masters)")
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                p.grad = p.grad.float().contiguous()
            live.append(p)
        return live

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        _lib.require_device()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            live = self._live(group)
            if not live:
                continue
            states = [self._init_state(p, group) for p in live]
            slots_d, bmap_d, partial, nblocks = self._table(gi, live, states)
            sq_ptr = None
            max_norm = float(self.max_norm) if self.max_norm else 0.0
            if max_norm > 0.0:
                sq = partial[nblocks:]
                _lib.check(lib.hn_grad_sqnorm(slots_d.data_ptr(), bmap_d.data_ptr(), nblocks, partial.data_ptr(), sq.data_ptr(), E._stream()))
                E._count(2)
                sq_ptr, self.last_sqnorm = sq.data_ptr(), sq
            self._launch(lib, group, live, slots_d, bmap_d, nblocks, max_norm, sq_ptr)
            E._count()
            torch.autograd.graph.increment_version(live)      # the kernel writes behind torch's back: refresh packed-weight caches
        return loss

    def total_grad_norm(self):
        """sqrt of the last clip pass's squared norm, times |grad_scale| (what clip_grad_norm returns); one D2H read."""
        return None if self.last_sqnorm is None else float(self.last_sqnorm.item()) ** 0.5 * abs(self.grad_scale)


class RMSprop(_FusedOptimizer):
    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=0, momentum=0, centered=False, max_norm=None, grad_scale=1.0):
        if centered:
            raise NotImplementedError("centered RMSprop is not on the hot path")
        if lr < 0.0 or eps < 0.0 or momentum < 0.0 or weight_decay < 0.0 or alpha < 0.0:
            raise ValueError("Invalid hyper-parameter (negative lr / eps / momentum / weight_decay / alpha)")
        super().__init__(params, dict(lr=lr, alpha=alpha, eps=eps, weight_decay=weight_decay, momentum=momentum, centered=False),
                         max_norm, grad_scale)

    def _init_state(self, p, group):
        st = self.state[p]
        if len(st) == 0:
            st['step'] = torch.tensor(0.0)
            st['square_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if group['momentum'] > 0:
                st['momentum_buffer'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st['step'] += 1
        return st['square_avg'], st.get('momentum_buffer')

    def _launch(self, lib, group, live, slots_d, bmap_d, nblocks, max_norm, sq_ptr):
        _lib.check(lib.hn_rmsprop_step(slots_d.data_ptr(), bmap_d.data_ptr(), nblocks, group['lr'], group['alpha'], group['eps'],
                                       group['weight_decay'], group['momentum'], self.grad_scale, max_norm, sq_ptr, E._stream()))


class Adam(_FusedOptimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, max_norm=None, grad_scale=1.0):
        if amsgrad:
            raise NotImplementedError("amsgrad is not on the hot path")
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("Invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False), max_norm, grad_scale)

    def _init_state(self, p, group):
        st = self.state[p]
        if len(st) == 0:
            st['step'] = torch.tensor(0.0)
            st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st['step'] += 1
        return st['exp_avg'], st['exp_avg_sq']

    def _launch(self, lib, group, live, slots_d, bmap_d, nblocks, max_norm, sq_ptr):
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("heatnet_pub_b200.optim.Adam: the bias-correction step count is a launch argument; a captured step "
                               "would replay it unchanged -- use RMSprop inside graphs.GraphedStep, or step Adam eagerly")
        steps = {int(self.state[p]['step'].item()) for p in live}
        # one launch per distinct step count (parameters that joined later, e.g. after a phase switch, have their own)
        if len(steps) == 1:
            _lib.check(lib.hn_adam_step(slots_d.data_ptr(), bmap_d.data_ptr(), nblocks, group['lr'], group['betas'][0], group['betas'][1],
                                        group['eps'], group['weight_decay'], steps.pop(), self.grad_scale, max_norm, sq_ptr, E._stream()))
            return
        bmap = bmap_d.cpu().numpy()
        for s in sorted(steps):
            idx = [i for i, p in enumerate(live) if int(self.state[p]['step'].item()) == s]
            sel = np.ascontiguousarray(bmap[np.isin(bmap[:, 0], idx)])
            sub = torch.from_numpy(sel).to(bmap_d.device)
            _lib.check(lib.hn_adam_step(slots_d.data_ptr(), sub.data_ptr(), sel.shape[0], group['lr'], group['betas'][0], group['betas'][1],
                                        group['eps'], group['weight_decay'], s, self.grad_scale, max_norm, sq_ptr, E._stream()))


def poly_lr_scheduler(optimizer, init_lr, iter, max_iter=100, power=0.9):
    """helper/utils.py:71-84 (called at scripts/main.py:232): same signature, same arithmetic, returns the new lr."""
    lr = init_lr * (1 - iter / max_iter) ** power
    for param_group in optimizer.param_groups:
        param_group['lr'] = lr
    return lr
