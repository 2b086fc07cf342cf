"""torch.autograd integration: one autograd.Function per network call.

The forward runs the same kernels as inference with a Tape active (engine ops record their backward
closures and keep what they need: FP32 pre-BN tensors, batch statistics, max-pool indices, dropout masks);
the backward seeds a Grads store with the incoming gradients and replays the tape in reverse.  Parameters
are passed to the Function explicitly so that `loss.backward()` populates `.grad` of the FP32 masters,
`requires_grad` toggling (conv_segnet.setPhase) prunes wgrad / dgrad work, and any torch optimizer works.
"""
from typing import Callable, List, Sequence

import torch

from . import engine as E
from .engine import Act


def needs_autograd(module: torch.nn.Module, *inputs) -> bool:
    if not torch.is_grad_enabled():
        return False
    return any(p.requires_grad for p in module.parameters()) or any(torch.is_tensor(t) and t.requires_grad for t in inputs)


class _NetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner: Callable, n_inputs: int, *args):
        inputs, params = args[:n_inputs], args[n_inputs:]
        tape = E.Tape()
        if E.grad_arena is not None:
            E.grad_arena.on_forward()
        prev = E.current_tape
        E.current_tape = tape
        try:
            in_acts, out_acts, out_tensors = runner(tape, inputs)
        finally:
            E.current_tape = prev
        ctx.tape, ctx.in_acts, ctx.out_acts = tape, in_acts, out_acts
        ctx.params, ctx.inputs_meta = params, [(t.dtype, E.act_from_view(t) is not None) if torch.is_tensor(t) else None for t in inputs]
        ctx.n_inputs = n_inputs
        return tuple(out_tensors)

    @staticmethod
    def backward(ctx, *douts):
        grads = E.Grads()
        if E.grad_arena is not None:
            E.grad_arena.on_backward()
        for act, d in zip(ctx.out_acts, douts):
            if d is None or act is None:
                continue
            g, inited = grads.target(act)
            view = E.act_from_view(d)
            if view is not None:
                E.accumulate(view, g, inited)
            else:
                tmp = E.from_nchw(d, g.dtype, None if inited else g)
                if inited:
                    E.accumulate(tmp, g, True)
            grads.mark(act)
        ctx.tape.backward(grads)
        in_grads: List = []
        for act, meta, need in zip(ctx.in_acts, ctx.inputs_meta, ctx.needs_input_grad[2:2 + ctx.n_inputs]):
            g = grads.get(act) if (need and act is not None) else None
            if g is None:
                in_grads.append(None)
            elif meta is not None and meta[1] and meta[0] == g.dtype:
                in_grads.append(g.nchw())                       # the input was one of our NHWC views: same layout back
            else:
                in_grads.append(E.to_nchw_f32(g).to(meta[0]))
        # gradients written into the arena of a parallel.GradientReducer are bound to .grad here (a view of the parameter's slot:
        # no copy, static address) and torch.autograd gets None for them; the rest travels back through autograd as usual
        for p, slot in grads.arena_params.items():
            if p.grad is None or p.grad.data_ptr() != slot.data_ptr():
                p.grad = slot
        p_grads = [grads.params.get(p) if p.requires_grad else None for p in ctx.params]
        ctx.tape = ctx.in_acts = ctx.out_acts = None
        return (None, None, *in_grads, *p_grads)


def apply(runner: Callable, inputs: Sequence, module: torch.nn.Module):
    params = [p for p in module.parameters() if p.requires_grad]
    return _NetFunction.apply(runner, len(inputs), *inputs, *params)
