"""CUDA-graph replay of a whole training step.

A conv_segnet step (two seg-net forwards, twelve critic forwards, losses, backward, optimizer) is ~2300 kernel launches issued
from Python at ~27 us each: below ~10 day/night pairs of 320x640 per GPU the step takes ~62 ms whatever the batch (the reference's
default is `--batch_size 4`, cm/train_trgb_segnet_conf.py:156).  Every buffer of the step comes from torch's caching allocator and
every kernel is enqueued on the current stream, so the step can be captured once with torch.cuda.CUDAGraph and replayed:

    step = graphs.GraphedStep(train_step, [rgb_day, ir_day, rgb_night, ir_night, label], module=model)
    loss = step(rgb_day, ir_day, rgb_night, ir_night, label)        # copies into the static inputs, replays, returns static outputs

Rules (the same as for any captured training loop):
  * `train_step(*tensors)` must do all of its work on the GPU without host synchronisation (`.item()`, `.cpu()`, printing a loss):
    return the loss tensor and read it outside.  `p.grad = None` / `optimizer.zero_grad(set_to_none=True)`, the fused losses,
    `loss.backward()` and `heatnet_pub_b200.optim.RMSprop.step()` are all fine; Adam is not (its step count is a launch argument).
  * launch arguments are frozen at capture: learning rate, loss weights, the phase (`conv_segnet.setPhase`), Dropout2d
    probabilities.  Call `recapture()` after changing any of them (StepLR changes lr once per N epochs).
  * multi-GPU: parallel.GradientReducer enqueues its bucketed ncclAllReduce calls on a communication stream ordered by CUDA
    events (no work handles, no host waits), so the exchange is captured with the rest of the step and replayed: run the step
    eagerly at least twice first (the first step of a phase learns the bucket plan; GraphedStep's own warm-up does that).
  * Python-side effects of the step happen once, at capture; the parameters and BN buffers the kernels update behind torch's back
    get their version counters bumped after every replay, so packed-weight / folded-BN caches of a later eager or eval forward
    stay coherent.
"""
from typing import Callable, Optional, Sequence

import torch

from . import engine as E


class GraphedStep:
    def __init__(self, step_fn: Callable, example_inputs: Sequence[torch.Tensor], module: Optional[torch.nn.Module] = None, warmup: int = 3):
        self.step_fn, self.module, self.warmup = step_fn, module, max(int(warmup), 2 if E.grad_arena is not None else 1)
        self.static_inputs = [t.clone() for t in example_inputs]
        self.graph = None
        self.recapture()

    def recapture(self):
        self.graph = None
        dev = self.static_inputs[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):            # allocator pools, weight packs, optimizer state and tables, smem attributes
            for _ in range(self.warmup):
                self.step_fn(*self.static_inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        l0 = E.launch_count
        E.reset_stats_pool()                     # zero-once accumulator chunks must be created (and zeroed) inside the graph
        # scratch workspaces whose addresses the captured kernels carry stay alive as long as this graph does
        self._workspaces = E.begin_capture_refs()
        try:
            with torch.cuda.graph(graph):
                self.static_outputs = self.step_fn(*self.static_inputs)
        finally:
            E.end_capture_refs()
        E.reset_stats_pool()                     # ... and must not leak into eager code afterwards
        self.launches = E.launch_count - l0
        self.graph = graph
        self._touched = (list(self.module.parameters()) + list(self.module.buffers())) if self.module is not None else []

    def __call__(self, *inputs):
        assert len(inputs) == len(self.static_inputs), "GraphedStep: same number of inputs as at capture"
        for s, t in zip(self.static_inputs, inputs):
            if t is not s:
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        E.launch_count += self.launches
        if self._touched:
            torch.autograd.graph.increment_version(self._touched)
        return self.static_outputs
