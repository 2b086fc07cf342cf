#!/usr/bin/env python
"""HeatNet hot-path benchmark (driver contract: one JSON line on stdout from rank 0).

Metric (BASELINE.json): RGB+thermal segmentation images/sec.  A "step" is one forward of the two-stream
PSPNet-ResNet50 over one synthetic batch; at N=1 the workload is BASELINE.json configs[1]: BF16, batch 16,
650x1920 frames (logits 16x13x656x1920).  N>1: the path shards over images -- every rank runs the same
batch size on its own GPU, no data-path collective ("scaling": "weak").

  value     images/s with inputs resident in HBM, CUDA-event timed, max over ranks.
  e2e       same metric through the public module API from PINNED HOST buffers: H2D of the RGB+IR frames,
            forward, class argmax on the device, D2H of the uint8 label maps -- all inside the timed region.
  roofline  the tensor-core convolution kernel (conv_tc_kernel): algorithmic conv FLOPs per step (2*MAC of
            every nn.Conv2d the reference executes) / device time summed over the
            step's conv launches (CUDA events on the launching stream, separate instrumented pass).
  cpu_baseline  the reference's own CPU path (baseline/_ref when installed, else the oracle port) timed on the
            box's host cores on a bounded sample of the same workload.
  gpu_torch_baseline  stock PyTorch on the same B200 (the reference graph with torch / cuDNN: channels_last,
            BF16 autocast, cudnn.benchmark) on the same batch -- the bar SURVEY.md section 2.2 names.

The same run ALSO measures, after the headline, the other BASELINE configs at the same N and reports them inside
`config` / `legs` (the headline stays configs[1] so that N=1 of a scaling run equals the single-GPU bench):
  train_seg / train_critic   configs[2]/[3]: one adversarial conv_segnet step per rank on 16 day+night pairs of
            320x640 -- forward, fused losses, backward, gradient all-reduce over NCCL (overlapped with backward,
            replayed inside the step's CUDA graph), fused RMSprop -- images/s, ms/step, exposed all-reduce time;
  iou_eval  configs[4]: IoU(14).add over 500 label maps of 320x640 per step.
`--workload train_seg|train_critic|iou_eval` runs one of them alone as its own bench line; `--legs none` skips them.

`--impl reference` times the reference's CPU path as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line.  Libraries print to fd 1 behind Python's back (NCCL's "NCCL version ..." banner is a raw
# printf), so fd 1 is pointed at stderr for the whole run and the result line is written to a private duplicate of the real stdout.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


METRIC = "rgb_thermal_seg_images_per_sec"
UNIT = "images/s"
SEED = 1203412412           # the reference's own seed, scripts/main.py:126


def conv_flops_per_image(net, h, w, late=True):
    """2*MAC of every convolution of the forward at input h x w, from the module tree (matches SURVEY appendix A:
    325.84 GFLOP at 320x640, 2001.24 GFLOP at 650x1920 for late fusion)."""
    import torch.nn as nn

    def out(n, k, s, p, d):
        return (n + 2 * p - d * (k - 1) - 1) // s + 1

    total = 0

    def conv(m, hh, ww):
        nonlocal total
        k, s, p, d = m.kernel_size[0], m.stride[0], m.padding[0], m.dilation[0]
        ho, wo = out(hh, k, s, p, d), out(ww, k, s, p, d)
        total += 2 * ho * wo * m.out_channels * m.in_channels * k * k
        return ho, wo

    f = net.feats
    streams = [(f.conv1, f.layer1, f.layer2)]
    if f.late_fusion:
        streams.append((f.conv1_2, f.layer1_2, f.layer2_2))
    for c1, l1, l2 in streams:
        hh, ww = conv(c1, h, w)
        hh, ww = out(hh, 3, 2, 1, 1), out(ww, 3, 2, 1, 1)
        for layer in (l1, l2):
            for blk in layer:
                conv(blk.conv1, hh, ww)
                h2, w2 = conv(blk.conv2, hh, ww)
                conv(blk.conv3, h2, w2)
                if blk.downsample is not None:
                    conv(blk.downsample[0], hh, ww)
                hh, ww = h2, w2
    for layer in (f.layer3, f.layer4):
        for blk in layer:
            conv(blk.conv1, hh, ww)
            conv(blk.conv2, hh, ww)
            conv(blk.conv3, hh, ww)
            if blk.downsample is not None:
                conv(blk.downsample[0], hh, ww)
    for st in net.psp.stages:
        s = st[0].output_size
        s = s if isinstance(s, int) else s[0]
        conv(st[1], s, s)
    conv(net.psp.bottleneck, hh, ww)
    for up in (net.up_1, net.up_2, net.up_3):
        hh, ww = 2 * hh, 2 * ww
        conv(up.conv[0], hh, ww)
    conv(net.final[0], hh, ww)
    return total


def he_init_(net, seed=0):
    """Random-init weights of the named architecture (no checkpoints offline): fan-in He init for convs, BN
    gamma ~ U(.8,1.2) (x0.5 on block-closing bn3), running stats near (0,1) -- keeps eval-mode activations O(1)."""
    import torch
    import torch.nn as nn
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, m in net.named_modules():
            if isinstance(m, nn.Conv2d):
                fan_in = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
            elif isinstance(m, nn.BatchNorm2d):
                gain = 0.5 if name.endswith("bn3") else 1.0
                m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 0.4 + 0.8) * gain)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
                m.running_mean.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
                m.running_var.copy_(torch.rand(m.bias.shape, generator=g) * 0.8 + 0.6)
    return net


def synthetic_batch(batch, h, w, seed=SEED):
    """RGB / IR in U(-1,1), the value range the reference's loaders produce (thermal_loader.py:649-659)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(batch, 3, h, w, generator=g) * 2 - 1
    ir = torch.rand(batch, 1, h, w, generator=g) * 2 - 1
    return rgb, ir


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, reasons, smax, pw = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax = float(parts[2]); pw.append(float(parts[3]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm), power_w_max=max(pw))
        return out


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ---------------------------------------------------------------------------------------------------- process context
class Ctx:
    """One process per GPU (torchrun): rank / world / device, barrier + synchronize, max-over-ranks of a CUDA-event interval."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K calls of fn bracketed by barrier + synchronize on both sides -> elapsed ms, max over ranks."""
        import torch
        import torch.distributed as dist
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        out = None
        for _ in range(steps):
            out = fn()
        ev1.record()
        self.barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), out

    def close(self):
        import torch.distributed as dist
        if self.world > 1 and dist.is_initialized():
            dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------- CPU arms
def reference_cpu_forward(h, w):
    """-> (callable running ONE eval forward of one h x w frame on the CPU, kind, keep): the UNMODIFIED reference module from
    baseline/_ref when installed ("reference"), else the oracle restatement ("port").  Same recipe weights and frame."""
    import torch
    from oracle import heatnet_oracle as O
    from oracle import reference_loader as RL
    sd = O.recipe_fill(O.pspnet_state_dict(True, 4), seed=0)
    rgb, ir = O.synthetic_inputs(1, h, w)
    keep = dict(sd=sd, rgb=rgb, ir=ir)
    PSPNet = None
    try:
        PSPNet = RL.load_cm_pspnet()
    except Exception:
        PSPNet = None
    if PSPNet is not None:
        import warnings
        warnings.filterwarnings("ignore", message=".*upsample.*")
        net = PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50', in_channels=4, pretrained=False,
                     late_fusion=True)
        net.load_state_dict(sd)
        net.eval()                                    # scripts/inference.py / the eval path of config 1

        def run():
            with torch.no_grad():
                keep["logits"] = net(rgb, ir)[0]
        return run, "reference", keep

    def run():
        with torch.no_grad():
            keep["logits"] = O.pspnet_forward(sd, rgb, ir, late_fusion=True, training=False)[0]
    return run, "port", keep


def cpu_forward_images_per_sec(h, w, images, threads=None, keep=None):
    """The reference's CPU path on `images` frames of h x w, one at a time.  `keep` (a dict) receives the weights, the input
    frame and the CPU logits for the parity check of the same run."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    run, kind, kp = reference_cpu_forward(h, w)
    t0 = time.perf_counter()
    for _ in range(images):
        run()
    dt = time.perf_counter() - t0
    if keep is not None:
        keep.update(kp)
    return images / dt, torch.get_num_threads(), kind


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (the unmodified modules from baseline/_ref; the oracle
    port only if they did not travel) on all host cores, one frame per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count())
    h, w = args.height, args.width
    run, kind, _ = reference_cpu_forward(h, w)
    steps, warmup = args.steps, args.warmup
    budget_s = 240.0
    t0 = time.perf_counter()
    run()
    first = time.perf_counter() - t0
    done_w = 1
    while done_w < warmup and (done_w + 1) * first < 0.3 * budget_s:
        run()
        done_w += 1
    k = max(1, min(steps, int((budget_s - done_w * first) / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(k):
        run()
    dt = time.perf_counter() - t0
    v = k / dt
    what = ("the reference's own PSPNet module (baseline/_ref, unmodified, .eval())" if kind == "reference"
            else "oracle.pspnet_forward (restatement of the reference; baseline/_ref not present)")
    sample = f"{k} timed steps of 1 frame {h}x{w} each through {what}, FP32, torch CPU, {done_w} warm-up; steps capped to fit ~4 min"
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": k, "warmup": done_w,
            "ms_per_step": 1000.0 * dt / k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"PSPNet-ResNet50 RGB+thermal late-fusion forward, {h}x{w}, 1 frame per step on CPU"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------- training step
TRAIN_GFLOP_PER_PAIR_320x640 = {"train_seg": 2051.2, "train_critic": 771.4}


def measure_train(ctx, args, workload, B, H, W, steps, warmup):
    """One adversarial step of cm/train_trgb_segnet_conf.py:428-568 per rank on its shard: conv_segnet forward on a day and a
    night batch, the phase's loss, backward with the bucketed NCCL gradient all-reduce overlapped (parallel.GradientReducer),
    fused RMSprop.  -> dict (rank 0 fills the derived fields)."""
    import contextlib
    import io
    import torch
    import torch.nn as nn
    from heatnet_pub_b200 import conf_segnet, engine as E, graphs, losses, optim, parallel

    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    with contextlib.redirect_stdout(io.StringIO()):
        model = conf_segnet.conv_segnet(pretrained=False, disc_arch='cyclegan', num_critics=6, no_conf=False, modalities='ir_rgb',
                                        arch='pspnet', late_fusion=True)
        he_init_(model.trgb_segnet)
        model = model.to(dev).train()
        model.setPhase(workload)
    model.trgb_segnet.set_precision(args.precision)
    for c in model.critics:
        c.precision = args.precision
    parallel.broadcast_parameters(model, 0)
    reducer = parallel.GradientReducer(model.parameters(), bucket_mb=args.bucket_mb)
    g = torch.Generator().manual_seed(SEED + rank)
    mk = lambda c: (torch.rand(B, c, H, W, generator=g) * 2 - 1).to(dev)
    rgb_d, ir_d, rgb_n, ir_n = mk(3), mk(1), mk(3), mk(1)
    label = torch.randint(0, 13, (B, H, W), generator=g).to(dev)
    if args.torch_losses:
        mse, ce = nn.MSELoss(), nn.CrossEntropyLoss()
        tgt = lambda c, v: torch.full_like(c, v)
    else:       # fused forward+backward criteria (hn_ce_loss_fwd_bwd / hn_critic_loss_fwd_bwd); constant targets as floats
        mse, ce = losses.MSELoss(), losses.CrossEntropyLoss()
        tgt = lambda c, v: float(v)
    # cm/train_trgb_segnet_conf.py:270: ONE RMSprop over all parameters; the frozen half has no .grad and is skipped
    optimizer = None if args.no_optimizer else optim.RMSprop(model.parameters(), lr=1e-5)

    def train_step(rgb_d, ir_d, rgb_n, ir_n, label):
        reducer.zero_grad()
        o = model([rgb_d, ir_d], [rgb_n, ir_n])
        if workload == "train_seg":
            conf = sum(mse(c, tgt(c, 1)) for c in o['critics_a']) + sum(mse(c, tgt(c, 1)) for c in o['critics_b'])
            total = ce(o['pred_label_a'], label) + 0.1 * conf
        else:
            total = sum(mse(c, tgt(c, 1)) for c in o['critics_a']) + sum(mse(c, tgt(c, 0)) for c in o['critics_b'])
        total.backward()
        reducer.finish()                 # gradients are rank averages from here on (no-op on one GPU)
        if optimizer is not None:
            optimizer.step()
        return total

    batch = (rgb_d, ir_d, rgb_n, ir_n, label)
    # --cuda-graph on / auto: the WHOLE step -- forward, losses, backward, the all-reduce on its communication stream, optimizer --
    # is captured once and replayed (graphs.GraphedStep): a step is ~2300 launches, Python-bound below ~10 pairs per GPU
    use_graph = args.cuda_graph in ("on", "auto") and not args.torch_losses and not (args.layer_table and workload == args.workload)

    def make_step():
        if not use_graph:
            for _ in range(max(warmup, 3)):
                train_step(*batch)
            return lambda: train_step(*batch)
        gstep = graphs.GraphedStep(train_step, batch, module=model, warmup=max(warmup, 3))
        return lambda: gstep(*batch)     # copies the batch into the graph's static inputs, replays, returns the static loss

    step = make_step()
    for _ in range(warmup):
        loss = step()
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    l0 = E.launch_count
    prof_range = bool(os.environ.get("HN_PROFILE_RANGE"))      # `ncu --profile-from-start off`: only the timed steps are profiled
    if prof_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ms, loss = ctx.timed(step, steps)
    if prof_range:
        torch.cuda.profiler.stop()
    launches = E.launch_count - l0
    clocks = sampler.stop() if sampler else None
    buckets, nbytes, copied = reducer.last_buckets, reducer.last_bytes, reducer.copied
    # the same step with the exchange switched off (gradients stay local): the difference is the EXPOSED all-reduce time
    ms_local = None
    if world > 1:
        reducer.enabled = False
        step_local = make_step()
        step_local()
        ms_local, _ = ctx.timed(step_local, steps)
        reducer.enabled = True
    if rank == 0 and args.layer_table and workload == args.workload:
        # per C-ABI-call device timeline of ONE step (needs HN_TIMELINE=1 at start-up), aggregated by call + geometry
        from heatnet_pub_b200 import _lib as L
        if not os.environ.get("HN_TIMELINE"):
            raise SystemExit("--layer-table on a training workload needs HN_TIMELINE=1 in the environment")
        L.timeline = []
        step()
        torch.cuda.synchronize()
        recs, L.timeline = L.timeline, None
        agg = {}
        for name, desc, a, b in recs:
            e = agg.setdefault((name, desc), [0, 0.0])
            e[0] += 1
            e[1] += a.elapsed_time(b)
        rows = [{"call": k[0], "what": k[1], "launches_per_step": v[0], "ms_per_step": v[1]} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
        by_call = {}
        for r in rows:
            by_call[r["call"]] = by_call.get(r["call"], 0.0) + r["ms_per_step"]
        json.dump({"by_call_ms": dict(sorted(by_call.items(), key=lambda kv: -kv[1])), "rows": rows}, open(args.layer_table, "w"), indent=1)
    res = None
    if rank == 0:
        peaks, peak_src = load_peaks()
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        scale = (H * W) / (320.0 * 640.0)
        tflops = TRAIN_GFLOP_PER_PAIR_320x640[workload] * scale * B * steps / (ms / 1000.0) / 1e3
        res = {"images_per_s": 2 * B * world * steps / (ms / 1000.0), "ms_per_step": ms / steps, "steps": steps, "per_gpu_pairs": B,
               "global_pairs": B * world, "height": H, "width": W, "cuda_graph": bool(use_graph), "loss": float(loss.detach()), "gpu_launches": launches,
               "tflops_per_gpu": tflops, "frac_of_sustained_peak": tflops / peak, "frac_of_burst_peak": tflops / peaks["bf16_tflops"],
               "peak_source": peak_src, "clocks": clocks,
               "what": f"conv_segnet {workload} step (PSPNet-ResNet50 late fusion + 6 FCDiscriminator critics), {B} day+night pairs per GPU at "
                       f"{H}x{W}: fwd + {'torch' if args.torch_losses else 'fused'} losses + bwd + gradient all-reduce"
                       f"{' (optimizer excluded)' if args.no_optimizer else ' + fused RMSprop step'}; 2 images per pair",
               "allreduce": ({"ranks": world, "transport": "ncclAllReduce(avg) on a communication stream, per bucket, event-ordered behind the "
                              "bucket's last wgrad; inside the CUDA graph" if use_graph else "ncclAllReduce(avg), event-ordered, eager",
                              "mbytes_per_step": nbytes / 1e6, "buckets": buckets, "gradients_copied_into_arena": copied,
                              "ms_per_step_without_exchange": ms_local / steps, "exposed_ms": (ms - ms_local) / steps}
                             if world > 1 else None)}
    del model, reducer, optimizer, step
    from heatnet_pub_b200 import engine as _E
    _E.grad_arena = None
    torch.cuda.empty_cache()
    return res


def run_train(args):
    ctx = Ctx()
    B = args.batch
    H, W = (args.height, args.width) if (args.height, args.width) != (650, 1920) else (320, 640)
    r = measure_train(ctx, args, args.workload, B, H, W, args.steps, args.warmup)
    if ctx.rank == 0:
        line = {"metric": "rgb_thermal_seg_train_images_per_sec", "value": r["images_per_s"], "unit": UNIT, "n_gpus": ctx.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": {"workload": r["what"], "per_gpu_pairs": B, "global_pairs": B * ctx.world, "cuda_graph": r["cuda_graph"],
                           "parallelism": (f"batch-sharded x{ctx.world}, NCCL all-reduce of {r['allreduce']['mbytes_per_step']:.1f} MB in "
                                           f"{r['allreduce']['buckets']} buckets overlapped with backward" if ctx.world > 1 else "single GPU"),
                           "allreduce": r["allreduce"]},
                "loss": r["loss"], "gpu_launches": r["gpu_launches"], "clocks": r["clocks"],
                "roofline": {"bound": "tensor", "achieved": r["tflops_per_gpu"], "peak": load_peaks()[0].get("bf16_tflops_sustained"),
                             "unit": "TFLOP/s", "frac": r["frac_of_sustained_peak"], "frac_of_burst_peak": r["frac_of_burst_peak"],
                             "note": "whole step (all kernels, not only convs) against algorithmic conv FLOPs of SURVEY.md section 8d -- the graph as the "
                                     "reference writes it; the BF16 step runs the PSP bottleneck in projected form (2048 instead of 10240 input "
                                     "channels in forward, dgrad and wgrad), so executed FLOPs are ~16 % lower than algorithmic",
                             "peak_source": r["peak_source"]}}
        emit(line)
    ctx.close()


# ---------------------------------------------------------------------------------------------------- iou_eval (config 5)
def measure_iou_eval(ctx, args, maps, steps, warmup, cpu_baseline=False):
    """BASELINE.json configs[4]: iou_eval.IoU(14, ignore [12,13]) over synthetic 320x640 int64 label maps in chunks of
    `maps` (default 500; 20 steps = 10 000 maps = 2.048e9 pixels).  A step = IoU.add(pred, target) on one chunk.
    value: label maps/s with the chunk resident in HBM (includes the K*K D2H + host int32 accumulate every step);
    e2e: the same call on pinned HOST int64 tensors (H2D of 16 B/pixel inside the timed region);
    roofline: hn_confusion alone, 16 algorithmic bytes per pixel, against the measured HBM copy bandwidth."""
    import ctypes as C
    import numpy as np
    import torch
    from heatnet_pub_b200 import _lib, iou_eval
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    K, H, W = 14, 320, 640
    g = torch.Generator(device=dev).manual_seed(SEED + rank)
    pred = torch.randint(0, K, (maps, H, W), generator=g, device=dev)
    target = torch.randint(0, K, (maps, H, W), generator=g, device=dev)
    npix = maps * H * W
    metric = iou_eval.IoU(K, False, [12, 13])
    for _ in range(warmup):
        metric.add(pred, target)
    metric.reset()
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    ms, _ = ctx.timed(lambda: metric.add(pred, target), steps)
    clocks = sampler.stop() if sampler else None
    # size-independent property: every pixel counted exactly once per step (modulo 2^32: the reference's accumulator is int32)
    assert int(metric.conf_metric.conf.astype(np.uint32).sum(dtype=np.uint64)) % (1 << 32) == (npix * steps) % (1 << 32)
    iou, miou = metric.value()

    # kernel alone (no D2H): CUDA events on the launching stream
    lib = _lib.load()
    out = torch.zeros(K * K + 1, dtype=torch.int64, device=dev)
    flags = out[K * K:].view(torch.int32)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    kern = lambda: _lib.check(lib.hn_confusion(pred.data_ptr(), None, 1, npix, target.data_ptr(), K, out.data_ptr(), flags.data_ptr(), st))
    for _ in range(3):
        kern()
    kms = ctx.timed(kern, steps)[0] / steps
    assert int(out[:K * K].sum().item()) == npix * (steps + 3)

    # e2e from pinned host memory
    pred_h, target_h = pred.cpu().pin_memory(), target.cpu().pin_memory()
    metric.reset()
    metric.add(pred_h, target_h)
    e2e_ms, _ = ctx.timed(lambda: metric.add(pred_h, target_h), steps)
    res = None
    if rank == 0:
        peaks, peak_src = load_peaks()
        hbm = peaks.get("hbm_gbs", 6650.0)
        achieved = 16.0 * npix / (kms / 1e3) / 1e9
        cpu = None
        if cpu_baseline:
            from oracle.iou_oracle import IoUOracle
            o = IoUOracle(K, False, [12, 13])
            n_cpu = min(maps, 200)
            ph, th = pred_h[:n_cpu].numpy(), target_h[:n_cpu].numpy()
            t0 = time.perf_counter()
            o.add(ph, th)
            dt = time.perf_counter() - t0
            cpu = {"value": n_cpu / dt, "unit": "label maps/s", "cores": 1, "kind": "port",
                   "sample": f"{n_cpu} maps of {H}x{W} through oracle.iou_oracle.IoUOracle.add (numpy bincount, the reference's arithmetic)"}
        res = {"maps_per_s": maps * world * steps / (ms / 1e3), "ms_per_step": ms / steps, "steps": steps, "maps_per_step": maps, "clocks": clocks,
               "miou": float(miou), "cpu_baseline": cpu,
               "what": f"iou_eval.IoU(14, ignore [12,13]).add over {maps} pred/target int64 label maps of {H}x{W} per step per GPU "
                       f"({steps} steps = {maps * steps} maps per GPU)",
               "e2e": {"value": maps * world * steps / (e2e_ms / 1e3), "unit": "label maps/s", "h2d_bytes_per_step": 16 * npix,
                       "d2h_bytes_per_step": 8 * (K * K + 1), "what": "IoU.add on pinned host int64 tensors: H2D + histogram + K*K D2H per step"},
               "roofline": {"kernel": "confusion_labels_tma_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                            "traffic": None, "kernel_ms": kms, "algorithmic_bytes_per_pixel": 16, "peak_source": peak_src}}
    del pred, target, pred_h, target_h
    torch.cuda.empty_cache()
    return res


def run_iou_eval(args):
    ctx = Ctx()
    maps = args.batch if args.batch != 16 else 500
    r = measure_iou_eval(ctx, args, maps, args.steps, args.warmup, cpu_baseline=(ctx.world == 1 and not args.no_cpu_baseline))
    if ctx.rank == 0:
        line = {"metric": "iou_eval_label_maps_per_sec", "value": r["maps_per_s"], "unit": "label maps/s", "n_gpus": ctx.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "config": {"workload": r["what"], "maps_per_step": maps, "parallelism": f"map-sharded x{ctx.world}, no collective",
                           "l2": "each step streams 1.6 GB of labels (>> 126 MB L2); no explicit flush"},
                "e2e": r["e2e"], "gpu_launches": args.steps, "clocks": r["clocks"], "miou": r["miou"], "roofline": r["roofline"],
                "cpu_baseline": r["cpu_baseline"]}
        emit(line)
    ctx.close()


# ---------------------------------------------------------------------------------------------------- stock PyTorch on the same GPU
def measure_torch_gpu_baseline(ctx, args, B, H, W, ours_value, train_legs):
    """The bar SURVEY.md section 2.2 names: the reference graph executed by stock PyTorch (ATen + cuDNN) on this B200 -- state dict
    on the device, channels_last, torch.autocast(bfloat16), cudnn.benchmark -- on the same synthetic batch; plus one train_seg
    step of the same graph with torch autograd + torch.optim.RMSprop.  Reported next to cpu_baseline, never on the product path."""
    import torch
    import torch.nn.functional as F
    from oracle import heatnet_oracle as O
    dev = ctx.dev
    out = {"what": "oracle graph (= the reference's modules as torch functional calls) on cuda: channels_last weights and inputs, "
                   "torch.autocast(bfloat16), cudnn.benchmark=True, TF32 irrelevant under autocast", "torch": torch.__version__}
    old_bench = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    cl = lambda t: t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t
    try:
        sd = {k: cl(v.to(dev)) for k, v in O.recipe_fill(O.pspnet_state_dict(True, 4), seed=0).items()}
        rgb, ir = synthetic_batch(B, H, W)
        rgb, ir = cl(rgb.to(dev)), cl(ir.to(dev))

        def fwd():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return O.pspnet_forward(sd, rgb, ir, late_fusion=True, training=False)[0]

        for _ in range(3):
            fwd()
        n = max(3, min(args.steps, 5))
        ms, _ = ctx.timed(fwd, n)
        v = B * n / (ms / 1e3)
        out["infer"] = {"value": v, "unit": UNIT, "ms_per_step": ms / n, "steps": n, "batch": B, "height": H, "width": W,
                        "ours_over_torch": ours_value / v}
        del sd, rgb, ir
        torch.cuda.empty_cache()
    except Exception as e:
        out["infer"] = {"value": None, "failed": repr(e)[:300]}
    # train_seg step of the oracle graph at the training leg's shape
    ts = (train_legs or {}).get("train_seg")
    if ts and "images_per_s" in ts:
        Bt, Ht, Wt = ts["per_gpu_pairs"], ts["height"], ts["width"]
        for pairs in (Bt, Bt // 2, Bt // 4):
            if pairs < 1:
                break
            try:
                sd = {k: cl(v.to(dev)) for k, v in O.recipe_fill(O.conf_segnet_state_dict(True, 6), seed=0).items()}
                live = [v.requires_grad_(True) for k, v in sd.items() if k.startswith("trgb_segnet.") and v.is_floating_point() and "running_" not in k]
                opt = torch.optim.RMSprop(live, lr=1e-5)
                g = torch.Generator().manual_seed(SEED)
                mk = lambda c: cl((torch.rand(pairs, c, Ht, Wt, generator=g) * 2 - 1).to(dev))
                day, night = [mk(3), mk(1)], [mk(3), mk(1)]
                label = torch.randint(0, 13, (pairs, Ht, Wt), generator=g).to(dev)

                def step():
                    opt.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        o = O.conf_segnet_forward(sd, day, night, training=True, dropout=True)
                    total, _, _ = O.train_seg_loss({k: ([t.float() for t in v] if isinstance(v, list) else (v.float() if torch.is_tensor(v) else v))
                                                    for k, v in o.items()}, label)
                    total.backward()
                    opt.step()
                    return total

                for _ in range(3):
                    step()
                n = 3
                ms, _ = ctx.timed(step, n)
                v = 2 * pairs * n / (ms / 1e3)
                out["train_seg"] = {"value": v, "unit": UNIT, "ms_per_step": ms / n, "steps": n, "pairs": pairs, "height": Ht, "width": Wt,
                                    "ours_over_torch": ts["images_per_s"] / ctx.world / v,
                                    "what": "fwd (autocast bf16) + CE + 0.1 * 12 MSE + torch autograd bwd + torch.optim.RMSprop (foreach), eager"}
                del sd, live, opt, day, night, label
                torch.cuda.empty_cache()
                break
            except torch.OutOfMemoryError:
                out["train_seg"] = {"value": None, "failed": f"out of memory at {pairs} pairs"}
                sd = live = opt = day = night = label = None
                torch.cuda.empty_cache()
            except Exception as e:
                out["train_seg"] = {"value": None, "failed": repr(e)[:300]}
                break
    torch.backends.cudnn.benchmark = old_bench
    return out


# ---------------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--height", type=int, default=650)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-logits-e2e", action="store_true", help="skip the second e2e variant that copies the FP32 logits back to the host")
    ap.add_argument("--no-optimizer", action="store_true", help="training workloads: leave the optimizer step out of the timed step")
    ap.add_argument("--torch-losses", action="store_true", help="training workloads: torch criteria instead of the fused loss kernels")
    ap.add_argument("--cuda-graph", default="auto", choices=["auto", "on", "off"],
                    help="infer workload: replay the forward as a CUDA graph (PSPNet.set_cuda_graph).  auto = on for launch-bound "
                         "shapes (at most 8 images of 320x640 per step), off for the headline full-frame batch, which is GPU-bound")
    ap.add_argument("--layer-table", default=None, help="write the per-conv-launch timing table (JSON) here")
    ap.add_argument("--workload", default="infer", choices=["infer", "train_seg", "train_critic", "iou_eval"],
                    help="infer = the headline (BASELINE configs[1]) followed by the train_seg / train_critic / iou_eval legs; train_* = one "
                         "adversarial training step (configs[2]/[3]) alone: per-GPU batch of --batch day+night pairs at --height x --width, "
                         "fwd + bwd + NCCL gradient all-reduce + optimizer")
    ap.add_argument("--legs", default="all", help="infer workload: comma list of the extra legs measured after the headline "
                    "(train_seg,train_critic,iou_eval,torch) or all / none")
    ap.add_argument("--train-pairs", type=int, default=16, help="day+night pairs per GPU of the training legs (320x640)")
    ap.add_argument("--bucket-mb", type=float, default=32.0, help="gradient all-reduce bucket size")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    if args.workload == "iou_eval":
        return run_iou_eval(args)
    if args.workload != "infer":
        return run_train(args)

    import ctypes as C
    import torch
    import torch.distributed as dist

    ctx = Ctx()
    rank, world, local_rank, dev = ctx.rank, ctx.world, ctx.local_rank, ctx.dev

    import __graft_entry__ as ge
    if rank == 0 and not os.path.exists(ge.LIB):
        ge.build()
    if world > 1:
        dist.barrier()
    from heatnet_pub_b200 import _lib, engine as E, pspnet

    B, H, W = args.batch, args.height, args.width
    net = pspnet.PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50', in_channels=4,
                        pretrained=False, late_fusion=True)
    he_init_(net)
    net = net.to(dev).eval().set_precision(args.precision)
    use_graph = args.cuda_graph == "on" or (args.cuda_graph == "auto" and B * H * W <= 8 * 320 * 640)
    if use_graph:
        net.set_cuda_graph(True)
    rgb_h, ir_h = synthetic_batch(B, H, W, seed=SEED + rank)
    rgb_h, ir_h = rgb_h.pin_memory(), ir_h.pin_memory()
    rgb, ir = rgb_h.to(dev), ir_h.to(dev)
    flops_img = conv_flops_per_image(net, H, W)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        with torch.no_grad():
            logits, taps, _ = net(rgb, ir)
        return logits

    for _ in range(args.warmup):
        out = step()
    del out
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = E.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    prof_range = bool(os.environ.get("HN_PROFILE_RANGE"))      # `ncu --profile-from-start off`: only the timed steps are profiled
    if prof_range:
        torch.cuda.profiler.start()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    barrier()
    if prof_range:
        torch.cuda.profiler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = E.launch_count - l0
    clocks = sampler.stop() if sampler else None
    Hout, Wout = out.shape[2], out.shape[3]
    del out
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    value = world * B * args.steps / (ms_max / 1000.0)

    # ---- e2e: pinned host frames -> H2D -> forward -> device argmax -> D2H uint8 label maps
    lib = _lib.load()
    labels_h = torch.empty((B, Hout, Wout), dtype=torch.uint8).pin_memory()
    # Two sets of device buffers and a copy stream: the H2D of step i+1 and the D2H of step i-1 overlap the forward of
    # step i, as a serving loop would do.  Every step still moves its own frames in and its own label maps out inside
    # the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [{"rgb": torch.empty_like(rgb), "ir": torch.empty_like(ir), "labels": torch.empty((B, Hout, Wout), dtype=torch.uint8, device=dev),
             "in_ready": torch.cuda.Event(), "out_ready": torch.cuda.Event(), "consumed": torch.cuda.Event()} for _ in range(2)]
    main_stream = torch.cuda.current_stream()

    def stage_in(b):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(b["consumed"])            # the forward that last read these buffers has finished
            b["rgb"].copy_(rgb_h, non_blocking=True)
            b["ir"].copy_(ir_h, non_blocking=True)
            b["in_ready"].record(copy_stream)

    def compute(b):
        main_stream.wait_event(b["in_ready"])
        with torch.no_grad():
            logits, _, _ = net(b["rgb"], b["ir"])
        _lib.check(lib.hn_argmax_labels(logits.data_ptr(), B, Hout * Wout, logits.shape[1], b["labels"].data_ptr(), None,
                                        C.c_void_p(main_stream.cuda_stream)))
        b["consumed"].record(main_stream)
        b["out_ready"].record(main_stream)

    def stage_out(b):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(b["out_ready"])
            labels_h.copy_(b["labels"], non_blocking=True)

    def e2e_run(nsteps):
        for b in bufs:
            b["consumed"].record(main_stream)
        stage_in(bufs[0])
        for i in range(nsteps):
            cur = bufs[i & 1]
            if i + 1 < nsteps:
                stage_in(bufs[(i + 1) & 1])
            compute(cur)
            stage_out(cur)
        main_stream.wait_stream(copy_stream)

    e2e_run(2)
    barrier()
    ev0.record()
    e2e_run(args.steps)
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (t.item() / 1000.0)
    h2d = rgb_h.numel() * 4 + ir_h.numel() * 4
    d2h = labels_h.numel()

    # ---- the same serving loop fed with RAW camera frames: uint8 RGB (N,H,W,3) + uint16 IR (N,H,W) in pinned host memory,
    # normalised on the device by heatnet_pub_b200.inputs (the loaders' arithmetic, cm/thermal_loader.py:649-659,715-728)
    from heatnet_pub_b200 import inputs as hn_inputs
    gq = torch.Generator().manual_seed(SEED + rank)
    rgb8_h = torch.randint(0, 256, (B, H, W, 3), generator=gq, dtype=torch.uint8).pin_memory()
    ir16_h = torch.randint(21000, 26000, (B, H, W), generator=gq, dtype=torch.int32).to(torch.int16).pin_memory()   # counts < 2^15: int16 == uint16 bits
    raw_bufs = [{"rgb": torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev), "ir": torch.empty((B, H, W), dtype=torch.int16, device=dev),
                 "labels": torch.empty((B, Hout, Wout), dtype=torch.uint8, device=dev),
                 "in_ready": torch.cuda.Event(), "out_ready": torch.cuda.Event(), "consumed": torch.cuda.Event()} for _ in range(2)]

    def raw_stage_in(b):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(b["consumed"])
            b["rgb"].copy_(rgb8_h, non_blocking=True)
            b["ir"].copy_(ir16_h, non_blocking=True)
            b["in_ready"].record(copy_stream)

    def raw_compute(b):
        main_stream.wait_event(b["in_ready"])
        with torch.no_grad():
            logits, _, _ = net(hn_inputs.prepare_rgb(b["rgb"], precision=args.precision), hn_inputs.prepare_ir(b["ir"], precision=args.precision))
        _lib.check(lib.hn_argmax_labels(logits.data_ptr(), B, Hout * Wout, logits.shape[1], b["labels"].data_ptr(), None,
                                        C.c_void_p(main_stream.cuda_stream)))
        b["consumed"].record(main_stream)
        b["out_ready"].record(main_stream)

    def raw_run(nsteps):
        for b in raw_bufs:
            b["consumed"].record(main_stream)
        raw_stage_in(raw_bufs[0])
        for i in range(nsteps):
            cur = raw_bufs[i & 1]
            if i + 1 < nsteps:
                raw_stage_in(raw_bufs[(i + 1) & 1])
            raw_compute(cur)
            stage_out(cur)
        main_stream.wait_stream(copy_stream)

    raw_run(2)
    barrier()
    ev0.record()
    raw_run(args.steps)
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_raw = {"value": world * B * args.steps / (t.item() / 1000.0), "unit": UNIT, "h2d_bytes_per_step": rgb8_h.numel() + 2 * ir16_h.numel(),
               "d2h_bytes_per_step": d2h, "what": "same loop from raw frames: pinned uint8 RGB + uint16 IR -> H2D -> device normalisation "
               "(heatnet_pub_b200.inputs, zero-copy into the stems) -> forward -> argmax -> D2H uint8 label maps"}
    del raw_bufs

    # ---- module-API e2e: the call's own result -- the FP32 logits (validation_bdd_mf.py:302-331 does `.cpu()` on them) -- comes
    # back to pinned host memory every step instead of device-side label maps: 52x more D2H bytes, PCIe-bound
    e2e_logits = None
    if not args.no_logits_e2e:
        logits_h = torch.empty((B, 13, Hout, Wout), dtype=torch.float32).pin_memory()

        def logits_compute(b):
            main_stream.wait_event(b["in_ready"])
            with torch.no_grad():
                logits, _, _ = net(b["rgb"], b["ir"])
            b["consumed"].record(main_stream)
            b["out_ready"].record(main_stream)
            logits.record_stream(copy_stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(b["out_ready"])
                logits_h.copy_(logits, non_blocking=True)

        def logits_run(nsteps):
            for b in bufs:
                b["consumed"].record(main_stream)
            stage_in(bufs[0])
            for i in range(nsteps):
                cur = bufs[i & 1]
                if i + 1 < nsteps:
                    stage_in(bufs[(i + 1) & 1])
                logits_compute(cur)
            main_stream.wait_stream(copy_stream)

        logits_run(2)
        n_l = max(3, min(args.steps, 8))
        barrier()
        ev0.record()
        logits_run(n_l)
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_logits = {"value": world * B * n_l / (t.item() / 1000.0), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": logits_h.numel() * 4,
                      "steps": n_l, "what": "every step: pinned host FP32 frames -> H2D -> PSPNet forward -> D2H of the module's FP32 logits "
                      "(1.05 GB per step, PCIe-bound); copies double-buffered on a second stream"}
        del logits_h

    # ---- roofline of the dominant kernel: events around every conv launch, separate instrumented pass
    roofline = None
    if rank == 0:
        peaks, peak_src = load_peaks()
        E.conv_timer, E.op_timer = [], []
        n_prof = min(args.steps, 3)
        for _ in range(n_prof):
            step()
        torch.cuda.synchronize()
        recs, oprecs = E.conv_timer, E.op_timer
        E.conv_timer = E.op_timer = None
        other = {}
        for (name, a, b) in oprecs:
            other[name] = other.get(name, 0.0) + a.elapsed_time(b) / n_prof
        conv_ms = sum(a.elapsed_time(b) for (_, a, b) in recs) / n_prof
        n_conv = len(recs) // n_prof
        total_flops = flops_img * B
        achieved = total_flops / (conv_ms / 1000.0) / 1e12
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])       # kernel timed inside a long step
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_roofline_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            wl = tj.get("workload", {})
            if (wl.get("batch"), wl.get("height"), wl.get("width"), wl.get("precision")) == (B, H, W, args.precision):
                traffic = tj["conv_dram_bytes_per_step"]        # ncu dram__bytes_read+write summed over one step's conv launches
        roofline = {"kernel": "conv_tc_kernel / conv_halo_kernel (tcgen05 implicit GEMM; %d launches/step incl. im2col gathers)" % n_conv,
                    "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "frac_of_burst_peak": achieved / peaks["bf16_tflops"], "peak_source": peak_src + " (sustained cuBLAS bf16)",
                    "traffic": traffic, "traffic_unit": "bytes of DRAM traffic per step, all conv launches (ncu, profiles/r1_conv_step_metrics.csv)",
                    "conv_ms_per_step": conv_ms, "step_ms": ms_max / args.steps,
                    "conv_share_of_step": conv_ms / (ms_max / args.steps), "algorithmic_gflop_per_image": flops_img / 1e9,
                    "other_ops_ms_per_step": {k: round(v, 3) for k, v in sorted(other.items(), key=lambda kv: -kv[1])},
                    "note": "achieved = the reference's algorithmic conv FLOPs / conv kernel time; on this inference path the PSP head "
                            "runs the 10240->1024 bottleneck as 2048->1024 + 4 tiny GEMMs (1x1 conv commutes with the upsample), so "
                            "executed FLOPs are ~16 % lower than algorithmic"}
        if args.layer_table:
            per = {}
            for (desc, a, b) in recs:
                per.setdefault(desc, []).append(a.elapsed_time(b))
            rows = [{"layer": k, "launches_per_step": len(v) // n_prof, "ms_per_step": sum(v) / n_prof} for k, v in per.items()]
            json.dump(rows, open(args.layer_table, "w"), indent=1)

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            keep = {}
            v, cores, kind = cpu_forward_images_per_sec(H, W, images=2, threads=os.cpu_count(), keep=keep)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"2 frames of {H}x{W} through " + ("the reference's own PSPNet module (baseline/_ref, unmodified)" if kind == "reference"
                                                                 else "oracle.pspnet_forward (restatement)") + ", torch CPU FP32, one frame per call"}
            # parity of THIS build on the frame the oracle just computed (BASELINE metric: "argmax agree"): same weights, same input
            net2 = pspnet.PSPNet(sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024, backend='resnet50', in_channels=4,
                                 pretrained=False, late_fusion=True)
            net2.load_state_dict(keep["sd"])
            net2 = net2.to(dev).eval().set_precision(args.precision)
            with torch.no_grad():
                got = net2(keep["rgb"].to(dev), keep["ir"].to(dev))[0].cpu()
            ref = keep["logits"]
            parity = {"vs": "the CPU baseline's own logits (FP32) on the same weights and frame", "frames": f"1 x {H}x{W}",
                      "logits_rel_err": float((got - ref).abs().max() / ref.abs().max()),
                      "argmax_agreement": float((got.argmax(1) == ref.argmax(1)).float().mean()),
                      "tolerance": 2e-2 if args.precision == "bf16" else 1e-4}
            del net2, got, keep
        except Exception as e:                                   # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}

    # ---- the other BASELINE configs at the same N: training steps (with the gradient all-reduce) and iou_eval
    del net, rgb, ir, bufs
    torch.cuda.empty_cache()
    want = {"train_seg", "train_critic", "iou_eval", "torch"} if args.legs == "all" else set(x for x in args.legs.split(",") if x and x != "none")
    legs = {}
    leg_steps = max(3, min(args.steps, 10))
    for wl in ("train_seg", "train_critic"):
        if wl in want:
            try:
                legs[wl] = measure_train(ctx, args, wl, args.train_pairs, 320, 640, leg_steps, 3)
            except Exception as e:                               # a leg never takes the headline down with it (the same code runs
                legs[wl] = {"failed": repr(e)[:400]}             # on every rank, so a failure is symmetric and nobody is left waiting)
                from heatnet_pub_b200 import engine as _E2
                _E2.grad_arena = None
                _E2.bn_groups = 1
                torch.cuda.empty_cache()
    if "iou_eval" in want:
        try:
            legs["iou_eval"] = measure_iou_eval(ctx, args, 500, leg_steps, 3)
        except Exception as e:
            legs["iou_eval"] = {"failed": repr(e)[:400]}
    torch_gpu = None
    if "torch" in want and rank == 0 and world == 1:
        try:
            torch_gpu = measure_torch_gpu_baseline(ctx, args, B, H, W, value, legs)
        except Exception as e:
            torch_gpu = {"failed": repr(e)[:400]}

    if rank == 0:
        cfg = {"workload": f"PSPNet-ResNet50 RGB+thermal late-fusion forward (eval), batch {B} per GPU, {H}x{W} frames -> "
                           f"{Hout}x{Wout} logits, random-init weights",
               "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"image-sharded x{world}, no collective",
               "cuda_graph": bool(use_graph),
               "l2": "inputs and activations (GBs per step) far exceed the 126 MB L2; no explicit flush"}
        ts, tc, ie = legs.get("train_seg"), legs.get("train_critic"), legs.get("iou_eval")
        if ts and "images_per_s" in ts:
            cfg.update(train_seg_images_per_s=ts["images_per_s"], train_seg_ms_per_step=ts["ms_per_step"],
                       train_seg_frac_of_sustained_bf16_peak=ts["frac_of_sustained_peak"], train_seg_pairs_per_gpu=ts["per_gpu_pairs"],
                       train_seg_cuda_graph=ts["cuda_graph"],
                       allreduce_exposed_ms=(ts["allreduce"]["exposed_ms"] if ts["allreduce"] else 0.0),
                       allreduce_mbytes_per_step=(ts["allreduce"]["mbytes_per_step"] if ts["allreduce"] else 0.0),
                       allreduce_ranks=world)
        if tc and "images_per_s" in tc:
            cfg.update(train_critic_images_per_s=tc["images_per_s"], train_critic_ms_per_step=tc["ms_per_step"],
                       train_critic_allreduce_exposed_ms=(tc["allreduce"]["exposed_ms"] if tc["allreduce"] else 0.0))
        if ts and tc and "images_per_s" in ts and "images_per_s" in tc:
            # cm/train_trgb_segnet_conf.py:157-158,577-592: 500 critic iterations, then 50 seg iterations, repeated
            pairs = ts["global_pairs"]
            cfg["train_mix_500_critic_50_seg_images_per_s"] = 2 * pairs * 550 / ((500 * tc["ms_per_step"] + 50 * ts["ms_per_step"]) / 1e3)
        if ie and "maps_per_s" in ie:
            cfg.update(iou_eval_label_maps_per_s=ie["maps_per_s"], iou_eval_hbm_frac=ie["roofline"]["frac"])
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic",
                "config": cfg,
                "e2e_raw_frames": e2e_raw, "e2e_logits_d2h": e2e_logits,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "what": "every step: pinned host FP32 frames -> H2D -> PSPNet forward (public module API) -> device argmax -> D2H uint8 label maps; "
                                "copies run on a second stream and overlap the neighbouring steps' compute (double-buffered)"},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "gpu_torch_baseline": torch_gpu,
                "parity": parity, "legs": legs}
        emit(line)
    ctx.close()


if __name__ == "__main__":
    main()
