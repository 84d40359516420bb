#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json): the Semantic-Pyramid GAN training step at 256x256.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|3] [--batch B]
                  [--channel-factor F] [--precision bf16|split] [--no-graph]

  --config 3 (default)  full G+D training step (model_wrapper.py:136-190 of the reference), 20 images per GPU,
                        channel_factor 1 -- BASELINE.json configs[2] on one GPU, configs[3] under torchrun (N ranks, NCCL);
                        `--channel-factor 2 --batch 32` is configs[4]
  --config 2            VGG-16 feature pyramid + Generator forward only (model_wrapper.py:144-151), configs[1]

N > 1 is launched by torchrun (one rank per GPU, NCCL); the JSON line is printed by rank 0.
  value         whole-job images/s, inputs resident in HBM, K steps timed with CUDA events (max over ranks)
  e2e           the same through the public API with host (pinned) inputs: H2D of images/labels/masks and D2H of the
                step's result inside the timed region
  roofline      dominant tensor-core kernel: algorithmic FLOPs of every launch of one step / their CUDA-event time, each
                launch bracketed alone (phase and leaf streams OFF for that pass, so nothing shares the GPU with the
                timed kernel); keyed by the real kernel name; `traffic` = DRAM bytes per launch from the committed ncu
                capture (profiles/r02_dram_traffic.json); `hbm` = the dominant bandwidth-bound pass against the HBM peak
  cpu_baseline  oracle/spyramid_oracle.py (CPU restatement of the reference step) on the host cores, rank 0, N=1, at the
                benchmark batch (and at batch 2, the reference's own CPU case)
`--impl reference` times that CPU path alone with the same metric/config keys.
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "G+D train images/sec at 256x256, bs=20/GPU"
METRIC_FWD = "VGG-16 pyramid + Generator forward images/sec at 256x256, bs=20"
# algorithmically necessary conv+linear+attention FLOPs per image: full step / forward only, by channel factor
FLOPS_PER_IMAGE = {1.0: 409.8e9, 2.0: 197.6e9, 0.5: 1239e9}
FLOPS_PER_IMAGE_FWD = {1.0: 65.35e9, 2.0: 47.90e9, 0.5: 129.95e9}
BATCH = 20
LR = 1e-5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3], help="BASELINE.json configs[] entry (1-based)")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--channel-factor", type=float, default=1.0)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "split"],
                    help="bf16: single-plane BF16 operands (throughput mode); split: hi+lo planes, the strict parity mode")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=None, help="CPU baseline batch (default: the benchmark batch)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds the reference arm may spend on timed steps")
    ap.add_argument("--dump-profile", default=None, help="write the per-launch tensor-kernel timings of one step here")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(source="measured", tflops=p["bf16_tflops_sustained"], tflops_burst=p["bf16_tflops"], hbm=p["hbm_gbs"])
    return dict(source="fallback", tflops=1400.0, tflops_burst=1590.0, hbm=6650.0)


def dram_traffic():
    """DRAM bytes per launch of the tensor kernels, from the committed ncu capture (dram__bytes_read + _write)."""
    path = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f)
    return {}


def host_batch(batch, seed):
    """Synthetic batch in the data loader's format (data.py:76-90): images U(-1,1) FP32 NCHW, one-hot int64 labels, seven
    masks with get_masks_for_training semantics.  Returned in pinned host memory."""
    from semantic_pyramid_for_image_generation_b200 import misc
    g = torch.Generator().manual_seed(1000 + seed)
    random.seed(seed)
    np.random.seed(seed)
    images = torch.rand(batch, 3, 256, 256, generator=g) * 2 - 1
    classes = torch.randint(0, 365, (batch,), generator=g)
    labels = torch.nn.functional.one_hot(classes, 365).long()
    per_sample = [misc.get_masks_for_training() for _ in range(batch)]
    masks = [torch.stack([per_sample[b][lvl] for b in range(batch)], dim=0).contiguous() for lvl in range(7)]
    pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
    return pin(images), pin(labels), [pin(m) for m in masks]


class ClockSampler(object):
    QUERY = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # raw samples, no filtering: the sampler runs only while the timed loop does
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: oracle/spyramid_oracle.py (checker code) used only as the reported baseline
# ------------------------------------------------------------------------------------------------
def cpu_step_times(config, cf, batch, steps, warmup, threads, budget_s):
    """Seconds per step of the CPU restatement of the reference (full step, or VGG + G forward for config 2): `warmup`
    untimed steps, then up to `steps` timed ones while the time budget lasts (at least one)."""
    from oracle import spyramid_oracle as O
    torch.set_num_threads(threads)
    g_sd, d_sd, v_sd = O.init_generator_state(cf, seed=0), O.init_discriminator_state(cf, seed=1), O.init_vgg_state(seed=2)
    images, labels, masks, z_d, z_g = O.synthetic_batch(batch, seed=0, mask_mode="inference")
    g_opt, d_opt = {}, {}

    def one():
        if config == 2:
            with torch.no_grad():
                feats = O.vgg16_features(v_sd, images)
                O.generator_forward(g_sd, z_d, feats, masks, labels.float(), training=True)
        else:
            O.train_step(g_sd, d_sd, v_sd, images, labels, masks, z_d, z_g, g_opt, d_opt, lr=LR)

    times = []
    spent = 0.0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            spent += dt
            if spent + dt > budget_s:  # the next step would overrun the budget
                break
    return times


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cf = args.channel_factor
    b = args.cpu_batch or args.batch
    # one cheap warm-up at batch 2 (thread pools, allocator) when the real batch is expensive
    if b > 4:
        cpu_step_times(args.config, cf, 2, 1, 0, threads, 1e9)
        warm = 0
    else:
        warm = max(1, min(args.warmup, 1))
    times = cpu_step_times(args.config, cf, b, args.steps, warm, threads, args.cpu_budget)
    sec = sum(times) / len(times)
    value = b / sec
    what = "full G+D training step (model_wrapper.py:136-190)" if args.config == 3 else \
        "VGG-16 pyramid + Generator forward (model_wrapper.py:144-151)"
    sample = "oracle/spyramid_oracle.py (CPU FP32 restatement of the reference: %s), batch %d per step, %d timed step(s) " \
             "of the %d requested within a %.0f s budget, %d threads" % (what, b, len(times), args.steps, args.cpu_budget,
                                                                        threads)
    line = {"impl": "reference", "metric": METRIC if args.config == 3 else METRIC_FWD, "value": value, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "steps_timed": len(times), "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "%s, channel_factor=%g, batch %d, 256x256, CPU" % (what, cf, b), "batch_per_step": b},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# roofline bookkeeping shared by both configs
# ------------------------------------------------------------------------------------------------
def profile_pass(fn, ops):
    """One eager execution of `fn` with every tensor-core launch (and the dominant bandwidth-bound pass) bracketed by its
    own CUDA-event pair.  Phase streams and leaf streams are off, so the timed kernel has the GPU to itself."""
    prev = os.environ.get("SPYR_PHASE_STREAMS")
    os.environ["SPYR_PHASE_STREAMS"] = "0"
    ops.PROFILE = []
    try:
        # Park the GPU behind a ~0.25 s spin so that the whole eager pass is ENQUEUED before any of it runs: the event
        # pairs then bracket back-to-back kernels.  (Without it the GPU idles between an event record and the launch
        # that Python issues ~5-10 us later, and that idle time lands inside every bracket: +20 % on 233 launches.)
        if hasattr(torch.cuda, "_sleep"):
            torch.cuda._sleep(int(5e8))
        fn()
        torch.cuda.synchronize()
        prof = ops.PROFILE
    finally:
        ops.PROFILE = None
        if prev is None:
            os.environ.pop("SPYR_PHASE_STREAMS", None)
        else:
            os.environ["SPYR_PHASE_STREAMS"] = prev
    return prof


def roofline_from(prof, pk, dump_path=None):
    if dump_path:
        rows = {}
        for kernel, flops, nbytes, a, b, label in prof:
            r = rows.setdefault((kernel, label), [0, 0.0, flops, nbytes])
            r[0] += 1
            r[1] += a.elapsed_time(b)
        with open(dump_path, "w") as f:
            for (kernel, label), (n, ms, flops, nbytes) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
                f.write("%-20s %-46s n=%3d total %8.3f ms  avg %7.1f us  %7.1f TFLOP/s  %6.0f GB/s(alg)\n" % (
                    kernel, label, n, ms, ms / n * 1e3, flops * n / (ms * 1e-3) / 1e12, nbytes * n / (ms * 1e-3) / 1e9))
    agg = {}
    for kernel, flops, nbytes, a, b, _label in prof:
        rec = agg.setdefault(kernel, [0.0, 0.0, 0.0, 0])
        rec[0] += flops
        rec[1] += nbytes
        rec[2] += a.elapsed_time(b) * 1e-3
        rec[3] += 1
    traffic = dram_traffic()
    tensor = {k: v for k, v in agg.items() if v[0] > 0}
    hbm = {k: v for k, v in agg.items() if v[0] == 0 and v[1] > 0}
    kernels = {}
    for kernel, (flops, nbytes, sec, n) in tensor.items():
        kernels[kernel] = {"launches": n, "tflops": flops / sec / 1e12, "ms_per_step": sec * 1e3,
                           "avg_launch_us": sec / n * 1e6, "frac_of_peak": flops / sec / 1e12 / pk["tflops"],
                           "dram_bytes_per_launch": traffic.get(kernel, {}).get("dram_bytes_per_launch")}
    roofline = None
    if tensor:
        dom = max(tensor.items(), key=lambda kv: kv[1][2])[0]
        flops, nbytes, sec, n = tensor[dom]
        tot_f = sum(v[0] for v in tensor.values())
        tot_s = sum(v[2] for v in tensor.values())
        roofline = {"bound": "tensor", "kernel": dom, "achieved": flops / sec / 1e12, "peak": pk["tflops"],
                    "unit": "TFLOP/s", "frac": flops / sec / 1e12 / pk["tflops"],
                    "traffic": traffic.get(dom, {}).get("dram_bytes_per_launch"),
                    "traffic_source": traffic.get(dom, {}).get("source"),
                    "algorithmic_flops_per_launch": flops / n, "avg_launch_us": sec / n * 1e6,
                    "peak_source": "%s bf16 sustained (MEASURED_PEAKS.json)" % pk["source"],
                    "launches_per_step": n, "kernel_ms_per_step": sec * 1e3,
                    "how": "sum of algorithmic FLOPs (2*pixels*Cout*Cin*taps) of all launches of this kernel in one eager "
                           "step / sum of their CUDA-event durations on the launching stream; phase and leaf streams off "
                           "during this pass, so each event pair brackets one kernel running alone",
                    "all_tensor_kernels": kernels,
                    "all_tensor_kernels_frac": tot_f / tot_s / 1e12 / pk["tflops"]}
        if hbm:
            hk = max(hbm.items(), key=lambda kv: kv[1][2])[0]
            _, nbytes, sec, n = hbm[hk]
            roofline["hbm"] = {"bound": "hbm", "kernel": hk, "achieved": nbytes / sec / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                               "frac": nbytes / sec / 1e9 / pk["hbm"], "launches_per_step": n, "kernel_ms_per_step": sec * 1e3,
                               "how": "algorithmic bytes (input read once + outputs written once) / CUDA-event time"}
    return roofline


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); the hot path has no CPU fallback. "
                         "Use --impl reference for the CPU baseline.")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    from semantic_pyramid_for_image_generation_b200 import _native, distributed, models, ops
    from semantic_pyramid_for_image_generation_b200.model_wrapper import METRICS, ModelWrapper
    from semantic_pyramid_for_image_generation_b200.optim import FusedAdam
    ops.set_precision(args.precision)
    reducer = distributed.init_from_env("nccl") if world > 1 else distributed.GradientReducer()

    torch.manual_seed(0)  # identical replicas on every rank
    cf = args.channel_factor
    G = models.Generator(channels_factor=cf).to(device)
    D = models.Discriminator(channel_factor=cf).to(device)
    V = models.VGG16().to(device).eval()
    G.train()
    D.train()
    wrapper = ModelWrapper(G, D, None, None, vgg16=V, generator_optimizer=FusedAdam(G.parameters(), lr=LR),
                           discriminator_optimizer=FusedAdam(D.parameters(), lr=LR),
                           save_data_path=os.path.join("/tmp", "spyr_bench_%d" % os.getpid()),
                           reducer=reducer if world > 1 else None)
    torch.manual_seed(1234 + rank)  # own noise stream per rank
    B = args.batch
    h_images, h_labels, h_masks = host_batch(B, seed=rank)
    s_images = h_images.to(device)
    s_labels = h_labels.to(device)
    s_masks = [m.to(device) for m in h_masks]
    h2d_bytes = h_images.numel() * 4 + h_labels.numel() * 8 + sum(m.numel() * 4 for m in h_masks)
    pk = peaks()
    dtype = "bf16" if args.precision == "bf16" else "bf16 hi+lo planes (split), fp32 accumulate"

    if args.config == 2:
        run_forward_config(args, rank, world, local_rank, device, wrapper, G, V, reducer, s_images, s_labels, s_masks,
                           h_images, h_labels, h_masks, h2d_bytes, pk, dtype, ops, _native)
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    def eager_step():
        return wrapper.training_step(s_images, s_labels, s_masks)

    for _ in range(max(args.warmup, 3)):
        losses = eager_step()
    torch.cuda.synchronize()

    # ---- the iteration as CUDA graphs (three graphs, the two NCCL gradient averages run between them) ----
    captured = None
    graph_note = "eager"
    launches_per_step = None
    if not args.no_graph:
        try:
            captured = wrapper.capture_training_step(s_images, s_labels, s_masks)
            launches_per_step = captured.launches_per_step
            losses = captured()
            torch.cuda.synchronize()
            graph_note = "cuda-graphs (3 per step)"
        except Exception as exc:  # noqa: BLE001
            captured = None
            graph_note = "eager (graph capture failed: %s)" % str(exc)[:160]
            torch.cuda.synchronize()
    if launches_per_step is None:
        _native.launch_count_reset()
        losses = eager_step()
        torch.cuda.synchronize()
        launches_per_step = _native.launch_count()

    def step():
        if captured is not None:
            captured()
        else:
            eager_step()

    for _ in range(3):
        step()
    # ---- device-resident timing ----
    reducer.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    reducer.barrier()
    ms_total = reducer.max_over_ranks(e0.elapsed_time(e1), device)
    ms_per_step = ms_total / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the five losses, every step ----
    loss_host = torch.empty(len(METRICS), dtype=torch.float32).pin_memory()
    # untimed warm-up of the end-to-end path itself (first use of the copy stream, of the staging buffers and of the DMA
    # mappings of the pinned batch: on a cold box these one-off costs were worth 4 ms/step over a 20-step region)
    for i in range(max(args.warmup, 3)):
        if captured is not None:
            if i == 0:
                captured.load(h_images, h_labels, h_masks)
            out = captured()
            if i + 1 < max(args.warmup, 3):
                captured.prefetch(h_images, h_labels, h_masks)
        else:
            s_images.copy_(h_images, non_blocking=True)
            out = eager_step()
        loss_host.copy_(torch.stack([out[name] for name in METRICS]), non_blocking=True)
    torch.cuda.synchronize()
    reducer.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    if captured is not None:
        # every step's inputs cross PCIe from pinned memory inside the timed region; the copy of step i+1 is issued on the
        # copy stream right after step i is launched (CapturedTrainingStep.prefetch) and lands device-to-device
        captured.load(h_images, h_labels, h_masks)  # step 0: on the timed stream itself
    for i in range(args.steps):
        if captured is not None:
            out = captured()
            if i + 1 < args.steps:
                captured.prefetch(h_images, h_labels, h_masks)
        else:
            s_images.copy_(h_images, non_blocking=True)
            s_labels.copy_(h_labels, non_blocking=True)
            for dm, hm in zip(s_masks, h_masks):
                dm.copy_(hm, non_blocking=True)
            out = eager_step()
        loss_host.copy_(torch.stack([out[name] for name in METRICS]), non_blocking=True)
    t1.record()
    torch.cuda.synchronize()
    reducer.barrier()
    e2e_ms = reducer.max_over_ranks(t0.elapsed_time(t1), device) / args.steps
    e2e_value = B * world / (e2e_ms * 1e-3)
    loss_values = loss_host.tolist()

    # ---- roofline: every tensor-core launch of one eager step, each timed alone ----
    roofline = roofline_from(profile_pass(eager_step, ops), pk, args.dump_profile if rank == 0 else None)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = args.cpu_batch or B
        cpu_step_times(3, cf, 2, 1, 0, threads, 1e9)  # warm-up at batch 2
        t_small = cpu_step_times(3, cf, 2, 2, 0, threads, 20.0)
        t_full = cpu_step_times(3, cf, cb, 2, 0, threads, 25.0)
        sec = sum(t_full) / len(t_full)
        cpu_baseline = {"value": cb / sec, "unit": "images/s", "cores": threads, "kind": "port",
                        "sample": "oracle.train_step (CPU FP32 restatement of the reference step), batch %d (the benchmark "
                                  "batch), %d timed step(s) after a batch-2 warm-up" % (cb, len(t_full)),
                        "batch2": {"value": 2 / (sum(t_small) / len(t_small)), "steps": len(t_small),
                                   "note": "BASELINE.json configs[0]: the reference's own CPU case"}}
    if rank == 0:
        fpi = FLOPS_PER_IMAGE.get(cf)
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": {"workload": "full G+D training step (LSGAN + semantic reconstruction + diversity losses, two "
                                       "Adam updates), channel_factor=%g, batch %d/GPU, 256x256" % (cf, B),
                           "baseline_config": "configs[4]" if (cf == 2.0 and B == 32) else
                                              ("configs[3]" if world > 1 else "configs[2]"),
                           "global_batch": B * world, "parallelism": "dp%d" % world, "execution": graph_note,
                           "precision": args.precision, "deterministic": True,
                           "l2_policy": "per-step working set (>2 GB of activations) exceeds the 126 MB L2"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": 4 * len(METRICS)},
                "gpu_launches": int(launches_per_step) * args.steps,
                "launches_per_step": int(launches_per_step),
                "model_tflops": fpi * B * world / (ms_per_step * 1e-3) / 1e12 if fpi else None,
                "roofline": roofline, "cpu_baseline": cpu_baseline,
                "losses": dict(zip(METRICS, loss_values))}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_forward_config(args, rank, world, local_rank, device, wrapper, G, V, reducer, s_images, s_labels, s_masks,
                       h_images, h_labels, h_masks, h2d_bytes, pk, dtype, ops, _native):
    """BASELINE.json configs[1]: VGG-16 feature-pyramid extraction + Generator forward only (model_wrapper.py:144-151)."""
    B = args.batch
    cf = args.channel_factor
    labels_f = s_labels.float()
    z = torch.randn((B, wrapper.latent_dimensions), dtype=torch.float32, device=device)

    def forward():
        with torch.no_grad():
            feats = V(s_images)
            return G(input=z, features=feats, masks=s_masks, class_id=labels_f)

    for _ in range(max(args.warmup, 3)):
        img = forward()
    torch.cuda.synchronize()
    graph, graph_note = None, "eager"
    _native.launch_count_reset()
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                forward()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            _native.launch_count_reset()
            with torch.cuda.graph(graph):
                img = forward()
            graph_note = "cuda-graph (1 per step)"
        except Exception as exc:  # noqa: BLE001
            graph, graph_note = None, "eager (graph capture failed: %s)" % str(exc)[:160]
            torch.cuda.synchronize()
    if graph is None:
        _native.launch_count_reset()
        img = forward()
        torch.cuda.synchronize()
    launches_per_step = _native.launch_count()

    def step():
        if graph is not None:
            graph.replay()
        else:
            forward()

    for _ in range(3):
        step()
    reducer.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    reducer.barrier()
    ms_per_step = reducer.max_over_ranks(e0.elapsed_time(e1), device) / args.steps
    value = B * world / (ms_per_step * 1e-3)
    # end to end: images / labels / masks from pinned host memory every step, the generated batch's checksum back
    result_host = torch.empty(1, dtype=torch.float32).pin_memory()
    reducer.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # every step's inputs cross PCIe from pinned memory inside the timed region: step 0's on the timed stream, step i+1's on
    # a copy stream into staging buffers while step i computes, then device-to-device into the graph's inputs
    stage = (torch.empty_like(s_images), torch.empty_like(s_labels), [torch.empty_like(m) for m in s_masks])
    copy_stream = torch.cuda.Stream()
    staged, consumed = torch.cuda.Event(), torch.cuda.Event()
    main = torch.cuda.current_stream()

    def load(images, labels, masks):
        s_images.copy_(images, non_blocking=True)
        s_labels.copy_(labels, non_blocking=True)
        for dm, hm in zip(s_masks, masks):
            dm.copy_(hm, non_blocking=True)
        labels_f.copy_(s_labels)

    def e2e_pass(steps):
        load(h_images, h_labels, h_masks)
        consumed.record(main)
        for i in range(steps):
            if i > 0:
                main.wait_event(staged)
                load(*stage)
                consumed.record(main)
            if i + 1 < steps:
                copy_stream.wait_event(consumed)
                with torch.cuda.stream(copy_stream):
                    stage[0].copy_(h_images, non_blocking=True)
                    stage[1].copy_(h_labels, non_blocking=True)
                    for dm, hm in zip(stage[2], h_masks):
                        dm.copy_(hm, non_blocking=True)
                    staged.record(copy_stream)
            step()
            result_host.copy_(img.abs().mean().reshape(1), non_blocking=True)

    e2e_pass(max(args.warmup, 3))  # untimed warm-up of the end-to-end path (copy stream, staging, DMA mappings)
    torch.cuda.synchronize()
    t0.record()
    e2e_pass(args.steps)
    t1.record()
    torch.cuda.synchronize()
    reducer.barrier()
    e2e_ms = reducer.max_over_ranks(t0.elapsed_time(t1), device) / args.steps
    roofline = roofline_from(profile_pass(forward, ops), pk, args.dump_profile if rank == 0 else None)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = args.cpu_batch or B
        cpu_step_times(2, cf, 2, 1, 0, threads, 1e9)
        t_full = cpu_step_times(2, cf, cb, 3, 0, threads, 25.0)
        sec = sum(t_full) / len(t_full)
        cpu_baseline = {"value": cb / sec, "unit": "images/s", "cores": threads, "kind": "port",
                        "sample": "oracle VGG-16 pyramid + generator forward (CPU FP32 restatement of the reference), batch "
                                  "%d, %d timed pass(es) after a batch-2 warm-up" % (cb, len(t_full))}
    if rank == 0:
        fpi = FLOPS_PER_IMAGE_FWD.get(cf)
        line = {"metric": METRIC_FWD, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": {"workload": "VGG-16 feature pyramid + Generator forward only (model_wrapper.py:144-151), "
                                       "channel_factor=%g, batch %d/GPU, 256x256" % (cf, B),
                           "baseline_config": "configs[1]", "global_batch": B * world, "parallelism": "dp%d" % world,
                           "execution": graph_note, "precision": args.precision, "deterministic": True,
                           "l2_policy": "per-step working set (>1 GB of activations) exceeds the 126 MB L2"},
                "clocks": clocks,
                "e2e": {"value": B * world / (e2e_ms * 1e-3), "unit": "images/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches_per_step) * args.steps, "launches_per_step": int(launches_per_step),
                "model_tflops": fpi * B * world / (ms_per_step * 1e-3) / 1e12 if fpi else None,
                "roofline": roofline, "cpu_baseline": cpu_baseline,
                "result_checksum": float(result_host[0])}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
