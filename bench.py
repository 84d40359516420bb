#!/usr/bin/env python
"""Benchmark of the hot path: one full G+D training step (model_wrapper.py:136-190 of the reference) at 256x256,
20 images per GPU, channel_factor 1 -- BASELINE.json configs[2]/[3].

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-graph]

N > 1 is launched by torchrun (one rank per GPU, NCCL); the JSON line is printed by rank 0.
  value      whole-job images/s, inputs resident in HBM, K steps timed with CUDA events (max over ranks)
  e2e        the same through ModelWrapper with host (pinned) inputs: H2D of images/labels/masks and D2H of the five
             losses inside the timed region
  roofline   tensor-core convolution kernel: algorithmic FLOPs of every launch of one step / their CUDA-event time
  cpu_baseline  oracle/spyramid_oracle.py (CPU restatement of the reference step) on the host cores, rank 0, N=1
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
FLOPS_PER_IMAGE = 409.8e9  # algorithmically necessary conv+linear+attention FLOPs per image per step
BATCH = 20
LR = 1e-5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--channel-factor", type=float, default=1.0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--dump-profile", default=None, help="write the per-launch tensor-kernel timings of one step here")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(source="measured", tflops=p["bf16_tflops_sustained"], tflops_burst=p["bf16_tflops"], hbm=p["hbm_gbs"])
    return dict(source="fallback", tflops=1400.0, tflops_burst=1590.0, hbm=6650.0)


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
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
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
        time.sleep(0.15)
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
        # idle samples (before the first kernel) would drag the median down: keep the upper half
        sm_sorted = sorted(sm)
        busy = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_step_time(batch, steps, warmup, threads):
    """Times oracle.train_step (the CPU restatement of the reference's step) -- checker code, used here only as the
    reported baseline."""
    from oracle import spyramid_oracle as O
    torch.set_num_threads(threads)
    g_sd, d_sd, v_sd = O.init_generator_state(1, seed=0), O.init_discriminator_state(1, seed=1), O.init_vgg_state(seed=2)
    images, labels, masks, z_d, z_g = O.synthetic_batch(batch, seed=0, mask_mode="inference")
    g_opt, d_opt = {}, {}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(g_sd, d_sd, v_sd, images, labels, masks, z_d, z_g, g_opt, d_opt, lr=LR)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times)


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    b = args.cpu_batch
    sec = cpu_reference_step_time(b, args.steps, max(1, min(args.warmup, 1)), threads)
    value = b / sec
    sample = "oracle.train_step (CPU FP32 restatement of model_wrapper.py:136-190), batch %d per step, %d timed steps, " \
             "%d threads" % (b, args.steps, threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "full G+D training step, channel_factor=1, 256x256, CPU sample batch %d" % b,
                       "batch_per_step": b},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
    sampler = ClockSampler(local_rank)
    sampler.start()
    reducer.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    reducer.barrier()
    ms_total = reducer.max_over_ranks(e0.elapsed_time(e1), device)
    clocks = sampler.stop()
    ms_per_step = ms_total / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the five losses, every step ----
    loss_host = torch.empty(len(METRICS), dtype=torch.float32).pin_memory()
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

    # ---- roofline of the dominant kernel: every tensor-core conv launch of one eager step, CUDA-event timed ----
    pk = peaks()
    ops.PROFILE = []
    eager_step()
    torch.cuda.synchronize()
    prof = ops.PROFILE
    ops.PROFILE = None
    agg = {}
    if args.dump_profile and rank == 0:
        rows = {}
        for kernel, flops, nbytes, a, b, label in prof:
            r = rows.setdefault((kernel, label), [0, 0.0, flops, nbytes])
            r[0] += 1
            r[1] += a.elapsed_time(b)
        with open(args.dump_profile, "w") as f:
            for (kernel, label), (n, ms, flops, nbytes) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
                f.write("%-18s %-46s n=%3d total %8.3f ms  avg %7.1f us  %7.1f TFLOP/s  %6.0f GB/s(alg)\n" % (
                    kernel, label, n, ms, ms / n * 1e3, flops * n / (ms * 1e-3) / 1e12, nbytes * n / (ms * 1e-3) / 1e9))
    for kernel, flops, nbytes, a, b, _label in prof:
        rec = agg.setdefault(kernel, [0.0, 0.0, 0.0, 0])
        rec[0] += flops
        rec[1] += nbytes
        rec[2] += a.elapsed_time(b) * 1e-3
        rec[3] += 1
    kernels = {}
    for kernel, (flops, nbytes, sec, n) in agg.items():
        kernels[kernel] = {"launches": n, "tflops": flops / sec / 1e12, "ms_per_step": sec * 1e3,
                           "frac_of_peak": flops / sec / 1e12 / pk["tflops"]}
    dom = max(agg.items(), key=lambda kv: kv[1][2])[0] if agg else None
    roofline = None
    if dom is not None:
        flops, nbytes, sec, n = agg[dom]
        roofline = {"bound": "tensor", "kernel": dom, "achieved": flops / sec / 1e12, "peak": pk["tflops"],
                    "unit": "TFLOP/s", "frac": flops / sec / 1e12 / pk["tflops"], "traffic": None,
                    "peak_source": "%s bf16 sustained (MEASURED_PEAKS.json)" % pk["source"],
                    "launches_per_step": n, "kernel_ms_per_step": sec * 1e3,
                    "how": "sum of algorithmic FLOPs (2*pixels*Cout*Cin*taps) of all launches in one eager step / sum of "
                           "their CUDA-event durations on the launching stream", "all_tensor_kernels": kernels}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sec = cpu_reference_step_time(args.cpu_batch, 2, 1, threads)
        cpu_baseline = {"value": args.cpu_batch / sec, "unit": "images/s", "cores": threads, "kind": "port",
                        "sample": "oracle.train_step (CPU FP32 restatement of the reference step), batch %d, 2 timed steps "
                                  "after 1 warm-up" % args.cpu_batch}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "full G+D training step (LSGAN + semantic reconstruction + diversity losses, two "
                                       "Adam updates), channel_factor=%g, batch %d/GPU, 256x256" % (cf, B),
                           "global_batch": B * world, "parallelism": "dp%d" % world, "execution": graph_note,
                           "l2_policy": "per-step working set (>2 GB of activations) exceeds the 126 MB L2"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": 4 * len(METRICS)},
                "gpu_launches": int(launches_per_step) * args.steps,
                "launches_per_step": int(launches_per_step),
                "model_tflops": FLOPS_PER_IMAGE * B * world / (ms_per_step * 1e-3) / 1e12 if cf == 1.0 else None,
                "roofline": roofline, "cpu_baseline": cpu_baseline,
                "losses": dict(zip(METRICS, loss_values))}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
