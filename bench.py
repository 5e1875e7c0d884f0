#!/usr/bin/env python
"""Benchmark of the ViT-UNet hot path: Base (depth_te=2, hidden 128, 8 heads, patch 32, 49 patches) denoising
training step -- zero_grad + forward + L1 loss + backward (+ gradient all-reduce for N>1) -- on synthetic
3x224x224 batches, images/s over all ranks (BASELINE.json metric, configs[2]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU PyTorch path (oracle port)

Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for how every field is produced.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOPS_PER_IMAGE_FWD_BWD = 23.26e9      # Base, 2xMAC of the reference math, fwd+bwd = 3x fwd
METRIC = "ViT-UNet Base 224^2 images/s fwd+bwd"
BASE_KW = dict(depth=2, depth_te=2, size_bottleneck=2, preprocessing="conv", im_size=224, patch_size=32,
               num_channels=3, hidden_dim=128, num_heads=8, attn_drop=0.2, proj_drop=0.2, linear_drop=0)


# BASELINE.json configs; "base_train" (configs[2]) is the headline metric, the others are extra bench modes
WORKLOADS = {
    "base_train": dict(kw={}, train=True, loss="l1", flops=23.26e9,
                       name="ViT_UNet Base denoising training step, L1 loss, 3x224x224 (BASELINE configs[2])"),
    "lite_infer": dict(kw=dict(depth_te=1, patch_size=16, hidden_dim=64, num_heads=4), train=False, loss="l1", flops=9.311e9,
                       name="ViT_UNet Lite inference (eval forward), 3x224x224 (BASELINE configs[1])"),
    "lite_train": dict(kw=dict(depth_te=1, patch_size=16, hidden_dim=64, num_heads=4), train=True, loss="l1", flops=27.9e9,
                       name="ViT_UNet Lite denoising training step, L1 loss (README usage line: run_denoising.py --model_string lite)"),
    "base_infer": dict(kw={}, train=False, loss="l1", flops=7.755e9,
                       name="ViT_UNet Base inference (eval forward), 3x224x224"),
    "large_train": dict(kw=dict(depth_te=4, size_bottleneck=4), train=True, loss="l1", flops=42.4e9,
                        name="ViT_UNet Large denoising training step, L1 loss (BASELINE configs[3]; pass --dtype bf16 for its bf16 storage / compute mode)"),
    "base1ch_dice": dict(kw=dict(num_channels=1), train=True, loss="dice", flops=5.46e9,
                         name="ViT_UNet Base 1-channel segmentation step, soft-Dice loss (BASELINE configs[4])"),
}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    hbm=float(d.get("hbm_gbs", 6650.0)), src="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


# kernel class of the live table -> regex over the ncu kernel names of profiles/r02_step_traffic.json
_NCU_CLASS = {
    # gemm_tf32_tc_kernel<BLOCK_N, STAGES, A_MN, B_MN, BF16, EPI_UN>
    "gemm_tcgen05_bf16:tokens": r"gemm_tf32_tc_kernel<(64|128), \d, \d, \d, 1, \d>$",     # bf16 mode: token GEMMs on bf16 operands
    "gemm_tcgen05_tf32:tokens": r"gemm_tf32_tc_kernel<\d+, \d, \d, \d, 0, \d>$",         # fp32-operand tcgen05 GEMMs (a few L0/L1 map products included)
    "gemm_tcgen05_tf32:map_in": r"gemm_tf32_tc_kernel<32, \d, \d, \d, 1, \d>$",          # bf16-operand (map-reading) tcgen05 GEMMs
    "gemm_tcgen05_tf32:map_out": r"gemm_tf32_tc_kernel<\d+, \d, \d, \d, 0, \d>$",
    "gemm_mma_tf32:map_out": r"scores_mma_kernel", "vu_reattn_bwd_rows": r"reattn_bwd_rows", "vu_softmax_stats": r"softmax_stats",
    "vu_reattn_mix_reduce": r"reattn_mix_reduce", "vu_reattn_mix": r"reattn_mix_mma_kernel|reattn_mix_kernel",
    "vu_reattn_stream_fwd": r"stream_fwd_kernel", "vu_reattn_stream_bwd_ds": r"stream_bwd_ds", "vu_reattn_stream_bwd_reduce": r"stream_bwd_reduce",
}


def _ncu_traffic(kernel_class, batch):
    """HBM traffic from this round's committed ncu capture of ONE benchmarked step (profiles/r02_step_traffic.json, made by
    tools/ncu_step.sh + profiles/make_step_summary.py: dram__bytes_read.sum + dram__bytes_write.sum of every kernel),
    scaled linearly from the captured batch to `batch`.  Returns (bytes per launch of the class, bytes per step of the
    class, bytes per step of ALL kernels, note) or Nones when the capture is absent."""
    import re
    path = os.path.join(ROOT, "profiles", "r02_step_traffic.json")
    if not os.path.exists(path):
        return None, None, None, None
    d = json.load(open(path))
    scale = batch / float(d["batch"])
    rx = next((v for k, v in _NCU_CLASS.items() if kernel_class.startswith(k)), re.escape(kernel_class.replace("vu_", "")))
    hit = [v for k, v in d["kernels"].items() if re.search(rx, k)]
    total = d["step_dram_bytes"] * scale
    if not hit:
        return None, None, total, "no kernel of the capture matches this class"
    cls_bytes = sum(v["dram_bytes"] for v in hit) * scale
    n = sum(v["launches"] for v in hit)
    return cls_bytes / n, cls_bytes, total, (f"ncu capture of one Base step at {d['batch']} images (profiles/r02_step.md), kernels matching /{rx}/: "
                                             f"{n} launches per step, scaled to {batch} images")


def _synthetic(B, gen_seed=0):
    """clean=rand, x=clamp(clean+0.1*randn,0,1) normalised like run_denoising.py:54."""
    g = torch.Generator().manual_seed(gen_seed)
    clean = torch.rand(B, 3, 224, 224, generator=g)
    noisy = (clean + 0.1 * torch.randn(B, 3, 224, 224, generator=g)).clamp(0, 1)
    return ((noisy - 0.456) / 0.224).contiguous(), clean.contiguous()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _dist_env(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and "RANK" not in os.environ:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    return rank, world, local


# ------------------------------------------------------------------------------------------- reference arm
def cpu_reference_step_time(B, steps, warmup, threads):
    """The reference's own CPU implementation of the path (oracle port of vit_unet/torch/model.py), train step with
    L1 loss and the preset's dropout (0.2/0.2/0), all host threads.  Returns seconds per step."""
    from oracle import vit_unet_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = O.get_vit_unet("base", variant="head")
    m.train()
    x, y = _synthetic(B)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        m.zero_grad(set_to_none=True)
        loss = torch.nn.functional.l1_loss(m(x), y)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times), float(loss.item())


def run_reference(args):
    rank, world, _ = _dist_env(args)
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B = args.cpu_batch
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    sec, _ = cpu_reference_step_time(B, steps, warmup, cores)
    ips = B / sec
    sample = f"Base train step (zero_grad+fwd+L1+bwd, dropout 0.2/0.2/0), batch {B}, fp32, {steps} timed + {warmup} warm-up"
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ViT_UNet Base denoising training step, L1 loss, 3x224x224 (BASELINE configs[2])",
                       "batch_per_step": B, "where": "host CPU, oracle port of the reference (reference model.py is "
                       "unconstructible at HEAD and has no build; SURVEY.md F2)"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print_line(json.dumps(line))


# ------------------------------------------------------------------------------------------- eager-PyTorch-on-GPU arm
def run_eager(args):
    """SURVEY.md 8(d), second baseline: "what a user of the reference gets today" -- the SAME plain-PyTorch model
    (oracle port of vit_unet/torch/model.py) moved to one B200 and run eagerly through cuBLAS / cuDNN, with TF32 matmuls
    off (stock default) and on.  Same workload, loss, dropout and timing rules as the CUDA arm; none of this repo's
    kernels are loaded.  Reported next to the CUDA arm, never as its value."""
    from oracle import vit_unet_oracle as O
    rank, world, local = _dist_env(args)
    if rank != 0:
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = WORKLOADS[args.workload]
    preset = {"base_train": "base", "lite_infer": "lite", "lite_train": "lite", "base_infer": "base", "large_train": "large", "base1ch_dice": "base"}[args.workload]
    B = args.eager_batch
    res = {}
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            m = O.get_vit_unet(preset, variant="head", **({"num_channels": 1} if args.workload == "base1ch_dice" else {}))
        m.to(dev).train(wl["train"])
        x, y = _synthetic(B)
        if args.workload == "base1ch_dice":
            x, y = (x[:, :1] * 0.224 + 0.456).contiguous(), (y[:, :1] > 0.5).float().contiguous()
        x, y = x.to(dev), y.to(dev)
        loss_fn = O.dice_loss if wl["loss"] == "dice" else torch.nn.functional.l1_loss

        def step():
            if not wl["train"]:
                with torch.no_grad():
                    return (m(x) - y).abs().mean()
            m.zero_grad(set_to_none=True)
            loss = loss_fn(m(x), y)
            loss.backward()
            return loss
        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        res["tf32" if tf32 else "fp32"] = {"images_per_s": B / (ms / 1e3), "ms_per_step": ms,
                                           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
        del m, x, y
        torch.cuda.empty_cache()
    best = max(res.values(), key=lambda r: r["images_per_s"])
    line = {"impl": "eager", "metric": METRIC if args.workload == "base_train" else wl["name"] + " images/s",
            "value": best["images_per_s"], "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": best["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32/f32", "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": B,
                       "where": "one B200, stock eager PyTorch (cuBLAS / cuDNN / ATen kernels) running the oracle port of the "
                                "reference model; value = the faster of TF32-off / TF32-on"},
            "eager": res, "gpu_launches": 0}
    print_line(json.dumps(line))


# ------------------------------------------------------------------------------------------- CUDA arm
def run_cuda(args):
    import torch.distributed as dist
    import vit_unet_b200 as vu
    from vit_unet_b200 import _lib, ops
    from vit_unet_b200.dp import DataParallel

    rank, world, local = _dist_env(args)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    vu.set_precision(args.precision)
    if args.streamed is not None:
        vu.set_streamed(bool(args.streamed), inference=bool(args.streamed))
    B = args.batch
    if args.global_batch:                   # strong scaling (fixed global batch, e.g. 1024 -> 128 per GPU at N=8)
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} must be divisible by the {world} ranks")
        B = args.global_batch // world
    torch.manual_seed(0)
    wl = WORKLOADS[args.workload]
    kw = dict(BASE_KW, **wl["kw"])
    if args.dropout is not None:           # experiments only; the headline run keeps the preset's 0.2/0.2/0
        kw.update(attn_drop=args.dropout, proj_drop=args.dropout)
    with contextlib.redirect_stdout(io.StringIO()):
        net = vu.HViT_UNet(**kw)
    net.to(dev).train(wl["train"])
    model = DataParallel(net) if (world > 1 and wl["train"]) else net
    x_h, y_h = _synthetic(B, gen_seed=rank)
    if kw["num_channels"] == 1:            # CT-like slices, binary disc masks
        x_h = (x_h[:, :1] * 0.224 + 0.456).contiguous()
        y_h = (y_h[:, :1] > 0.5).float().contiguous()
    x_pin, y_pin = x_h.pin_memory(), y_h.pin_memory()
    x_d, y_d = x_h.to(dev), y_h.to(dev)
    params = [p for p in net.parameters()]
    loss_fn = {"l1": vu.l1_loss, "dice": vu.dice_loss}[wl["loss"]]

    def step(x, y):
        if not wl["train"]:                # inference: batch-sharded forward, no collective
            with torch.no_grad():
                return (model(x) - y).abs().mean()
        for p in params:
            p.grad = None
        loss = loss_fn(model(x), y)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, args.min_warmup)):
        step(x_d, y_d)
    barrier()

    # ---- timed region 1: device-resident inputs (value) ------------------------------------------------
    prof = ops.KernelTimer() if args.kernel_timing else None
    _lib.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        ops.set_kernel_timer(prof)
        ev0.record()
        for _ in range(args.steps):
            step(x_d, y_d)
        ev1.record()
        ops.set_kernel_timer(None)
        barrier()
    launches = _lib.launch_count()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_per_step = ms / args.steps
    value = B * world * args.steps / (ms / 1e3)

    # ---- timed region 2: end to end through the public API with pinned HOST buffers (e2e) -----------------
    step(x_pin.to(dev, non_blocking=True), y_pin.to(dev, non_blocking=True)).item()      # one untimed e2e warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # input pipeline of the e2e loop: every step's inputs are copied from pinned host memory inside the timed region;
    # the copy of step i+1 runs on a copy stream (double-buffered device inputs) while step i computes
    copy_s = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(x_d), torch.empty_like(y_d)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    for ev in free:
        ev.record()

    def issue(i):
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(free[i % 2])          # the step that last read this buffer pair has finished
            bufs[i % 2][0].copy_(x_pin, non_blocking=True)
            bufs[i % 2][1].copy_(y_pin, non_blocking=True)
            ready[i % 2].record(copy_s)

    e0.record()
    issue(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            issue(i + 1)
        torch.cuda.current_stream().wait_event(ready[i % 2])
        loss = step(*bufs[i % 2])
        free[i % 2].record()
        loss_host = loss.item()                    # device -> host read of the step's result
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item()
    e2e = B * world * args.steps / (ms_e2e / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = _peaks()
    roof = {"bound": "tensor", "achieved": None, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": None,
            "traffic": None, "peak_source": peaks["src"]}
    if prof is not None:
        k = prof.summary(peaks["tflops"], peaks["hbm"])
        by = k["by_kernel"]
        # the dominant kernel = the class with the largest device time inside the timed region; its bound is the
        # roofline it sits closest to (algorithmic flops vs algorithmic HBM bytes of the op)
        top_name = max(by, key=lambda n: by[n]["ms"])
        top = by[top_name]
        # token GEMMs (proj / FF / dgrad / wgrad) are tensor-core work: judged against the tensor peak even when the
        # thin ones among them are epilogue / HBM limited
        tensor_bound = top["tensor_frac"] >= top["hbm_frac"] or top_name.startswith("gemm_simt") or ":tokens" in top_name
        roof.update({"kernel": top_name, "bound": "tensor" if tensor_bound else "hbm",
                     "achieved": top["tflops"] if tensor_bound else top["gbs"],
                     "peak": peaks["tflops"] if tensor_bound else peaks["hbm"],
                     "unit": "TFLOP/s" if tensor_bound else "GB/s",
                     "frac": top["tensor_frac"] if tensor_bound else top["hbm_frac"],
                     "kernel_ms_per_step": top["ms"] / args.steps, "kernel_share_of_step": top["ms"] / ms,
                     "launches_timed": top["launches"], "all_kernels_ms_per_step": k["total_kernel_ms"] / args.steps,
                     "by_kernel": by})
        if args.workload == "base_train":
            tr, cls_b, step_b, note = _ncu_traffic(top_name, B)
            if tr is not None:
                roof["traffic"] = tr                      # dram bytes per launch of the dominant kernel class (average)
                roof["traffic_per_step"] = cls_b
                roof["traffic_note"] = note
            if step_b is not None:
                alg = sum(v.get("alg_bytes", 0.0) for v in by.values()) / args.steps
                roof["step_dram_bytes"] = step_b          # every kernel of the step, from the same capture
                roof["step_algorithmic_bytes"] = alg      # sum of the ops' algorithmic bytes (what the maths must move)
                roof["step_compulsory_bytes"] = 87e6 * B  # ~87 MB per image if no attention map ever touched HBM
                roof["wasted_traffic_ratio"] = step_b / (87e6 * B)
    roof["step_tflops"] = value / world * wl["flops"] / 1e12
    roof["step_tensor_frac"] = roof["step_tflops"] / peaks["tflops"]

    line = {"metric": METRIC if args.workload == "base_train" else wl["name"] + " images/s", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, args.min_warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if args.global_batch else "weak",
            "vs_baseline": None, "dtype": {"tf32": "tf32", "bf16": "bf16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": wl["name"],
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "step": ("zero_grad + forward + loss + backward" + (" + bucketed NCCL grad all-reduce" if world > 1 else ""))
                               if wl["train"] else "eval forward under no_grad (batch-sharded, no collective)",
                       "dropout": "attn 0.2 / proj 0.2 / linear 0 (preset)" if args.dropout is None else f"OVERRIDDEN to {args.dropout}", "precision": args.precision,
                       "maps": ("P centred bf16 where N >= 256 (else fp32); mixed map A and gradient map dA/dS bf16 where N % 8 == 0; 8-head map kernels on TF32 warp MMAs" if (args.precision in ("tf32", "bf16") and os.environ.get("VU_BF16_MAPS", "1") == "1") else "fp32"),
                       "reattention": ("streamed (no attention maps)" if (args.streamed == 1 or (args.streamed is None and not wl["train"]))
                                       else "materialised maps") + " at the levels vu_reattn_stream_supported covers",
                       "l2": "per-step working set (saved activations + attention maps, GBs) >> 126 MB L2"},
            "clocks": clocks.summary(), "roofline": roof,
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": int(x_pin.nbytes + y_pin.nbytes),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps, "last_loss": loss_host,
                    "input_pipeline": "pinned host -> device every step, double-buffered on a copy stream (copy of step i+1 overlaps step i)"},
            "gpu_launches": launches,
            "gemm_tf32_fallbacks": ops.gemm_tf32_fallbacks()}      # TF32 requests that ran on the CUDA-core GEMM (0 = none)
    if world == 1 and not args.no_cpu_baseline and args.workload == "base_train":
        cores = os.cpu_count() or 1
        sec, _ = cpu_reference_step_time(args.cpu_batch, 2, 1, cores)
        line["cpu_baseline"] = {"value": args.cpu_batch / sec, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"oracle port of the reference, Base train step, batch {args.cpu_batch}, "
                                          f"fp32, 2 timed + 1 warm-up steps, {cores} threads"}
    print_line(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


print_line = print


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference", "eager"],
                    help="cuda: this repo; reference: the reference's CPU path (oracle port); eager: the same plain-PyTorch "
                         "model run eagerly on one B200 through cuBLAS/cuDNN (SURVEY 8(d) second baseline)")
    ap.add_argument("--eager-batch", type=int, default=64, help="batch of the eager-PyTorch arm (it materialises every "
                    "(B,h,N,N) map in fp32 and keeps them for autograd: ~1 GB per image for Base)")
    ap.add_argument("--workload", default="base_train", choices=["base_train", "lite_infer", "lite_train", "base_infer", "large_train", "base1ch_dice"],
                    help="base_train is the headline (BASELINE.json metric); the others are extra modes")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step (SURVEY 8(d) C3: 64..256; ~40 GB of the 180 GB at 256)")
    ap.add_argument("--global-batch", type=int, default=0, help="strong scaling: fixed GLOBAL batch split over the ranks "
                    "(overrides --batch; reported with scaling=strong)")
    ap.add_argument("--cpu-batch", type=int, default=8, help="batch of the bounded CPU sample")
    ap.add_argument("--precision", "--dtype", dest="precision", default=os.environ.get("VU_PRECISION", "bf16"),
                    choices=["fp32", "tf32", "bf16"],
                    help="bf16 (default): bf16 storage of every GEMM operand / saved activation outside the residual stream + bf16 "
                         "tcgen05 token GEMMs, fp32 residual stream / statistics / master weights / gradients; tf32: tcgen05 "
                         "contractions on fp32 storage; fp32: CUDA-core exact mode")
    ap.add_argument("--streamed", type=int, default=None, help="1/0: force the streamed Re-Attention kernels on/off "
                    "(default: on for no-grad inference, off for training steps; see DESIGN.md)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dropout", type=float, default=None, help="override attn/proj dropout (experiments; default = preset 0.2)")
    ap.add_argument("--min-warmup", type=int, default=3, help="lower only for profiler runs (numbers under ncu are never bench values)")
    ap.add_argument("--kernel-timing", type=int, default=1, help="CUDA-event timing of every GEMM launch (roofline)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 on their own (NCCL prints its version banner
    # there) are diverted to stderr for the duration of the run; the JSON line goes to the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real_stdout, "w")
    global print_line
    print_line = lambda text: (out.write(text + "\n"), out.flush())
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "eager":
        run_eager(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
