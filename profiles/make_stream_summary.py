#!/usr/bin/env python
"""profiles/r02_stream.md from the ncu exports of tools/ncu_stream.sh (gpurun_out/r2_ncu_stream_raw.csv and the three
source-page CSVs) plus the pipe micro-benchmarks (gpurun_out/r2_pipes.txt).  usage: python profiles/make_stream_summary.py"""
import collections
import csv
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
B_CAPTURE = 64


def fnum(v):
    return float(v.replace(",", ""))


rows = list(csv.reader(open(os.path.join(G, "r2_ncu_stream_raw.csv"))))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time ms", 1.0), ("launch__registers_per_thread", "regs", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1.0),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %", 1.0),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu %", 1.0),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %", 1.0),
        ("dram__bytes_read.sum", "dram rd GB", 1.0), ("dram__bytes_write.sum", "dram wr GB", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %", 1.0),
        ("smsp__inst_executed.sum", "warp instr (M)", 1e-6)]
elements = B_CAPTURE * 8 * 784 * 784
out = ["# r02: the streamed Re-Attention kernels under `ncu --set full --clock-control none` (B200)\n",
       f"Command: `sh tools/ncu_stream.sh {B_CAPTURE}` under gpurun: one Base-level-2-shaped block (784 tokens, 8 heads of 24), forward +",
       f"backward at {B_CAPTURE} images = {elements / 1e6:.0f} M attention-map elements per kernel.  Replayed, cold-cache launches: read the",
       "counters and shares, not the absolute times (CUDA-event times at 256 images are in DESIGN.md section 5).\n",
       "| kernel | " + " | ".join(n for _, n, _ in want) + " | thread instr / map element |", "|---|" + "---:|" * (len(want) + 1)]
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("void ", "").split("(")[0]
    cells = []
    for k, n, sc in want:
        v = fnum(r[col[k]]) * sc
        u = units[col[k]]
        if "GB" in n:
            v = fnum(r[col[k]]) * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1.0)
        if n == "time ms":
            v = fnum(r[col[k]]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        cells.append(f"{v:.2f}" if v < 100 else f"{v:.0f}")
    ipe = fnum(r[col["smsp__inst_executed.sum"]]) * 32 / elements
    out.append(f"| `{name}` | " + " | ".join(cells) + f" | {ipe:.1f} |")
out += ["", "`stream_fwd_kernel<8,24,1,1>` = train statistics (sweeps A + B, writes the centred bf16 probabilities), `<8,24,2,0>` = train apply",
        "(sweep C: mix + A.V).  Every kernel runs ONE 7-warp CTA per SM (all heads of a position live in one lane: 64 score registers +",
        "96 output accumulators or 72 reduction accumulators -> 240-255 registers per thread), i.e. 11 % occupancy, 1.75 warps per",
        "scheduler: no latency hiding between warps.  The issue slots are 15-42 % busy although the tensor pipe (10-16 %), the FMA",
        "pipe (13-18 %) and the XU pipe (MUFU.EX2, 14-22 %) are far from saturated: the kernels are LATENCY-bound, not throughput-bound.\n",
        "## Where the warps wait (ncu source page, warp-stall samples)\n"]


def stalls(path, title):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {k: i for i, k in enumerate(hdr)}
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot, byop, n = collections.Counter(), collections.Counter(), 0
    seen = set()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[0] in seen:
            continue
        seen.add(r[0])                      # the export lists every instruction twice (SASS + source view)
        try:
            s = int(r[col["# Samples"]] or 0)
        except ValueError:
            continue
        n += s
        for st in names:
            tot[st] += int(r[col[st]] or 0)
        toks = r[col["Source"]].split()
        op = (toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")).split(".")[0]
        byop[op] += s
    line = ", ".join(f"{k.replace('stall_', '')} {100.0 * v / max(n, 1):.0f} %" for k, v in tot.most_common(6))
    ops = ", ".join(f"{k} {100.0 * v / max(n, 1):.0f} %" for k, v in byop.most_common(8))
    return [f"* **{title}** ({n} samples): {line}.  Samples by opcode: {ops}."]


for k, t in (("stream_fwd_kernel", "forward (first captured launch: statistics)"), ("stream_bwd_reduce_kernel", "backward reductions"),
             ("stream_bwd_ds_kernel", "backward dS + dq")):
    p = os.path.join(G, f"r2_ncu_src_{k}.csv")
    if os.path.exists(p):
        out += stalls(p, t)
out += ["", "Reading: in the forward kernels the top stall is `short_sb` on the `HMMA` that consumes a just-issued `LDSM` (ldmatrix ->",
        "warp MMA with nobody else to run), then fixed-latency `wait`; in the dS kernel 70 % of the samples are `long_sb`: the bf16",
        "probability / dP words are loaded from global memory right where they are unpacked (`LOP3` / `SHF` on the loaded word), with no",
        "prefetch distance and no second warp to switch to.\n",
        "## Pipe micro-benchmarks (tools/ubench/pipes.cu, 16 warps per scheduler, B200 at 1965 MHz)\n", "```"]
pp = os.path.join(G, "r2_pipes.txt")
if os.path.exists(pp):
    out += [l.rstrip() for l in open(pp)]
out += ["```", "FFMA2 (`fma.rn.f32x2`) issues half as many instructions for the same FMA throughput (the FP32 pipe is the limit: 35-36 TFLOP/s",
        "either way), MUFU.EX2 runs at 1/8 of the FFMA rate, one dropout hash + 4 selects costs ~40 FFMA issue slots per quad (hence the",
        "cached keep-bits), and the legacy warp-MMA path peaks at 276 (TF32) / 552 (bf16) TFLOP/s -- enough for the K = 24 score",
        "contractions of this path, which is why these kernels use `mma.sync` and keep tcgen05 for the token GEMMs.", ""]
open(os.path.join(ROOT, "profiles", "r02_stream.md"), "w").write("\n".join(out))
print("\n".join(out[:14]))
