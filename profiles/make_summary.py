#!/usr/bin/env python
"""Turn the raw ncu exports brought back in gpurun_out/ into the committed summaries under profiles/.

  python profiles/make_summary.py r01        # reads gpurun_out/ncu_block_raw.csv, gpurun_out/launches_tf32.csv
"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
B_CAPTURE = int(sys.argv[2]) if len(sys.argv) > 2 else 32


def to_bytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    return float(v.replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


def block_table():
    path = os.path.join(ROOT, "gpurun_out", "ncu_block_raw.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in hdr}
    want = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "dram rd MB"),
            ("dram__bytes_write.sum", "dram wr MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
            ("launch__registers_per_thread", "regs"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %")]
    out = [f"# {tag}: `ncu --set full --clock-control none` of ONE Base-L2-shaped block (N=784 tokens, D=192, 8 heads, hd=24), "
           f"forward+backward, {B_CAPTURE} images, TF32 path\n",
           "Command: `sh tools/ncu_block.sh 32` under gpurun (tools/profile_block.py, cudaProfilerStart around step 3).",
           "Cold-cache, serialised replays: use the SHARES and the per-kernel counters, not absolute times.\n",
           "| kernel | grid | " + " | ".join(n for _, n in want) + " |", "|---|---|" + "---:|" * len(want)]
    traffic = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void vu::", "").replace("vu::", "").replace("void ", "").replace("mma::", "")
        if name.startswith("void at::") or name.startswith("at::"):
            continue
        cells = []
        for k, n in want:
            i = col.get(k)
            if i is None:
                cells.append("-"); continue
            v, u = r[i], units[i]
            if "MB" in n:
                cells.append(f"{to_bytes(v, u) / 1e6:.0f}")
            elif "us" in n:
                cells.append(f"{to_us(v, u):.0f}")
            else:
                try: cells.append(f"{float(v):.0f}")
                except ValueError: cells.append(v)
        out.append(f"| `{name}` | {r[col['Grid Size']].replace(' ', '')} | " + " | ".join(cells) + " |")
        tot = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
            to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        base = re.sub(r"<.*", "", name).replace("_kernel", "")
        key = "vu_" + base
        if tot > traffic.get(key, {}).get("bytes_per_launch", 0):
            traffic[key] = {"bytes_per_launch": tot, "bytes_per_image": tot / B_CAPTURE, "captured_batch": B_CAPTURE,
                            "shape": "Base L2 level block (N=784, h=8, hd=24)", "grid": r[col["Grid Size"]]}
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_block.md"), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)


def launches():
    path = os.path.join(ROOT, "gpurun_out", "launches_tf32.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if l.startswith('"')]
    agg = defaultdict(lambda: [0, 0.0])
    n = 0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r.get("Metric Unit", "ns"), 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void vu::", "").replace("vu::", "").replace("void ", "").replace("mma::", "")
        agg[name][0] += 1; agg[name][1] += ns; n += 1
    tot = sum(v[1] for v in agg.values())
    out = [f"# {tag}: launch list of `bench.py --steps 1 --warmup 1 --min-warmup 1 --batch 32 --precision tf32` under "
           "`ncu --metrics gpu__time_duration.sum --clock-control none`\n",
           f"{n} launches (4 training steps: warm-up, timed, e2e warm-up, e2e), {tot / 1e6:.1f} ms of kernel time "
           "(cold-cache, serialised: compare SHARES).\n", "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% |")
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")


def gemm_summary():
    raw = os.path.join(ROOT, "gpurun_out", "ncu_gemm_raw.csv")
    shapes = os.path.join(ROOT, "gpurun_out", "gemm_shapes.txt")
    if not (os.path.exists(raw) and os.path.exists(shapes)):
        return
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    keys = ["Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"]
    out = [f"# {tag}: tcgen05 TF32 GEMM (`gemm_tf32_tc_kernel`) on the token-GEMM shapes of a Base step, 64 images\n",
           "## CUDA-event timings (`python tools/gemm_bench.py 64`, 10 launches each after 3 warm-ups)\n", "```",
           open(shapes).read().rstrip(), "```\n",
           "## `ncu --set full --clock-control none -k regex:gemm_tf32_tc -c 2` on the first shape "
           "(L0 proj forward, M=3136 N=3072 K=3072)\n", "| metric | launch 1 | launch 2 |", "|---|---:|---:|"]
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            out.append(f"| `{k}` ({units[i]}) | " + " | ".join(r[i] for r in rows[2:4]) + " |")
    out.append("\nSASS check (`cuobjdump -sass vit_unet_b200/libvitunet_b200.so | grep -c ...`): `UTCHMMA` (tcgen05.mma), "
               "`UTMALDG.4D` (TMA), `LDTM` (tcgen05.ld), `SYNCS.*TRYWAIT` (mbarrier) are all present; no `HMMA`.")
    open(os.path.join(ROOT, "profiles", f"{tag}_gemm_tcgen05.md"), "w").write("\n".join(out) + "\n")


block_table()
launches()
gemm_summary()
print("written:", sorted(os.listdir(os.path.join(ROOT, "profiles"))))
