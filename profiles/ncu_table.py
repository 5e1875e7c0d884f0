#!/usr/bin/env python
"""Key metrics per kernel launch from an ncu report: python profiles/ncu_table.py gpurun_out/prof.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
print("| " + " | ".join(n for _, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    cells = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0].replace("void vu::", "")[:48]
        elif n in ("dram_rd", "dram_wr", "time"):
            v = f"{v} {units[i]}"
        else:
            try: v = f"{float(v):.1f}"
            except ValueError: pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
