#!/usr/bin/env python
"""profiles/r02_step.md + profiles/ncu_traffic.json from gpurun_out/r2_step_kernels.csv (tools/ncu_step.sh): every kernel of
one benchmarked Base training step with duration and dram__bytes_read/write -> per-kernel-class time share and HBM traffic."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
MODE = sys.argv[2] if len(sys.argv) > 2 else "bf16"
path = os.path.join(ROOT, "gpurun_out", "r2_step_kernels.csv")
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
col = {k: i for i, k in enumerate(hdr)}


def unit(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)


per = collections.defaultdict(lambda: {"ms": 0.0, "rd": 0.0, "wr": 0.0, "n": 0})
ids = collections.defaultdict(dict)
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"^void ", "", r[col["Kernel Name"]]).split("(")[0].replace("vu::", "")
    ids[r[col["ID"]]][r[col["Metric Name"]]] = (name, unit(r[col["Metric Value"]], r[col["Metric Unit"]]))
for _, m in ids.items():
    name = next(iter(m.values()))[0]
    e = per[name]
    e["n"] += 1
    e["ms"] += m.get("gpu__time_duration.sum", (name, 0.0))[1]
    e["rd"] += m.get("dram__bytes_read.sum", (name, 0.0))[1]
    e["wr"] += m.get("dram__bytes_write.sum", (name, 0.0))[1]
tot_ms = sum(e["ms"] for e in per.values())
tot_b = sum(e["rd"] + e["wr"] for e in per.values())
out = [f"# r02: every kernel of one benchmarked Base training step ({B} images, {MODE} mode = the bench default, HEAD) under ncu\n",
       f"Command: `sh tools/ncu_step.sh {B} {MODE}` under gpurun (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum",
       "--clock-control none --profile-from-start off`, tools/profile_step.py).  Serialised, cold-cache launches: compare SHARES.\n",
       f"Step total under ncu: {tot_ms:.1f} ms in {sum(e['n'] for e in per.values())} launches; DRAM traffic {tot_b / 1e9:.1f} GB per step "
       f"= {tot_b / B / 1e6:.0f} MB per image.\n",
       "| kernel | launches | ms | share | DRAM read GB | DRAM write GB | GB/s |", "|---|---:|---:|---:|---:|---:|---:|"]
for name, e in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
    out.append(f"| `{name}` | {e['n']} | {e['ms']:.2f} | {100 * e['ms'] / tot_ms:.1f} % | {e['rd'] / 1e9:.2f} | {e['wr'] / 1e9:.2f} | "
               f"{(e['rd'] + e['wr']) / max(e['ms'], 1e-9) / 1e6:.0f} |")
open(os.path.join(ROOT, "profiles", "r02_step.md"), "w").write("\n".join(out) + "\n")
json.dump({"batch": B, "step_ms_under_ncu": tot_ms, "step_dram_bytes": tot_b,
           "kernels": {k: {"launches": e["n"], "ms": e["ms"], "dram_bytes": e["rd"] + e["wr"]} for k, e in per.items()}},
          open(os.path.join(ROOT, "profiles", "r02_step_traffic.json"), "w"), indent=1)
print("\n".join(out[:30]))
