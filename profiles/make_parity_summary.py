#!/usr/bin/env python
"""profiles/r02_parity.md from the output of tools/parity_table.py on the GPU box (gpurun_out/r2_parity_table.txt)."""
import os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "r2_parity_table.txt")
out = ["# r02: measured distance of every path from the reference's golden vectors (B200)\n",
       "`python tools/parity_table.py` under gpurun: max-norm error relative to max|reference| (the reference evaluated in fp64 where",
       "stored), worst tensor per group.  Groups: eval_out; evg / trn = eval- / train-mode with the L1 loss (outputs, loss, dx, per-parameter",
       "gradients g); mev / mtr = the same with the MSE loss the reference trains with; buf = BatchNorm buffers after the train step.",
       "`chaotic` = tensors whose fp32 reference is itself > 1e-3 away from fp64 (not comparable by any fp32 implementation).",
       "L1 gradients of a tensor-core path are NOT a parity criterion (sign(out - y) flips where a 1e-3-accurate output crosses the target);",
       "they are listed to show exactly that effect.  Tolerances asserted by the tests: tests/_parity.py.\n"]
cur = None
for line in open(src):
    line = line.rstrip()
    m = re.match(r"== (\S+) prec=(\S+) bf16_maps=(\S+)(?: streamed=(\S+))?", line)
    if m:
        cur = m.groups()
        out.append(f"\n### {cur[0]} -- precision {cur[1]}, bf16 maps {cur[2]}" + (f", streamed {cur[3]}" if cur[3] else ""))
        out.append("| group | worst error | tensor | tensors | chaotic (max) |")
        out.append("|---|---:|---|---:|---|")
        continue
    m = re.match(r"\s+(\S+)\s+worst (\S+) \((.*?)\) over (\d+) tensors; chaotic (\d+) \(max (\S+)\)", line)
    if m and cur:
        g, e, k, n, nc, ec = m.groups()
        out.append(f"| {g} | {e} | `{k}` | {n} | {nc} ({ec}) |")
open(os.path.join(ROOT, "profiles", "r02_parity.md"), "w").write("\n".join(out) + "\n")
print(len(out), "lines")
