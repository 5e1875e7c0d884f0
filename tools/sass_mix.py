#!/usr/bin/env python
"""Instruction mix + stall-sample share per SASS opcode from an `ncu --page source --csv` export."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
ops = collections.Counter(); samp = collections.Counter(); tot = 0; ts = 0
for r in rows:
    if len(r) < 6: continue
    try: n = int(r[5]); s = int(r[2])
    except ValueError: continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1])
    if not m: continue
    op = m.group(2).split('.')[0]
    ops[op] += n; samp[op] += s; tot += n; ts += s
print('total warp-instr', tot, 'samples', ts)
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
    print(f'{k:10s} {v:11d} {100*v/tot:5.1f}%  stall samples {100*samp[k]/max(ts,1):5.1f}%')
