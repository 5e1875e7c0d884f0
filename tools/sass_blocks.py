#!/usr/bin/env python
"""Basic-block view of an `ncu --page source --csv` export: runs of SASS with equal execution counts."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
blocks = []; cur = None
for r in rows:
    if len(r) < 6: continue
    try: n = int(r[5]); s = int(r[2])
    except ValueError: continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1]); op = m.group(2).split('.')[0] if m else '?'
    if cur is None or cur['n'] != n:
        cur = {'n': n, 'len': 0, 'ops': collections.Counter(), 'stall': 0, 'first': r[1].strip()[:60]}
        blocks.append(cur)
    cur['len'] += 1; cur['ops'][op] += 1; cur['stall'] += s
tot = sum(b['n'] * b['len'] for b in blocks)
ts = sum(b['stall'] for b in blocks)
print('total', tot)
for b in blocks:
    w = b['n'] * b['len']
    if w < 0.01 * tot: continue
    print(f"{100*w/tot:5.1f}% instr  {100*b['stall']/max(ts,1):5.1f}% stall  exec {b['n']:9d} x {b['len']:4d} instrs  top {b['ops'].most_common(6)}  | {b['first']}")
