#!/usr/bin/env python
"""Print the parity table of the tensor-core paths against the golden vectors (GPU box): one line per config and
mode with the worst tensors.  Used to calibrate / document tests/_parity.py."""
import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import vit_unet_b200 as vu
from _parity import CONFIGS, build_net, parity_rows


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


names = [a for a in sys.argv[1:] if not a.startswith("--")] or list(CONFIGS)
MODES = (("fp32", False, False), ("tf32", False, False), ("tf32", True, False), ("tf32", True, True), ("bf16", True, False))
if "--bf16" in sys.argv:
    MODES = (("bf16", True, False),)
for prec, maps, streamed in MODES:
    vu.set_precision(prec); vu.set_bf16_maps(maps); vu.set_streamed(streamed)
    for name in names:
        if streamed and not name.startswith(("l2block", "base", "lite")):
            continue
        net, x, y = build_net(name, quiet)
        rows = parity_rows(name, net, x, y, l1_grads=(prec == "fp32"), chaos=(1e-3 if prec == "fp32" else 1e-4))
        by = {}
        for k, e, t, st in rows:
            grp = k.split(":")[0] if ":" in k else k
            a = by.setdefault(grp, [0.0, "", 0, 0, 0.0])
            if st == "chaotic":
                a[3] += 1; a[4] = max(a[4], e)
                continue
            a[2] += 1
            if e >= a[0]:
                a[0], a[1] = e, k
        print(f"== {name} prec={prec} bf16_maps={maps} streamed={streamed}")
        for grp, (e, k, n, nch, ech) in by.items():
            print(f"   {grp:10s} worst {e:.3e} ({k}) over {n} tensors; chaotic {nch} (max {ech:.2e})")
        fails = [r for r in rows if r[3] == "FAIL"]
        for k, e, t, st in sorted(fails, key=lambda r: -r[1])[:6]:
            print(f"   FAIL {k}: {e:.3e} > {t:.2e}")
