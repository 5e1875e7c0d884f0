"""Per-shape timing of the token GEMMs INSIDE a benchmarked step (CUDA events around every launch, VU_TIMER_SHAPES=1):
usage: VU_TIMER_SHAPES=1 python bench.py ... > line.json; python tools/step_gemm_shapes.py line.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
by, steps = d["roofline"]["by_kernel"], d["steps"]
rows = [(v["ms"] / steps, k, v) for k, v in by.items() if ":tokens" in k]
tot = sum(r[0] for r in rows)
print(f"{d['value']:.1f} images/s, {d['ms_per_step']:.2f} ms/step; token GEMMs {tot:.2f} ms/step")
for ms, k, v in sorted(rows, reverse=True):
    n = v["launches"] // steps
    print(f"{ms:7.3f} ms/step  {n:3d} x {1e3 * ms / n:7.1f} us  {v['tflops']:7.1f} TF/s {v['gbs']:7.0f} GB/s  {k.split(')', 1)[1]}")
