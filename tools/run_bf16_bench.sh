set -x
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --precision bf16 > gpurun_out/r2b_base_bf16.json 2> gpurun_out/r2b_base_bf16.err
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --precision tf32 > gpurun_out/r2b_base_tf32.json 2> gpurun_out/r2b_base_tf32.err
python bench.py --steps 6 --warmup 3 --workload large_train --batch 128 --precision bf16 > gpurun_out/r2b_large_bf16.json 2> gpurun_out/r2b_large_bf16.err
python - <<'PY'
import json
for f in ("r2b_base_bf16","r2b_base_tf32","r2b_large_bf16"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1))
        by=d["roofline"]["by_kernel"]
        for k,v in sorted(by.items(), key=lambda kv:-kv[1]["ms"])[:14]:
            print("   ", k, round(v["ms"]/d["steps"],2), "ms/step", v["tflops"], "TF", v["gbs"], "GB/s", v["launches"]//d["steps"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
