# single-GPU numbers of every bench mode at HEAD (bench default precision = bf16); usage: sh tools/run_workloads.sh [prefix]
PFX=${1:-r2c}
run() { name=$1; shift; timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${PFX}_$name.json 2> gpurun_out/${PFX}_$name.err; python -c "
import json
try:
    d=json.loads(open('gpurun_out/${PFX}_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), d['dtype'], d['roofline'].get('kernel'), d['roofline'].get('frac'))
except Exception as e: print('$name ERR', e)"; }
run base_bf16
run base_tf32 --precision tf32
run base_bf16_b128 --batch 128
run base_bf16_b64 --batch 64 --kernel-timing 0
run lite_infer --workload lite_infer --batch 512
run base_infer --workload base_infer --batch 256
run large_bf16 --workload large_train --batch 128
run large_tf32 --workload large_train --batch 128 --precision tf32
run base1ch_dice --workload base1ch_dice --batch 256
run lite_train --workload lite_train --batch 32
run lite_train_streamed --workload lite_train --batch 32 --streamed 1
