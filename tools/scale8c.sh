#!/bin/sh
# 8-GPU runs at HEAD (bench default precision bf16): weak 256 img/GPU, strong global 1024, Large bf16 global 1024 (configs[3])
PFX=${1:-r2c}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $T --master-port 29541 bench.py --gpus 8 --steps 8 --no-cpu-baseline > gpurun_out/${PFX}_n8_weak256.json 2> gpurun_out/${PFX}_n8_weak256.err
timeout 240 $T --master-port 29543 bench.py --gpus 8 --steps 8 --global-batch 1024 --no-cpu-baseline --kernel-timing 0 > gpurun_out/${PFX}_n8_strong1024.json 2> gpurun_out/${PFX}_n8_strong1024.err
timeout 240 $T --master-port 29544 bench.py --gpus 8 --steps 8 --workload large_train --global-batch 1024 --dtype bf16 --no-cpu-baseline --kernel-timing 0 > gpurun_out/${PFX}_n8_large_bf16.json 2> gpurun_out/${PFX}_n8_large_bf16.err
for f in n8_weak256 n8_strong1024 n8_large_bf16; do python -c "import json;d=json.loads(open('gpurun_out/${PFX}_$f.json').read().strip().splitlines()[-1]);print('$f',round(d['value'],1),round(d['ms_per_step'],2),round(d['e2e']['value'],1), d['dtype'], d['config']['global_batch'])"; done
tail -2 gpurun_out/${PFX}_n8_weak256.err
