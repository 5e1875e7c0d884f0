#!/bin/sh
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29541 bench.py --gpus 8 --steps 10 --no-cpu-baseline > gpurun_out/r2_n8b_weak256.json 2> gpurun_out/r2_n8b_weak256.err
timeout 300 $T --master-port 29542 bench.py --gpus 8 --steps 10 --batch 64 --no-cpu-baseline --kernel-timing 0 > gpurun_out/r2_n8b_weak64.json 2> gpurun_out/r2_n8b_weak64.err
timeout 300 $T --master-port 29543 bench.py --gpus 8 --steps 10 --global-batch 1024 --no-cpu-baseline > gpurun_out/r2_n8b_strong1024.json 2> gpurun_out/r2_n8b_strong1024.err
timeout 200 python bench.py --gpus 1 --steps 10 --no-cpu-baseline > gpurun_out/r2_n1b_256.json 2>/dev/null
timeout 200 python bench.py --gpus 1 --steps 10 --batch 64 --no-cpu-baseline --kernel-timing 0 > gpurun_out/r2_n1b_64.json 2>/dev/null
timeout 200 python bench.py --gpus 1 --steps 10 --batch 128 --no-cpu-baseline > gpurun_out/r2_n1b_128.json 2>/dev/null
for f in n8b_weak256 n8b_weak64 n8b_strong1024 n1b_256 n1b_64 n1b_128; do python -c "import json;d=json.load(open('gpurun_out/r2_$f.json'));print('$f',round(d['value'],1),round(d['ms_per_step'],2),round(d['e2e']['value'],1))"; done
tail -2 gpurun_out/r2_n8b_weak256.err
