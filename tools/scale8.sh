#!/bin/sh
# 8-GPU runs of the BASELINE configs that need a full box (run under: gpurun --gpus 8)
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29511 bench.py --gpus 8 --steps 10 --no-cpu-baseline > gpurun_out/r2_n8_weak256.json 2> gpurun_out/r2_n8_weak256.err
timeout 300 $T --master-port 29512 bench.py --gpus 8 --steps 10 --global-batch 1024 --no-cpu-baseline > gpurun_out/r2_n8_strong1024.json 2> gpurun_out/r2_n8_strong1024.err
timeout 300 $T --master-port 29513 bench.py --gpus 8 --steps 10 --batch 64 --no-cpu-baseline --kernel-timing 0 > gpurun_out/r2_n8_weak64.json 2> gpurun_out/r2_n8_weak64.err
timeout 300 $T --master-port 29514 bench.py --gpus 8 --steps 10 --workload base1ch_dice --batch 256 --no-cpu-baseline > gpurun_out/r2_n8_dice.json 2> gpurun_out/r2_n8_dice.err
timeout 300 $T --master-port 29515 bench.py --gpus 8 --steps 5 --workload large_train --global-batch 1024 --no-cpu-baseline > gpurun_out/r2_n8_large1024.json 2> gpurun_out/r2_n8_large1024.err
timeout 200 python bench.py --gpus 1 --steps 10 --batch 128 --no-cpu-baseline > gpurun_out/r2_n1_b128.json 2>/dev/null
timeout 200 python bench.py --gpus 1 --steps 10 --batch 64 --no-cpu-baseline --kernel-timing 0 > gpurun_out/r2_n1_b64.json 2>/dev/null
for f in weak256 strong1024 weak64 dice large1024; do head -c 220 gpurun_out/r2_n8_$f.json; echo; done
