#!/bin/sh
# Run on the GPU box: every kernel of ONE benchmarked Base training step (B images) with its duration and DRAM bytes.
set -e
B=${1:-256}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/r2_step_kernels.csv python tools/profile_step.py $B ${2:-bf16} > gpurun_out/r2_ncu_step.log 2>&1
wc -l gpurun_out/r2_step_kernels.csv
