#!/bin/sh
# Run on the GPU box: full ncu capture of kernels matching $1 in one Base-L2-shaped block (fwd+bwd), B=$2 images.
set -e
K=$1; B=${2:-32}
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$K" -f -o /tmp/prof_k \
    python tools/profile_block.py $B tf32 > gpurun_out/ncu_k.log 2>&1
ncu -i /tmp/prof_k.ncu-rep --page raw --csv > gpurun_out/ncu_k_raw.csv 2>/dev/null
ncu -i /tmp/prof_k.ncu-rep --page source --csv -c 1 > gpurun_out/ncu_k_src.csv 2>/dev/null || true
ls -la gpurun_out/ | tail -5
