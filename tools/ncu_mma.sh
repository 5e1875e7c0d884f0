#!/bin/sh
# Run on the GPU box: full ncu capture of the Re-Attention map kernels (mma path) of one Base-L2-shaped block.
set -e
B=${1:-32}
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"mma|softmax|reattn" -f -o /tmp/prof_mma \
    python tools/profile_block.py $B tf32 > gpurun_out/ncu_mma.log 2>&1
ncu -i /tmp/prof_mma.ncu-rep --page raw --csv > gpurun_out/ncu_mma_raw.csv 2>/dev/null
ncu -i /tmp/prof_mma.ncu-rep --page details --csv > gpurun_out/ncu_mma_details.csv 2>/dev/null
for k in softmax_stats_mma_bulk reattn_bwd_rows_mma_cta reattn_mix_reduce_mma reattn_mix_mma; do
  ncu -i /tmp/prof_mma.ncu-rep --page source --csv -k regex:$k -c 1 > gpurun_out/ncu_src_$k.csv 2>/dev/null || true
done
ls -la gpurun_out/ | tail -12
