#!/bin/sh
# Run on the GPU box: full ncu capture of one L2-shaped block (fwd+bwd), export compact CSVs, drop the big report.
set -e
B=${1:-32}
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/prof_block \
    python tools/profile_block.py $B ${2:-bf16} > gpurun_out/ncu_block.log 2>&1
ncu -i /tmp/prof_block.ncu-rep --page raw --csv > gpurun_out/ncu_block_raw.csv 2>/dev/null
for k in reattn_bwd_rows_mma_cta softmax_stats_mma_bulk reattn_mix_reduce_mma reattn_mix_mma scores_mma; do
  ncu -i /tmp/prof_block.ncu-rep --page source --csv -k regex:$k -c 1 > gpurun_out/ncu_src_$k.csv 2>/dev/null || true
done
ls -la gpurun_out/
