#!/bin/sh
# Run on the GPU box: full ncu capture of the tcgen05 GEMM on a few bf16-mode token shapes; compact CSV exports only.
set -e
B=${1:-256}
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tf32_tc -f -o /tmp/prof_gemm \
    python tools/profile_gemm.py $B > gpurun_out/r2b_ncu_gemm.log 2>&1
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > gpurun_out/r2b_ncu_gemm_raw.csv 2>/dev/null
for i in 0 1 2; do
  ncu -i /tmp/prof_gemm.ncu-rep --page source --csv --launch-skip $i -c 1 > gpurun_out/r2b_ncu_gemm_src$i.csv 2>/dev/null || true
done
ls -la gpurun_out/ | tail -6
