#!/bin/sh
# Run on the GPU box: full ncu capture of the streamed Re-Attention kernels inside one L2-shaped block (fwd+bwd),
# compact CSV exports (raw metrics + per-instruction source page), the big report is dropped.
set -e
B=${1:-64}
VU_STREAMED=1 VU_STREAMED_BWD=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stream_ -f -o /tmp/prof_stream \
    python tools/profile_block.py $B tf32 > gpurun_out/r2_ncu_stream.log 2>&1
ncu -i /tmp/prof_stream.ncu-rep --page raw --csv > gpurun_out/r2_ncu_stream_raw.csv 2>/dev/null
for k in stream_fwd_kernel stream_bwd_ds_kernel stream_bwd_reduce_kernel; do
  ncu -i /tmp/prof_stream.ncu-rep --page source --csv -k regex:$k -c 1 > gpurun_out/r2_ncu_src_$k.csv 2>/dev/null || true
done
ncu -i /tmp/prof_stream.ncu-rep --page details --csv > gpurun_out/r2_ncu_stream_details.csv 2>/dev/null || true
ls -la gpurun_out/ | tail -8
