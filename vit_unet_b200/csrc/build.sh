#!/bin/sh
# Build libvitunet_b200.so for sm_100a (cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libvitunet_b200.so"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC -shared -Xptxas -v \
     -o "$OUT" "$HERE"/vu_api.cu "$HERE"/vu_layout.cu "$HERE"/vu_conv.cu "$HERE"/vu_gemm_simt.cu \
     "$HERE"/vu_gemm_tc.cu "$HERE"/vu_gemm_scores.cu "$HERE"/vu_reattn.cu "$HERE"/vu_reattn_stream.cu "$HERE"/vu_norm.cu "$HERE"/vu_loss.cu "$HERE"/vu_input.cu "$@"
echo "built $OUT"
