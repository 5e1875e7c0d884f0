# A/B of the GEMM epilogue batching (rows per batch of prefetched residual loads) on ONE box: same step, three builds
for un in 1 4 8 1 4 8; do
  lib=$PWD/vit_unet_b200/libvitunet_b200_un$un.so
  [ $un = 4 ] && lib=$PWD/vit_unet_b200/libvitunet_b200.so
  VU_LIB_PATH=$lib VU_TIMER_SHAPES=1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --precision bf16 > gpurun_out/r2b_ab_un$un.json 2> gpurun_out/r2b_ab_un$un.err
  echo "== UN=$un"; python tools/step_gemm_shapes.py gpurun_out/r2b_ab_un$un.json | head -12
done
