# A/B of the persistent token-GEMM kernel on ONE box
for mode in bf16 tf32; do for p in 0 1 0 1; do
  VU_TC_PERSISTENT=$p VU_TIMER_SHAPES=1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --precision $mode > gpurun_out/r2b_abp_${mode}_$p.json 2> gpurun_out/r2b_abp_${mode}_$p.err
  echo "== $mode persistent=$p"; python tools/step_gemm_shapes.py gpurun_out/r2b_abp_${mode}_$p.json > gpurun_out/r2b_abp_${mode}_$p.txt; head -1 gpurun_out/r2b_abp_${mode}_$p.txt
done; done
sed -n 1,30p gpurun_out/r2b_abp_bf16_1.txt
