// tcgen05 / TMEM / TMA tensor-core GEMM (TF32 inputs, FP32 accumulate).  Placeholder until the kernel lands:
// reports "not handled" so vu_gemm uses the CUDA-core kernel.
#include "vu_common.cuh"
namespace vu {
int gemm_tc(const vu_gemm_desc& d, cudaStream_t s, bool* handled) { (void)d; (void)s; *handled = false; return VU_OK; }
}
