// 3xTF32 tcgen05 GEMM (placeholder until the tensor-core path lands; the exact-fp32 FFMA kernel runs).
#include "gemm.cuh"
bool fc_gemm_tc_supported(const GemmArgs&) { return false; }
int fc_launch_gemm_tc(const GemmArgs&, cudaStream_t) { return FC_ERR_UNSUPPORTED; }
