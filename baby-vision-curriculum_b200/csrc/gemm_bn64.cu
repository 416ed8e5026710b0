// gemm_bn64.cu -- instantiations of the tcgen05 GEMM for 128 x 64 output tiles.
#include "gemm_kernel.cuh"
namespace bvc {
int gemm_launch_bn64(const bvc_gemm_args* a, int epi, cudaStream_t s) { return gemm_dispatch_bn<64>(a, epi, s); }
}  // namespace bvc
