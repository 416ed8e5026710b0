// gemm_bn128.cu -- instantiations of the tcgen05 GEMM for 128 x 128 output tiles.
#include "gemm_kernel.cuh"
namespace bvc {
int gemm_launch_bn128(const bvc_gemm_args* a, int epi, cudaStream_t s) { return gemm_dispatch_bn<128>(a, epi, s); }
}  // namespace bvc
