// gemm_bn256.cu -- instantiations of the tcgen05 GEMM for 128 x 256 output tiles.
#include "gemm_kernel.cuh"
namespace bvc {
int gemm_launch_bn256(const bvc_gemm_args* a, int epi, cudaStream_t s) { return gemm_dispatch_bn<256>(a, epi, s); }
}  // namespace bvc
