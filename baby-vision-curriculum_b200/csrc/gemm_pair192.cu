// gemm_pair192.cu -- instantiations of the tcgen05 GEMM for 256 x 192 output tiles on CTA pairs (cta_group::2).
#include "gemm_kernel.cuh"
namespace bvc {
int gemm_launch_pair192(const bvc_gemm_args* a, int epi, cudaStream_t s) { return gemm_dispatch_bn_pair<192, 1>(a, epi, s); }
}  // namespace bvc
