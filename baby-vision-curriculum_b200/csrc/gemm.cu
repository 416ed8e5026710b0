// gemm.cu -- bvc_gemm_bf16: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[M,N] = epilogue(alpha * A[M,K] . B[N,K]^T)       bf16 operands, fp32 accumulation in TMEM
//
// One CTA per SM (persistent over work items = output tiles x k-splits):
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier full/empty)
//   warp 1      : MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16; commits free smem / publish TMEM)
//   warps 2..9  : epilogue       (tcgen05.ld TMEM -> registers -> fused bias / GELU / GELU' / residual / MSE ->
//                                 vectorised global stores); two TMEM accumulator stages overlap epilogue and MMA
// Operands may be K-major or MN-major in memory (UMMA descriptors + instruction-descriptor major bits), so the
// forward (A.W^T), dgrad (dY.W) and wgrad (dY^T.X) contractions all read the tensors where they already lie.
//
// Replaces: every F.linear / Conv3d-as-GEMM of HF modeling_videomae.py (see include/bvc.h for the line map).
#include "../../include/bvc.h"
#include "bvc_host.h"
#include "bvc_ptx.cuh"

namespace bvc {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
enum GemmEpi { EPI_GENERIC = 0, EPI_PLAIN = 1, EPI_GELU = 2, EPI_DGELU = 3, EPI_RES = 4, EPI_SPLITK = 5, EPI_LOSS = 6 };

int gemm_launch_bn64(const bvc_gemm_args* a, int epi, cudaStream_t s);
int gemm_launch_bn128(const bvc_gemm_args* a, int epi, cudaStream_t s);
int gemm_launch_bn192(const bvc_gemm_args* a, int epi, cudaStream_t s);
int gemm_launch_bn256(const bvc_gemm_args* a, int epi, cudaStream_t s);
int gemm_launch_pair128(const bvc_gemm_args* a, int epi, cudaStream_t s);
int gemm_launch_pair192(const bvc_gemm_args* a, int epi, cudaStream_t s);
int gemm_launch_pair256(const bvc_gemm_args* a, int epi, cudaStream_t s);

// CTA-pair (cta_group::2, 256 x BN) tiles exist for BN = 128 / 256, and for BN = 192 when B is K-major (each CTA
// stages BN / 2 rows of B: whole 64-element swizzle atoms when B is MN-major)
static bool pair_supported(int bn, int b_mn_major) { return bn == 128 || bn == 256 || (bn == 192 && !b_mn_major); }

// cta_pair == 0: choose.  Measured on B200 over every shape of the ViT-B step (tools/gpu_gemm_tune.py,
// profiles/r01_gemm_tune_pair.log): the pair tiles win 5-10 % on every contraction with K >= 1152 (mainloop-bound:
// half the B traffic per flop) and lose on the short-K, epilogue-bound ones (a pair hands its accumulators back in
// lockstep).  Tile width in pair mode: 256, except a residual epilogue on an N that is not a multiple of 256 (the
// fp32 residual rows of the wasted columns cost more than the narrower tile).
static bool pick_pair_auto(const bvc_gemm_args* a) {
  return a->block_n == 0 && a->target == nullptr && a->K >= 1152 && a->N >= 256 && a->M >= 256;
}
static int pair_auto_block_n(const bvc_gemm_args* a) { return (a->N % 256 != 0 && a->res != nullptr) ? 128 : 256; }

static int pick_block_n(int M, int N, int K, bool wgrad) {
  // measured on B200 over every shape of the ViT-B step (tools/gpu_gemm_tune.py, profiles/r01_gemm_tune.log)
  if (wgrad) {
    if (N % 256 == 0) return 256;
    if (N % 192 == 0) return 192;
  } else if (K <= 768 && N % 192 == 0 && N >= 1152) {
    return 192;  // short-K, epilogue-heavy: three 32-column chunks per warp hand the accumulator back sooner
  }
  // fewest wasted columns first, then the widest tile (B traffic and MMA efficiency), with at least ~1 wave.
  const int cands[4] = {256, 192, 128, 64};
  int best = 64;
  double best_cost = 1e30;
  const int sms = num_sms();
  const int tiles_m = (M + BM - 1) / BM;
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    const int tiles_n = (N + bn - 1) / bn;
    const long long tiles = (long long)tiles_m * tiles_n;
    const long long waves = (tiles + sms - 1) / sms;
    // time ~ waves * (tile cost); tile cost ~ bn (MMA) with a floor for narrow tiles (smem-bound below 128)
    const double tile_cost = (double)(bn < 128 ? 128 : bn) + 24.0;
    const double cost = (double)waves * tile_cost;
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

__global__ void loss_finalize_kernel(const float* __restrict__ partials, long long n, double inv_numel,
                                     const int* __restrict__ status, float* __restrict__ loss) {
  __shared__ double sh[256];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 256) s += (double)partials[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float v = (float)(sh[0] * inv_numel);
    if (status && *status != 0) v = __int_as_float(0x7fc00000);
    *loss = v;
  }
}

}  // namespace bvc

extern "C" int bvc_abi_version(void) { return BVC_ABI_VERSION; }

static int resolve_block_n(int M, int N, int block_n, int K = 1 << 30, bool wgrad = false) {
  if (block_n == 64 || block_n == 128 || block_n == 192 || block_n == 256) return block_n;
  return bvc::pick_block_n(M, N, K, wgrad);
}

extern "C" int64_t bvc_gemm_loss_slots(int32_t M, int32_t N, int32_t block_n) {
  const int bn = resolve_block_n(M, N, block_n);
  return (int64_t)((M + bvc::BM - 1) / bvc::BM) * ((N + bn - 1) / bn) * bvc::kEpiWarps;
}

extern "C" int bvc_gemm_bf16(const bvc_gemm_args* a, void* stream) {
  BVC_CHECK_ARG(a != nullptr && a->a != nullptr && a->b != nullptr);
  BVC_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0);
  BVC_CHECK_ARG(a->N % 8 == 0 && a->lda % 8 == 0 && a->ldb % 8 == 0 && a->ldo % 8 == 0);
  BVC_CHECK_ARG(a->block_n == 0 || a->block_n == 64 || a->block_n == 128 || a->block_n == 192 || a->block_n == 256);
  BVC_CHECK_ARG(a->cta_pair >= 0 && a->cta_pair <= 2);
  BVC_CHECK_ARG(a->out_f32 != nullptr || a->out_bf16 != nullptr);
  BVC_CHECK_ARG((((uintptr_t)a->a) & 15) == 0 && (((uintptr_t)a->b) & 15) == 0);
  BVC_CHECK_ARG(a->act == 0 || a->act == 1 || (a->act == 2 && a->aux_in != nullptr));
  BVC_CHECK_ARG(a->act == 0 || a->ld_aux % 8 == 0);
  BVC_CHECK_ARG(a->res == nullptr || a->ldr % 4 == 0);
  BVC_CHECK_ARG(a->target == nullptr || (a->ldt % 4 == 0 && a->loss_partial != nullptr));
  BVC_CHECK_ARG(a->colsum == nullptr || (a->target == nullptr && a->k_splits == 1 && a->out_seg == 0));
  BVC_CHECK_ARG(a->a_mn_major == 0 ? a->lda >= a->K : a->lda >= a->M);
  BVC_CHECK_ARG(a->b_mn_major == 0 ? a->ldb >= a->K : a->ldb >= a->N);
  cudaStream_t s = (cudaStream_t)stream;
  // loss GEMMs resolve the tile width exactly like bvc_gemm_loss_slots() (which does not know K)
  int bn = a->target ? resolve_block_n(a->M, a->N, a->block_n)
                     : resolve_block_n(a->M, a->N, a->block_n, a->K, a->a_mn_major && a->b_mn_major);
  bool pair = a->cta_pair == 2;
  if (a->cta_pair == 0 && bvc::pick_pair_auto(a)) {
    pair = true;
    bn = bvc::pair_auto_block_n(a);
  }
  BVC_CHECK_ARG(!pair || bvc::pair_supported(bn, a->b_mn_major));
  // pick the leanest epilogue variant that covers the request (gemm_kernel.cuh); anything unusual -> generic
  int epi = bvc::EPI_GENERIC;
  const bool seg = a->out_seg > 0;
  {
    const int kb_total = (a->K + bvc::BK - 1) / bvc::BK;
    int ks = a->k_splits;
    if (ks <= 0) {  // mirror of resolve_k_splits(): only the "is it > 1" answer is needed here
      const int tile_m = pair ? 2 * bvc::BM : bvc::BM, workers = pair ? bvc::num_sms() / 2 : bvc::num_sms();
      const long long tiles = (long long)((a->M + tile_m - 1) / tile_m) * ((a->N + bn - 1) / bn);
      ks = (int)((2LL * workers + tiles - 1) / tiles);
      if (ks > kb_total / 4) ks = kb_total / 4 > 0 ? kb_total / 4 : 1;
    }
    if (ks > kb_total) ks = kb_total;
    if (ks < 1) ks = 1;
    const int per = (kb_total + ks - 1) / ks;
    const bool split = (kb_total + per - 1) / per > 1;
    if (split)
      epi = bvc::EPI_SPLITK;
    else if (a->target && a->out_bf16 && !a->out_f32 && a->act == 0 && !a->res && !seg)
      epi = bvc::EPI_LOSS;
    else if (a->act == 1 && a->out_bf16 && !a->out_f32 && !a->res && !a->target && !seg)
      epi = bvc::EPI_GELU;
    else if (a->act == 2 && a->out_bf16 && !a->out_f32 && !a->res && !a->target && !seg && !a->bias)
      epi = bvc::EPI_DGELU;
    else if (a->res && !a->res_idx && a->out_f32 && !a->out_bf16 && a->act == 0 && !a->target && !seg)
      epi = bvc::EPI_RES;
    else if (a->act == 0 && !a->res && !a->target && !seg && a->out_bf16 && !a->out_f32)
      epi = bvc::EPI_PLAIN;
  }
  if (pair) {
    switch (bn) {
      case 256: return bvc::gemm_launch_pair256(a, epi, s);
      case 192: return bvc::gemm_launch_pair192(a, epi, s);
      default: return bvc::gemm_launch_pair128(a, epi, s);
    }
  }
  switch (bn) {
    case 256: return bvc::gemm_launch_bn256(a, epi, s);
    case 192: return bvc::gemm_launch_bn192(a, epi, s);
    case 128: return bvc::gemm_launch_bn128(a, epi, s);
    default: return bvc::gemm_launch_bn64(a, epi, s);
  }
}

extern "C" int bvc_loss_finalize(const float* partials, int64_t n, double numel, const int32_t* status, float* loss,
                                 void* stream) {
  BVC_CHECK_ARG(partials != nullptr && loss != nullptr && n > 0 && numel > 0);
  bvc::loss_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, (long long)n, 1.0 / numel, status, loss);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}
