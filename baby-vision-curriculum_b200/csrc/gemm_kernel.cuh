// gemm_kernel.cuh -- the tcgen05 GEMM kernel template and its launcher (see gemm.cu for the overview).  Included by
// one translation unit per tile width (gemm_bn*.cu) so that the instantiations compile in parallel.
#pragma once
#include "../../include/bvc.h"
#include "bvc_host.h"
#include "bvc_ptx.cuh"

namespace bvc {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;

// compile-time epilogue variants: each instantiation carries only the code it needs (the all-runtime-flags epilogue
// is ~56 KB of SASS and instruction-cache bound: ncu source view showed stall_no_inst on every flag test)
enum GemmEpi { EPI_GENERIC = 0, EPI_PLAIN = 1, EPI_GELU = 2, EPI_DGELU = 3, EPI_RES = 4, EPI_SPLITK = 5, EPI_LOSS = 6 };

struct GemmParams {
  int M, N, K;
  int k_splits, kb_total, kb_per_split;
  int tiles_m, tiles_n;
  float* out_f32;
  bf16* out_bf16;
  long long ldo;
  int out_seg, out_seg_stride, out_seg_off;
  float alpha;
  const float* alpha_dev;
  const float* bias;
  int act;
  bf16* aux_out;
  const bf16* aux_in;
  long long ld_aux;
  const float* res;
  long long ldr;
  const int* res_idx;
  const float* target;
  long long ldt;
  float* loss_partial;
  bf16* logits_out;
  float* colsum;
  int has_pre;  // tma_pre describes the epilogue's global operand (residual / target / GELU pre-activation)
};

// PAIR = 1: the kernel runs as clusters of two CTAs (the two SMs of a TPC) on 256 x BN output tiles with
// tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of A and HALF (BN / 2 rows) of the B tile and owns the
// accumulator of its 128 rows; see the protocol notes in bvc_ptx.cuh.  Per 128 x BN x 64 of MMA work an SM then pulls
// 16 + BN / 8 KB through the L2 -> SM fabric instead of 16 + BN / 4 KB, which is what capped the one-CTA tiles near
// 1050 TFLOP/s.
template <int BN, int PAIR = 0>
struct GemmCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBRows = PAIR ? BN / 2 : BN;   // B rows staged by one CTA
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = PAIR ? ((BN == 256) ? 6 : (BN == 192) ? 6 : 8)
                                      : ((BN == 256) ? 4 : (BN == 192) ? 4 : (BN == 128) ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kStagingBytes = kEpiWarps * 4096;  // per epilogue warp: 32 rows x 32 fp32, 128B-swizzled
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Phi(-a) = 2^(-Q(a)) for a in [0, 6]: degree-7 minimax fit of Q(a) = -log2(Phi(-a)) (tools/fit_gelu.py); relative
// error of Phi(-a) <= 3.3e-6 over the whole range (tails included), i.e. far below the bf16 rounding of the
// activation that follows.  One MUFU.EX2 and 7 FMAs instead of erff().
__device__ __forceinline__ float phi_neg_abs(float x) {
  const float a = fminf(fabsf(x), 6.0f);
  float q = -1.8348840982e-06f;
  q = fmaf(q, a, 6.1599723096e-05f);
  q = fmaf(q, a, -9.3053573547e-04f);
  q = fmaf(q, a, 8.5079311974e-03f);
  q = fmaf(q, a, -5.3960143443e-02f);
  q = fmaf(q, a, -4.5846433058e-01f);
  q = fmaf(q, a, -1.1512510639e+00f);
  q = fmaf(q, a, -9.9999529365e-01f);
  return exp2f(q);  // = Phi(-|x|)
}
// exact-erf GELU (HF:316, torch.nn.functional.gelu default): x * Phi(x)
__device__ __forceinline__ float gelu_erf(float x) {
  const float xw = x * phi_neg_abs(x);
  return x > 0.f ? x - xw : xw;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float w = phi_neg_abs(x);
  const float cdf = x > 0.f ? 1.0f - w : w;
  const float pdf = 0.3989422804014327f * exp2f(-0.72134752044448170f * x * x);
  return fmaf(x, pdf, cdf);
}

// ---- two elements at a time on the packed fp32x2 pipe (interior-tile fast path of the epilogue) ----
// Phi(-|x|) for a pair: the same degree-7 fit as phi_neg_abs, 7 FFMA2 + 2 MUFU.EX2
__device__ __forceinline__ uint64_t phi_neg_abs2(uint64_t a2) {
  uint64_t q = pack2(-1.8348840982e-06f, -1.8348840982e-06f);
  q = ffma2(q, a2, pack2(6.1599723096e-05f, 6.1599723096e-05f));
  q = ffma2(q, a2, pack2(-9.3053573547e-04f, -9.3053573547e-04f));
  q = ffma2(q, a2, pack2(8.5079311974e-03f, 8.5079311974e-03f));
  q = ffma2(q, a2, pack2(-5.3960143443e-02f, -5.3960143443e-02f));
  q = ffma2(q, a2, pack2(-4.5846433058e-01f, -4.5846433058e-01f));
  q = ffma2(q, a2, pack2(-1.1512510639e+00f, -1.1512510639e+00f));
  q = ffma2(q, a2, pack2(-9.9999529365e-01f, -9.9999529365e-01f));
  float q0, q1;
  unpack2(q, q0, q1);
  return pack2(exp2f(q0), exp2f(q1));
}
__device__ __forceinline__ uint64_t abs2(uint64_t x2) { return x2 & 0x7fffffff7fffffffull; }
// gelu(x) = x Phi(x) = 0.5 x + |x| (0.5 - Phi(-|x|))   (no compare / select)
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t x2) {
  const uint64_t a2 = abs2(x2);
  const uint64_t half2 = pack2(0.5f, 0.5f);
  const uint64_t t2 = ffma2(phi_neg_abs2(a2), pack2(-1.f, -1.f), half2);  // 0.5 - w
  return ffma2(x2, half2, fmul2(a2, t2));
}
// gelu'(x) = Phi(x) + x phi(x),  Phi(x) = 0.5 + copysign(0.5 - Phi(-|x|), x)
__device__ __forceinline__ uint64_t gelu_erf_grad2(uint64_t x2) {
  const uint64_t a2 = abs2(x2);
  const uint64_t half2 = pack2(0.5f, 0.5f);
  const uint64_t t2 = ffma2(phi_neg_abs2(a2), pack2(-1.f, -1.f), half2);  // 0.5 - w  (>= 0)
  const uint64_t cdf2 = fadd2(half2, t2 | (x2 & 0x8000000080000000ull));
  float e0, e1;
  unpack2(fmul2(fmul2(x2, x2), pack2(-0.72134752044448170f, -0.72134752044448170f)), e0, e1);
  const uint64_t pdf2 = fmul2(pack2(exp2f(e0), exp2f(e1)), pack2(0.3989422804014327f, 0.3989422804014327f));
  return ffma2(x2, pdf2, cdf2);
}
// gelu(x) and gelu'(x) from one evaluation of Phi(-|x|) (the forward epilogue stores gelu' for the backward GEMM)
__device__ __forceinline__ void gelu_erf_both2(uint64_t x2, uint64_t& y2, uint64_t& g2) {
  const uint64_t a2 = abs2(x2);
  const uint64_t half2 = pack2(0.5f, 0.5f);
  const uint64_t t2 = ffma2(phi_neg_abs2(a2), pack2(-1.f, -1.f), half2);  // 0.5 - Phi(-|x|)  (>= 0)
  y2 = ffma2(x2, half2, fmul2(a2, t2));
  const uint64_t cdf2 = fadd2(half2, t2 | (x2 & 0x8000000080000000ull));
  // phi(x) = 2^(-x^2 / (2 ln 2) + log2(1 / sqrt(2 pi)))
  float e0, e1;
  unpack2(ffma2(fmul2(x2, x2), pack2(-0.72134752044448170f, -0.72134752044448170f),
                pack2(-1.3257480647361593f, -1.3257480647361593f)),
          e0, e1);
  g2 = ffma2(x2, pack2(exp2f(e0), exp2f(e1)), cdf2);
}
__device__ __forceinline__ uint64_t bf16x2_to_f32x2(uint32_t h) {
  return pack2(__uint_as_float(h << 16), __uint_as_float(h & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f32x2_to_bf16x2(uint64_t v) {
  float a, b;
  unpack2(v, a, b);
  return pack_bf16x2(a, b);
}

template <int BN, int A_MN, int B_MN, int EPI, int PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_pre, const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, PAIR>;
  static_assert(!PAIR || (Cfg::kBRows % (B_MN ? 64 : 8) == 0), "pair tile: B half must be whole swizzle atoms");
  // pair mode: rank of this CTA in its cluster (0 = leader: owns the full barriers and issues the MMAs); the cluster
  // (not the CTA) is the persistent worker
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int kTileM = PAIR ? 2 * BM : BM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], PAIR ? 2 * kEpiWarps : kEpiWarps);  // pair: the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    else tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything is signalled on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_work = p.tiles_m * p.tiles_n * p.k_splits;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = worker; w < total_work; w += n_workers) {
        const int split = w % p.k_splits;
        const int tile = w / p.k_splits;
        const int m0 = (tile / p.tiles_n) * kTileM + (int)rank * BM;
        const int n0 = (tile % p.tiles_n) * BN;
        const int nb0 = n0 + (int)rank * Cfg::kBRows;  // first B row this CTA stages
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        // pull this tile's epilogue operand into L2 while its MMAs run: the epilogue's loads then see L2 latency,
        // not DRAM latency (the epilogue warps keep only ~4 KB each in flight)
        if ((EPI == EPI_RES || EPI == EPI_LOSS || EPI == EPI_DGELU) && p.has_pre)
          asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(
                           reinterpret_cast<uint64_t>(&tma_pre)),
                       "r"(n0), "r"(m0)
                       : "memory");
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (PAIR) {
            // both CTAs' bytes land on the leader's barrier; the leader arms it for the pair
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            const uint32_t lbar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (A_MN == 0) {
              tma_load_2d_pair(sa, &tma_a, lbar, kb * BK, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d_pair(sa + j * 8192, &tma_a, lbar, m0 + j * 64, kb * BK);
            }
            if (B_MN == 0) {
              tma_load_2d_pair(sb, &tma_b, lbar, kb * BK, nb0);
            } else {
#pragma unroll
              for (int j = 0; j < Cfg::kBRows / 64; ++j)
                tma_load_2d_pair(sb + j * 8192, &tma_b, lbar, nb0 + j * 64, kb * BK);
            }
          } else {
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            if (A_MN == 0) {
              tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BK, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d(sa + j * 8192, &tma_a, &full_bar[stage], m0 + j * 64, kb * BK);
            }
            if (B_MN == 0) {
              tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(sb + j * 8192, &tma_b, &full_bar[stage], n0 + j * 64, kb * BK);
            }
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer.  The whole warp walks the warp-uniform control flow and polls the barriers; one elected lane issues
    // the tcgen05 instructions.  (With the entire loop under `if (lane == 0)` ptxas kept the loop state in vector
    // registers and spent ~16 instructions -- R2UR moves, descriptor re-masking, an ELECT retry loop -- per MMA,
    // more than a 128 x N x 16 MMA takes to execute for N <= 192.)
    constexpr uint32_t idesc = umma_idesc_bf16(BN, A_MN, B_MN, kTileM);
    // descriptor start-address step of one K = 16 slice, in 16-byte units: 2048 B (MN-major) or 32 B (K-major)
    constexpr uint64_t kStepA = A_MN ? 128 : 2, kStepB = B_MN ? 128 : 2;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    // pair mode: only the leader CTA issues (its MMAs drive the tensor cores of both SMs); the peer's warp 1 just
    // holds the TMEM allocation
    for (int w = (PAIR && rank != 0) ? total_work : worker; w < total_work; w += n_workers) {
      const int split = w % p.k_splits;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
          const uint64_t da = A_MN ? umma_smem_desc(a_addr, 1024, 8192) : umma_smem_desc(a_addr, 1024, 16);
          const uint64_t db = B_MN ? umma_smem_desc(b_addr, 1024, 8192) : umma_smem_desc(b_addr, 1024, 16);
          const uint32_t acc0 = kb > kb0;
          if (PAIR) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss_pair(d_tmem, da + kStepA * k, db + kStepB * k, idesc, acc0 | (k > 0));
            umma_commit_pair(&empty_bar[stage]);  // frees the stage in both CTAs
            if (kb + 1 == kb1) umma_commit_pair(&tmem_full[acc]);
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss(d_tmem, da + kStepA * k, db + kStepB * k, idesc, acc0 | (k > 0));
            umma_commit(&empty_bar[stage]);
            if (kb + 1 == kb1) umma_commit(&tmem_full[acc]);
          }
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (kb0 >= kb1) {  // empty K range (never produced by the host-side split choice): keep the protocol alive
        if (elect_one()) {
          if (PAIR) umma_commit_pair(&tmem_full[acc]);
          else umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    // Phase 1: tcgen05.ld gives every lane one accumulator ROW (32 columns); the warp transposes the 32x32 fp32
    // chunk through a 128B-swizzled staging buffer.  Phase 2: 8 lanes cover one row's 128 bytes, 4 rows per
    // instruction, so every global access of the fused epilogue (residual, target, aux, outputs) is coalesced.
    const int e = warp - 2;
    const int q = warp & 3;          // TMEM lane quadrant this warp may read
    const int half = e >> 2;         // which half of the BN columns
    constexpr int kColsPerWarp = BN / 2;
    const uint32_t stg_s = smem_u32(staging + e * 4096);  // explicit LDS / STS (see lds_f4 in bvc_ptx.cuh)
    const int c4 = lane & 7;         // phase 2: this lane's 4-column group inside the 32-column chunk
    const int rsub = lane >> 3;      // phase 2: row within each group of 4 rows
    const float alpha = p.alpha * (p.alpha_dev ? __ldg(p.alpha_dev) : 1.0f);
    constexpr bool G = EPI == EPI_GENERIC;
    constexpr bool kFast = EPI == EPI_PLAIN || EPI == EPI_GELU || EPI == EPI_DGELU || EPI == EPI_RES;
    const bool kSplit = G ? p.k_splits > 1 : EPI == EPI_SPLITK;
    const bool kGelu = G ? p.act == 1 : EPI == EPI_GELU;
    const bool kDgelu = G ? p.act == 2 : EPI == EPI_DGELU;
    const bool kRes = G ? p.res != nullptr : EPI == EPI_RES;
    const bool kLoss = G ? p.target != nullptr : EPI == EPI_LOSS;
    const bool kSeg = G;
    const bool kOutF32 = G ? p.out_f32 != nullptr : EPI == EPI_RES;
    const bool kOutBf16 = G ? p.out_bf16 != nullptr : (EPI == EPI_PLAIN || EPI == EPI_GELU || EPI == EPI_DGELU ||
                                                         EPI == EPI_LOSS);
    int acc = 0;
    uint32_t acc_phase = 0;
    // pair mode: both CTAs hand their accumulator stage back to the leader's MMA warp
    const uint32_t tmem_empty_leader0 = PAIR ? mapa_shared(smem_u32(&tmem_empty[0]), 0) : 0u;
    const uint32_t tmem_empty_leader1 = PAIR ? mapa_shared(smem_u32(&tmem_empty[1]), 0) : 0u;
    for (int w = worker; w < total_work; w += n_workers) {
      const int tile = w / p.k_splits;
      const int m0 = (tile / p.tiles_n) * kTileM + (int)rank * BM;
      const int n0 = (tile % p.tiles_n) * BN;
      const int rbase = m0 + q * 32;
      const bool interior = m0 + BM <= p.M && n0 + BN <= p.N;  // warp-uniform
      // interior fast path: the global operand of the fused epilogue (fp32 residual rows / saved gelu' factors) is
      // software-pipelined one 32-column chunk ahead; chunk 0 is requested here, BEFORE waiting for the accumulator,
      // so its latency hides behind this tile's MMAs instead of following them
      float4 nxt_f[8];
      uint2 nxt_h[8];
      auto prefetch_chunk = [&](int cc_n) {
        const long long r0 = rbase + rsub;
        const int col_n = n0 + half * kColsPerWarp + cc_n + c4 * 4;
        if (EPI == EPI_RES) {
          const float* fb = p.res + r0 * p.ldr + col_n;
#pragma unroll
          for (int it = 0; it < 8; ++it)
            asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(nxt_f[it].x), "=f"(nxt_f[it].y), "=f"(nxt_f[it].z), "=f"(nxt_f[it].w)
                         : "l"(fb + (long long)it * 4 * p.ldr)
                         : "memory");
        }
        if (EPI == EPI_DGELU) {
          const bf16* hb = p.aux_in + r0 * p.ld_aux + col_n;
#pragma unroll
          for (int it = 0; it < 8; ++it)
            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];"
                         : "=r"(nxt_h[it].x), "=r"(nxt_h[it].y)
                         : "l"(hb + (long long)it * 4 * p.ld_aux)
                         : "memory");
        }
      };
      if (kFast && interior && (EPI == EPI_RES || EPI == EPI_DGELU)) prefetch_chunk(0);
      // the bias of the chunk after the current one is always in flight (volatile: ptxas otherwise sinks the load next
      // to its use, and its L2 latency was exposed once per 32-column chunk: 5 % of the GELU GEMM's stall samples)
      float4 bias_nxt = make_float4(0.f, 0.f, 0.f, 0.f);
      auto bias_chunk = [&](int cc_n) {
        const int col_n = n0 + half * kColsPerWarp + cc_n + c4 * 4;
        if (p.bias && col_n < p.N) bias_nxt = ldv_f4(p.bias + col_n);
      };
      bias_chunk(0);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      float lsum = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < kColsPerWarp; cc += 32) {
        const int c_tile = half * kColsPerWarp + cc;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c_tile), v);
        tmem_ld_wait();
        if (cc + 32 >= kColsPerWarp) {  // last chunk of this tile is in registers: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(acc ? tmem_empty_leader1 : tmem_empty_leader0);
            else mbar_arrive(&tmem_empty[acc]);
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts_u4(stg_s + lane * 128 + ((c ^ (lane & 7)) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        __syncwarp();
        const int col = n0 + c_tile + c4 * 4;
        const float4 bias_cur = bias_nxt;
        if (cc + 32 < kColsPerWarp) bias_chunk(cc + 32);
        // this lane's 8 rows of the transposed chunk, all requested before any of the row loop's global stores
        float4 tv[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rl = it * 4 + rsub;
          tv[it] = lds_f4(stg_s + rl * 128 + ((c4 ^ (rl & 7)) << 4));
        }
        if (kFast && interior) {
          // ---------------------------------------------------------------- interior tile, non-generic variant:
          // no row / column guards, no segment remap, pointers advance by a constant stride, math on fp32x2 pairs.
          // (The guarded path below costs ~30 instructions per output element on the GELU variant -- ncu: FFMA 8,
          // IMAD/IADD3/ISETP/BRA 9 -- which made the K = 384 decoder GEMMs epilogue-bound.)
          const long long r0 = rbase + rsub;
          const uint64_t b01 = pack2(bias_cur.x, bias_cur.y), b23 = pack2(bias_cur.z, bias_cur.w);  // zeros without bias
          const uint64_t al2 = pack2(alpha, alpha);
          float4 pre_f[8];
          uint2 pre_h[8];
          if (EPI == EPI_RES || EPI == EPI_DGELU) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              pre_f[it] = nxt_f[it];
              pre_h[it] = nxt_h[it];
            }
            if (cc + 32 < kColsPerWarp) prefetch_chunk(cc + 32);  // in flight while this chunk is processed
          }
          // the dispatcher (gemm.cu) guarantees: EPI_RES writes fp32 only, the other fast variants bf16 only
          bf16* ob = EPI != EPI_RES ? p.out_bf16 + r0 * p.ldo + col : nullptr;
          float* of = EPI == EPI_RES ? p.out_f32 + r0 * p.ldo + col : nullptr;
          bf16* oa = (EPI == EPI_GELU && p.aux_out) ? p.aux_out + r0 * p.ld_aux + col : nullptr;
          uint64_t cs01 = 0, cs23 = 0;  // (0.f, 0.f): this lane's column sums over its 8 rows
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const float4 t = tv[it];
            // alpha * acc + bias as one FFMA2 per pair (x * 1.0f is exact; a runtime `if (alpha != 1)` cost a predicated
            // multiply plus three register moves per pair: 2.2 of the GELU epilogue's 21 instructions per element)
            uint64_t v01 = ffma2(pack2(t.x, t.y), al2, b01), v23 = ffma2(pack2(t.z, t.w), al2, b23);
            if (EPI == EPI_GELU) {
              // the activation is evaluated on the bf16-rounded pre-activation (what HF's bf16 Linear output is);
              // aux_out receives gelu'(x) -- all the backward GEMM needs of this layer's pre-activation
              uint64_t g01, g23;
              gelu_erf_both2(bf16x2_to_f32x2(f32x2_to_bf16x2(v01)), v01, g01);
              gelu_erf_both2(bf16x2_to_f32x2(f32x2_to_bf16x2(v23)), v23, g23);
              if (oa) {
                uint2 pk;
                pk.x = f32x2_to_bf16x2(g01);
                pk.y = f32x2_to_bf16x2(g23);
                *reinterpret_cast<uint2*>(oa + (long long)it * 4 * p.ld_aux) = pk;
              }
            } else if (EPI == EPI_DGELU) {
              v01 = fmul2(v01, bf16x2_to_f32x2(pre_h[it].x));
              v23 = fmul2(v23, bf16x2_to_f32x2(pre_h[it].y));
            } else if (EPI == EPI_RES) {
              v01 = fadd2(v01, pack2(pre_f[it].x, pre_f[it].y));
              v23 = fadd2(v23, pack2(pre_f[it].z, pre_f[it].w));
            }
            if (EPI == EPI_DGELU || EPI == EPI_PLAIN) {
              cs01 = fadd2(cs01, v01);
              cs23 = fadd2(cs23, v23);
            }
            if (EPI == EPI_RES) {
              float x0, x1, x2, x3;
              unpack2(v01, x0, x1);
              unpack2(v23, x2, x3);
              *reinterpret_cast<float4*>(of + (long long)it * 4 * p.ldo) = make_float4(x0, x1, x2, x3);
            }
            if (EPI != EPI_RES) {
              uint2 pk;
              pk.x = f32x2_to_bf16x2(v01);
              pk.y = f32x2_to_bf16x2(v23);
              *reinterpret_cast<uint2*>(ob + (long long)it * 4 * p.ldo) = pk;
            }
          }
          if ((EPI == EPI_DGELU || EPI == EPI_PLAIN) && p.colsum) {
            // fused bias gradient: lanes that share a column group differ in rsub (lane bits 3, 4)
            float c0, c1, c2, c3;
            unpack2(cs01, c0, c1);
            unpack2(cs23, c2, c3);
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              c0 += __shfl_xor_sync(0xffffffffu, c0, o);
              c1 += __shfl_xor_sync(0xffffffffu, c1, o);
              c2 += __shfl_xor_sync(0xffffffffu, c2, o);
              c3 += __shfl_xor_sync(0xffffffffu, c3, o);
            }
            if (rsub == 0)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.colsum + col), "f"(c0), "f"(c1),
                           "f"(c2), "f"(c3)
                           : "memory");
          }
          __syncwarp();
          continue;
        }
        if (col < p.N) {
          const float4 bias4 = bias_cur;
          // global operands of the fused epilogue first, all 8 rows in flight at once (issuing them inside the
          // row loop serialises one DRAM round trip per row: measured 5x slower on the residual / GELU' GEMMs)
          float4 pre_f[8];
          uint2 pre_h[8];
          if (!G && (kRes || kLoss || kDgelu)) {
            // straight-line: clamp the row instead of predicating so the 8 loads issue back to back (the ncu source
            // view of the predicated version showed one full memory round trip per row)
            const float* fbase = kRes ? p.res : p.target;
            const long long fld = kRes ? p.ldr : p.ldt;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rc = min(rbase + it * 4 + rsub, p.M - 1);
              // volatile + memory clobber: ptxas otherwise sinks every load next to its use (one round trip per row)
              if (kRes || kLoss)
                asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(pre_f[it].x), "=f"(pre_f[it].y), "=f"(pre_f[it].z), "=f"(pre_f[it].w)
                             : "l"(fbase + (long long)rc * fld + col)
                             : "memory");
              if (kDgelu)
                asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];"
                             : "=r"(pre_h[it].x), "=r"(pre_h[it].y)
                             : "l"(p.aux_in + (long long)rc * p.ld_aux + col)
                             : "memory");
            }
            // one fake use of every loaded register: all 8 loads must have been issued before this point
            if (kRes || kLoss)
              asm volatile("" ::"f"(pre_f[0].x), "f"(pre_f[1].x), "f"(pre_f[2].x), "f"(pre_f[3].x), "f"(pre_f[4].x),
                           "f"(pre_f[5].x), "f"(pre_f[6].x), "f"(pre_f[7].x));
            if (kDgelu)
              asm volatile("" ::"r"(pre_h[0].x), "r"(pre_h[1].x), "r"(pre_h[2].x), "r"(pre_h[3].x), "r"(pre_h[4].x),
                           "r"(pre_h[5].x), "r"(pre_h[6].x), "r"(pre_h[7].x));
          } else if (G) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = rbase + it * 4 + rsub;
              pre_f[it] = make_float4(0.f, 0.f, 0.f, 0.f);
              pre_h[it] = make_uint2(0u, 0u);
              if (r < p.M) {
                if (kRes) {
                  const long long rr = p.res_idx ? (long long)__ldg(p.res_idx + r) : (long long)r;
                  pre_f[it] = __ldg(reinterpret_cast<const float4*>(p.res + rr * p.ldr + col));
                } else if (kLoss) {
                  pre_f[it] = __ldg(reinterpret_cast<const float4*>(p.target + (long long)r * p.ldt + col));
                }
                if (kDgelu) pre_h[it] = __ldg(reinterpret_cast<const uint2*>(p.aux_in + (long long)r * p.ld_aux + col));
              }
            }
          }
          float gcs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rl = it * 4 + rsub;
            const int r = rbase + rl;
            if (r >= p.M) continue;
            const float4 t = tv[it];
            float x[4] = {t.x * alpha, t.y * alpha, t.z * alpha, t.w * alpha};
            long long R = r;
            if (kSeg && p.out_seg > 0)
              R = (long long)(r / p.out_seg) * p.out_seg_stride + (r % p.out_seg) + p.out_seg_off;
            if (kSplit) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.out_f32 + R * p.ldo + col),
                           "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3])
                           : "memory");
              continue;
            }
            x[0] += bias4.x; x[1] += bias4.y; x[2] += bias4.z; x[3] += bias4.w;
            if (kGelu) {
              uint2 pk;
              pk.x = pack_bf16x2(x[0], x[1]);
              pk.y = pack_bf16x2(x[2], x[3]);
              const float xb[4] = {__uint_as_float(pk.x << 16), __uint_as_float(pk.x & 0xffff0000u),
                                   __uint_as_float(pk.y << 16), __uint_as_float(pk.y & 0xffff0000u)};
              if (p.aux_out) {
                uint2 gk;
                gk.x = pack_bf16x2(gelu_erf_grad(xb[0]), gelu_erf_grad(xb[1]));
                gk.y = pack_bf16x2(gelu_erf_grad(xb[2]), gelu_erf_grad(xb[3]));
                *reinterpret_cast<uint2*>(p.aux_out + (long long)r * p.ld_aux + col) = gk;
              }
              x[0] = gelu_erf(xb[0]);
              x[1] = gelu_erf(xb[1]);
              x[2] = gelu_erf(xb[2]);
              x[3] = gelu_erf(xb[3]);
            } else if (kDgelu) {
              const uint2 pk = pre_h[it];
              x[0] *= __uint_as_float(pk.x << 16);
              x[1] *= __uint_as_float(pk.x & 0xffff0000u);
              x[2] *= __uint_as_float(pk.y << 16);
              x[3] *= __uint_as_float(pk.y & 0xffff0000u);
            }
            if (kRes) {
              const float4 rv = pre_f[it];
              x[0] += rv.x; x[1] += rv.y; x[2] += rv.z; x[3] += rv.w;
            }
            if (kLoss) {
              if (p.logits_out) {
                uint2 pk;
                pk.x = pack_bf16x2(x[0], x[1]);
                pk.y = pack_bf16x2(x[2], x[3]);
                *reinterpret_cast<uint2*>(p.logits_out + R * p.ldo + col) = pk;
              }
              const float4 tv = pre_f[it];
              x[0] -= tv.x; x[1] -= tv.y; x[2] -= tv.z; x[3] -= tv.w;
              lsum = fmaf(x[0], x[0], lsum); lsum = fmaf(x[1], x[1], lsum);
              lsum = fmaf(x[2], x[2], lsum); lsum = fmaf(x[3], x[3], lsum);
            }
            if (kOutF32) *reinterpret_cast<float4*>(p.out_f32 + R * p.ldo + col) = make_float4(x[0], x[1], x[2], x[3]);
            if (kOutBf16) {
              uint2 pk;
              pk.x = pack_bf16x2(x[0], x[1]);
              pk.y = pack_bf16x2(x[2], x[3]);
              *reinterpret_cast<uint2*>(p.out_bf16 + R * p.ldo + col) = pk;
            }
            gcs[0] += x[0]; gcs[1] += x[1]; gcs[2] += x[2]; gcs[3] += x[3];
          }
          if (p.colsum && !kSplit && !kLoss)  // per-lane partial sums: a handful of atomics on the ragged edge tiles
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.colsum + col), "f"(gcs[0]),
                         "f"(gcs[1]), "f"(gcs[2]), "f"(gcs[3])
                         : "memory");
        }
        __syncwarp();
      }
      if (kLoss && p.loss_partial) {
        lsum = warp_sum(lsum);
        // slot = (128-row tile index) * tiles_n + column tile (pair mode: a CTA whose rows are all past M has none)
        if (lane == 0 && m0 < p.M)
          p.loss_partial[((long long)(m0 / BM) * p.tiles_n + tile % p.tiles_n) * kEpiWarps + e] = lsum;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer may still read this CTA's B half / signal its barriers until here
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// k-split resolution shared by the host dispatcher (it decides the epilogue variant) and the launcher
static inline void resolve_k_splits(const bvc_gemm_args* a, int bn, int pair, int* kb_per_split, int* k_splits) {
  const int tile_m = pair ? 2 * BM : BM;
  const int tiles_m = (a->M + tile_m - 1) / tile_m, tiles_n = (a->N + bn - 1) / bn;
  const int kb_total = (a->K + BK - 1) / BK;
  int ks = a->k_splits;
  const int sms = pair ? num_sms() / 2 : num_sms();  // persistent workers: CTAs, or CTA pairs
  if (ks <= 0) {
    const long long tiles = (long long)tiles_m * tiles_n;
    ks = (int)((2LL * sms + tiles - 1) / tiles);  // ~2 waves of work items
    if (ks < 1) ks = 1;
    if (ks > kb_total / 4) ks = kb_total / 4 > 0 ? kb_total / 4 : 1;  // >= 4 k-blocks per split
  }
  if (ks > kb_total) ks = kb_total;
  *kb_per_split = (kb_total + ks - 1) / ks;
  *k_splits = (kb_total + *kb_per_split - 1) / *kb_per_split;
}

template <int BN, int A_MN, int B_MN, int EPI, int PAIR = 0>
static int launch_gemm(const bvc_gemm_args* a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, PAIR>;
  static_assert(Cfg::kSmemBytes <= 232448, "stage ring does not fit the 227 KB of shared memory");
  // once per instantiation, thread-safe (C++11 static initialisation): the forward thread and autograd's backward
  // thread both launch
  static const int max_workers = []() -> int {  // co-resident CTAs (pair mode: clusters); 0 = setup failed
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, A_MN, B_MN, EPI, PAIR>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      fprintf(stderr, "bvc: cudaFuncSetAttribute(gemm) failed: %s\n", cudaGetErrorString(e));
      return 0;
    }
    int mw = PAIR ? num_sms() / 2 : num_sms();
    if (PAIR) {
      // how many CTA pairs the device can hold at once (GPCs with an odd number of usable SMs lose one)
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(2 * (num_sms() / 2));
      qc.blockDim = dim3(kGemmThreads);
      qc.dynamicSmemBytes = Cfg::kSmemBytes;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      qc.attrs = qa;
      qc.numAttrs = 1;
      int n_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&n_clusters, gemm_kernel<BN, A_MN, B_MN, EPI, PAIR>, &qc) == cudaSuccess &&
          n_clusters > 0 && n_clusters < mw)
        mw = n_clusters;
      (void)cudaGetLastError();
    }
    return mw;
  }();
  if (max_workers <= 0) return BVC_ERR_LAUNCH;
  CUtensorMap ta, tb;
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (A_MN == 0) {
      dims[0] = (uint64_t)a->K; dims[1] = (uint64_t)a->M; box[0] = BK; box[1] = BM;
    } else {
      dims[0] = (uint64_t)a->M; dims[1] = (uint64_t)a->K; box[0] = 64; box[1] = BK;
    }
    strides[0] = (uint64_t)a->lda * 2;
    int rc = make_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (B_MN == 0) {
      dims[0] = (uint64_t)a->K; dims[1] = (uint64_t)a->N; box[0] = BK; box[1] = Cfg::kBRows;
    } else {
      dims[0] = (uint64_t)a->N; dims[1] = (uint64_t)a->K; box[0] = 64; box[1] = BK;
    }
    strides[0] = (uint64_t)a->ldb * 2;
    rc = make_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->b, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  CUtensorMap tp = ta;  // placeholder when unused
  int has_pre = 0;
  {
    const void* base = nullptr;
    uint64_t ld_bytes = 0;
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if (EPI == EPI_RES && a->res_idx == nullptr) { base = a->res; ld_bytes = (uint64_t)a->ldr * 4; }
    if (EPI == EPI_LOSS) { base = a->target; ld_bytes = (uint64_t)a->ldt * 4; }
    if (EPI == EPI_DGELU) { base = a->aux_in; ld_bytes = (uint64_t)a->ld_aux * 2; dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; }
    if (base != nullptr && (((uintptr_t)base) & 15) == 0 && ld_bytes % 16 == 0) {
      const uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->M};
      const uint64_t strides[1] = {ld_bytes};
      const uint32_t box[2] = {(uint32_t)BN, (uint32_t)BM};
      if (make_tmap(&tp, dt, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE) == BVC_OK) has_pre = 1;
    }
  }
  GemmParams p;
  p.has_pre = has_pre;
  p.M = a->M; p.N = a->N; p.K = a->K;
  constexpr int kTileM = PAIR ? 2 * BM : BM;
  p.tiles_m = (a->M + kTileM - 1) / kTileM;
  p.tiles_n = (a->N + BN - 1) / BN;
  p.kb_total = (a->K + BK - 1) / BK;
  resolve_k_splits(a, BN, PAIR, &p.kb_per_split, &p.k_splits);
  p.out_f32 = a->out_f32; p.out_bf16 = (bf16*)a->out_bf16; p.ldo = a->ldo;
  p.out_seg = a->out_seg; p.out_seg_stride = a->out_seg_stride; p.out_seg_off = a->out_seg_off;
  p.alpha = a->alpha_host; p.alpha_dev = a->alpha_dev; p.bias = a->bias; p.act = a->act;
  p.aux_out = (bf16*)a->aux_out; p.aux_in = (const bf16*)a->aux_in; p.ld_aux = a->ld_aux;
  p.res = a->res; p.ldr = a->ldr; p.res_idx = a->res_idx;
  p.target = a->target; p.ldt = a->ldt; p.loss_partial = a->loss_partial; p.logits_out = (bf16*)a->logits_out;
  p.colsum = a->colsum;
  if (p.k_splits > 1) {
    BVC_CHECK_ARG(a->out_f32 != nullptr && a->out_bf16 == nullptr && a->bias == nullptr && a->act == 0 &&
                  a->res == nullptr && a->target == nullptr);
  }
  const long long total = (long long)p.tiles_m * p.tiles_n * p.k_splits;
  int cap = PAIR ? num_sms() / 2 : num_sms();
  if (cap > max_workers) cap = max_workers;
  const int workers = (int)(total < cap ? total : cap);
  if constexpr (PAIR != 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * workers);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, gemm_kernel<BN, A_MN, B_MN, EPI, PAIR>, ta, tb, tp, p) != cudaSuccess) {
      fprintf(stderr, "bvc: pair GEMM launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      return BVC_ERR_LAUNCH;
    }
  } else {
    gemm_kernel<BN, A_MN, B_MN, EPI, PAIR><<<workers, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tp, p);
  }
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

// every (operand-major, epilogue) combination the model uses gets its own lean kernel; anything else runs the
// generic (all-runtime-flags) epilogue.  PAIR variants exist for the tile widths whose B half is a whole number of
// swizzle atoms (BN = 128, 256; BN = 192 only with a K-major B).
template <int BN, int PAIR>
static int gemm_dispatch_bn_pair(const bvc_gemm_args* a, int epi, cudaStream_t s) {
  const int am = a->a_mn_major, bm = a->b_mn_major;
  if (am == 0 && bm == 0) {
    switch (epi) {
      case EPI_PLAIN: return launch_gemm<BN, 0, 0, EPI_PLAIN, PAIR>(a, s);
      case EPI_GELU: return launch_gemm<BN, 0, 0, EPI_GELU, PAIR>(a, s);
      case EPI_RES: return launch_gemm<BN, 0, 0, EPI_RES, PAIR>(a, s);
      case EPI_LOSS: return launch_gemm<BN, 0, 0, EPI_LOSS, PAIR>(a, s);
      default: return launch_gemm<BN, 0, 0, EPI_GENERIC, PAIR>(a, s);
    }
  }
  if constexpr (!PAIR || (BN / 2) % 64 == 0) {
    if (am == 0 && bm == 1) {
      switch (epi) {
        case EPI_PLAIN: return launch_gemm<BN, 0, 1, EPI_PLAIN, PAIR>(a, s);
        case EPI_DGELU: return launch_gemm<BN, 0, 1, EPI_DGELU, PAIR>(a, s);
        default: return launch_gemm<BN, 0, 1, EPI_GENERIC, PAIR>(a, s);
      }
    }
    if (am == 1 && bm == 1) {
      if (epi == EPI_SPLITK) return launch_gemm<BN, 1, 1, EPI_SPLITK, PAIR>(a, s);
      return launch_gemm<BN, 1, 1, EPI_GENERIC, PAIR>(a, s);
    }
  } else {
    if (bm == 1) return BVC_ERR_ARG;  // resolved away by the host dispatcher (gemm.cu: pair_supported)
  }
  return launch_gemm<BN, 1, 0, EPI_GENERIC, PAIR>(a, s);
}

template <int BN>
static int gemm_dispatch_bn(const bvc_gemm_args* a, int epi, cudaStream_t s) {
  return gemm_dispatch_bn_pair<BN, 0>(a, epi, s);
}

}  // namespace bvc
