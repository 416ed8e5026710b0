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
constexpr int kGemmThreads = 64 + kEpiWarps * 32;

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
};

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 192) ? 5 : (BN == 128) ? 6 : 8;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
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
      mbar_init(&tmem_empty[i], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_work = p.tiles_m * p.tiles_n * p.k_splits;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w % p.k_splits;
        const int tile = w / p.k_splits;
        const int m0 = (tile / p.tiles_n) * BM;
        const int n0 = (tile % p.tiles_n) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (A_MN == 0) {
            tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &tma_a, &full_bar[stage], m0 + j * 64, kb * BK);
          }
          if (B_MN == 0) {
            tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tma_b, &full_bar[stage], n0 + j * 64, kb * BK);
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BN, A_MN, B_MN, BM);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int split = w % p.k_splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? umma_smem_desc(a_addr + k * 2048, 1024, 8192)
                                     : umma_smem_desc(a_addr + k * 32, 1024, 16);
            const uint64_t db = B_MN ? umma_smem_desc(b_addr + k * 2048, 1024, 8192)
                                     : umma_smem_desc(b_addr + k * 32, 1024, 16);
            umma_bf16_ss(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;          // TMEM lane quadrant this warp may read
    const int half = e >> 2;         // which half of the BN columns
    constexpr int kColsPerWarp = BN / 2;
    const float alpha = p.alpha * (p.alpha_dev ? __ldg(p.alpha_dev) : 1.0f);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int tile = w / p.k_splits;
      const int m0 = (tile / p.tiles_n) * BM;
      const int n0 = (tile % p.tiles_n) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int r = m0 + q * 32 + lane;
      const bool row_ok = r < p.M;
      long long R = r;
      if (p.out_seg > 0) R = (long long)(r / p.out_seg) * p.out_seg_stride + (r % p.out_seg) + p.out_seg_off;
      const long long rr = (p.res && row_ok) ? (p.res_idx ? (long long)__ldg(p.res_idx + r) : (long long)r) : 0;
      float lsum = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < kColsPerWarp; cc += 32) {
        const int c_tile = half * kColsPerWarp + cc;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c_tile), v);
        tmem_ld_wait();
        if (!row_ok) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c = n0 + c_tile + g * 8;
          if (c >= p.N) continue;
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[g * 8 + j]) * alpha;
          if (p.k_splits > 1) {
            float* o = p.out_f32 + R * p.ldo + c;
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(o + j, x[j]);
            continue;
          }
          if (p.bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c) + 1);
            x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
            x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
          }
          if (p.act == 1) {
            uint4 pk;
            pk.x = pack_bf16x2(x[0], x[1]); pk.y = pack_bf16x2(x[2], x[3]);
            pk.z = pack_bf16x2(x[4], x[5]); pk.w = pack_bf16x2(x[6], x[7]);
            if (p.aux_out) *reinterpret_cast<uint4*>(p.aux_out + (long long)r * p.ld_aux + c) = pk;
            const uint32_t pw[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              x[2 * j] = gelu_erf(__uint_as_float(pw[j] << 16));
              x[2 * j + 1] = gelu_erf(__uint_as_float(pw[j] & 0xffff0000u));
            }
          } else if (p.act == 2) {
            const uint4 pk = __ldg(reinterpret_cast<const uint4*>(p.aux_in + (long long)r * p.ld_aux + c));
            const uint32_t pw[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              x[2 * j] *= gelu_erf_grad(__uint_as_float(pw[j] << 16));
              x[2 * j + 1] *= gelu_erf_grad(__uint_as_float(pw[j] & 0xffff0000u));
            }
          }
          if (p.res) {
            const float4* rp = reinterpret_cast<const float4*>(p.res + rr * p.ldr + c);
            const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            x[0] += r0.x; x[1] += r0.y; x[2] += r0.z; x[3] += r0.w;
            x[4] += r1.x; x[5] += r1.y; x[6] += r1.z; x[7] += r1.w;
          }
          if (p.target) {
            if (p.logits_out) {
              uint4 pk;
              pk.x = pack_bf16x2(x[0], x[1]); pk.y = pack_bf16x2(x[2], x[3]);
              pk.z = pack_bf16x2(x[4], x[5]); pk.w = pack_bf16x2(x[6], x[7]);
              *reinterpret_cast<uint4*>(p.logits_out + R * p.ldo + c) = pk;
            }
            const float4* tp = reinterpret_cast<const float4*>(p.target + (long long)r * p.ldt + c);
            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
            x[0] -= t0.x; x[1] -= t0.y; x[2] -= t0.z; x[3] -= t0.w;
            x[4] -= t1.x; x[5] -= t1.y; x[6] -= t1.z; x[7] -= t1.w;
#pragma unroll
            for (int j = 0; j < 8; ++j) lsum = fmaf(x[j], x[j], lsum);
          }
          if (p.out_f32) {
            float4* o = reinterpret_cast<float4*>(p.out_f32 + R * p.ldo + c);
            o[0] = make_float4(x[0], x[1], x[2], x[3]);
            o[1] = make_float4(x[4], x[5], x[6], x[7]);
          }
          if (p.out_bf16) {
            uint4 pk;
            pk.x = pack_bf16x2(x[0], x[1]); pk.y = pack_bf16x2(x[2], x[3]);
            pk.z = pack_bf16x2(x[4], x[5]); pk.w = pack_bf16x2(x[6], x[7]);
            *reinterpret_cast<uint4*>(p.out_bf16 + R * p.ldo + c) = pk;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (p.loss_partial) {
        lsum = warp_sum(lsum);
        if (lane == 0) p.loss_partial[(long long)tile * kEpiWarps + e] = lsum;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
static int pick_block_n(int M, int N) {
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

template <int BN, int A_MN, int B_MN>
static int launch_gemm(const bvc_gemm_args* a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      fprintf(stderr, "bvc: cudaFuncSetAttribute(gemm) failed: %s\n", cudaGetErrorString(e));
      return BVC_ERR_LAUNCH;
    }
    attr_done = true;
  }
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
      dims[0] = (uint64_t)a->K; dims[1] = (uint64_t)a->N; box[0] = BK; box[1] = BN;
    } else {
      dims[0] = (uint64_t)a->N; dims[1] = (uint64_t)a->K; box[0] = 64; box[1] = BK;
    }
    strides[0] = (uint64_t)a->ldb * 2;
    rc = make_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->b, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.tiles_m = (a->M + BM - 1) / BM;
  p.tiles_n = (a->N + BN - 1) / BN;
  p.kb_total = (a->K + BK - 1) / BK;
  int ks = a->k_splits;
  const int sms = num_sms();
  if (ks <= 0) {
    const long long tiles = (long long)p.tiles_m * p.tiles_n;
    ks = (int)((2LL * sms + tiles - 1) / tiles);  // ~2 waves of work items
    if (ks < 1) ks = 1;
    if (ks > p.kb_total / 4) ks = p.kb_total / 4 > 0 ? p.kb_total / 4 : 1;  // >= 4 k-blocks per split
  }
  if (ks > p.kb_total) ks = p.kb_total;
  p.kb_per_split = (p.kb_total + ks - 1) / ks;
  p.k_splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.out_f32 = a->out_f32; p.out_bf16 = (bf16*)a->out_bf16; p.ldo = a->ldo;
  p.out_seg = a->out_seg; p.out_seg_stride = a->out_seg_stride; p.out_seg_off = a->out_seg_off;
  p.alpha = a->alpha_host; p.alpha_dev = a->alpha_dev; p.bias = a->bias; p.act = a->act;
  p.aux_out = (bf16*)a->aux_out; p.aux_in = (const bf16*)a->aux_in; p.ld_aux = a->ld_aux;
  p.res = a->res; p.ldr = a->ldr; p.res_idx = a->res_idx;
  p.target = a->target; p.ldt = a->ldt; p.loss_partial = a->loss_partial; p.logits_out = (bf16*)a->logits_out;
  if (p.k_splits > 1) {
    BVC_CHECK_ARG(a->out_f32 != nullptr && a->out_bf16 == nullptr && a->bias == nullptr && a->act == 0 &&
                  a->res == nullptr && a->target == nullptr);
  }
  const long long total = (long long)p.tiles_m * p.tiles_n * p.k_splits;
  const int grid = (int)(total < sms ? total : sms);
  gemm_kernel<BN, A_MN, B_MN><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

template <int BN>
static int dispatch_major(const bvc_gemm_args* a, cudaStream_t s) {
  if (a->a_mn_major == 0 && a->b_mn_major == 0) return launch_gemm<BN, 0, 0>(a, s);
  if (a->a_mn_major == 0 && a->b_mn_major == 1) return launch_gemm<BN, 0, 1>(a, s);
  if (a->a_mn_major == 1 && a->b_mn_major == 0) return launch_gemm<BN, 1, 0>(a, s);
  return launch_gemm<BN, 1, 1>(a, s);
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

static int resolve_block_n(int M, int N, int block_n) {
  if (block_n == 64 || block_n == 128 || block_n == 192 || block_n == 256) return block_n;
  return bvc::pick_block_n(M, N);
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
  BVC_CHECK_ARG(a->out_f32 != nullptr || a->out_bf16 != nullptr);
  BVC_CHECK_ARG((((uintptr_t)a->a) & 15) == 0 && (((uintptr_t)a->b) & 15) == 0);
  BVC_CHECK_ARG(a->act == 0 || a->act == 1 || (a->act == 2 && a->aux_in != nullptr));
  BVC_CHECK_ARG(a->act == 0 || a->ld_aux % 8 == 0);
  BVC_CHECK_ARG(a->res == nullptr || a->ldr % 4 == 0);
  BVC_CHECK_ARG(a->target == nullptr || (a->ldt % 4 == 0 && a->loss_partial != nullptr));
  BVC_CHECK_ARG(a->a_mn_major == 0 ? a->lda >= a->K : a->lda >= a->M);
  BVC_CHECK_ARG(a->b_mn_major == 0 ? a->ldb >= a->K : a->ldb >= a->N);
  cudaStream_t s = (cudaStream_t)stream;
  switch (resolve_block_n(a->M, a->N, a->block_n)) {
    case 256: return bvc::dispatch_major<256>(a, s);
    case 192: return bvc::dispatch_major<192>(a, s);
    case 128: return bvc::dispatch_major<128>(a, s);
    default: return bvc::dispatch_major<64>(a, s);
  }
}

extern "C" int bvc_loss_finalize(const float* partials, int64_t n, double numel, const int32_t* status, float* loss,
                                 void* stream) {
  BVC_CHECK_ARG(partials != nullptr && loss != nullptr && n > 0 && numel > 0);
  bvc::loss_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, (long long)n, 1.0 / numel, status, loss);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}
