// attn.cu -- multi-head self-attention (head_dim 64, no mask, no dropout) on tcgen05, forward and backward.
// Replaces HF VideoMAESelfAttention (HF:236-266) + sdpa/eager attention (HF:181-206) and its autograd backward.
//
// Layouts: qkv bf16 [B, S, 3, H, 64] (the fused-QKV GEMM output), out/dout bf16 [B, S, H*64], lse/delta fp32 [B,H,S].
// All operand tiles are 128 rows x 64 bf16 (128-byte rows, 128B swizzle) fetched by one 4-D TMA box; rows past the
// end of the sequence are zero-filled by TMA, columns past the end are masked in registers.
//
// Forward (grid = q-tile x head x clip, 2 CTAs/SM so one CTA's softmax overlaps the other's MMAs):
//   warp 0 TMA (Q once, K/V ring), warp 1 MMA issuer, warps 2-5 softmax (one query row per thread).
//   S = Q K^T -> TMEM (128 cols); softmax threads read S, exp2 with a lazily updated running max (rescale O in TMEM
//   only when the max grows by > 8 in log2 units), write P (bf16) into swizzled smem; O += P V accumulates in TMEM.
// Backward (two passes, no atomics, deterministic):
//   MODE_KV: CTA owns a K/V tile, streams Q/dO tiles:  S, dP = dO V^T, P = exp2(S c - lse), dS = P (dP - delta) scale,
//            dV += P^T dO, dK += dS^T Q   (P / dS tiles in smem are read MN-major: no transposes anywhere)
//   MODE_Q : CTA owns a Q/dO tile, streams K/V tiles:  same S, dP, dS;  dQ += dS K.
#include "../../include/bvc.h"
#include "bvc_host.h"
#include "bvc_ptx.cuh"

namespace bvc {

constexpr int kTile = 128;          // rows per Q / KV tile
constexpr int kTileBytes = 16384;   // 128 x 64 bf16
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2 / FMUL2): halves the issue slots of the softmax-side math, which is
// what bounds these kernels at head_dim 64 (ncu: issue-active 43 %, tensor pipe 25 %)
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// K-major operand tile [rows][64]: k-step (16 elements) = +32 bytes inside the swizzle atom
__device__ __forceinline__ uint64_t desc_k(uint32_t base, int kstep) { return umma_smem_desc(base + kstep * 32, 1024, 16); }
// MN-major operand tile [k rows][64 mn]: k-step (16 rows) = +2048 bytes; lbo = distance between 64-wide mn blocks
__device__ __forceinline__ uint64_t desc_mn(uint32_t base, int kstep, uint32_t lbo) {
  return umma_smem_desc(base + kstep * 2048, 1024, lbo);
}
// byte offset of 8 consecutive bf16 (one 16-byte chunk) of element (row, col8*8) in a [128][128] bf16 tile stored as
// two [128 rows][64] 128B-swizzled sub-tiles (the layout TMA would have produced)
__device__ __forceinline__ uint32_t ptile_chunk_off(int row, int chunk16 /*0..15*/) {
  return (uint32_t)((chunk16 >> 3) * kTileBytes + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4));
}

// ================================================================================================ forward
constexpr int kFwdThreads = 192;
constexpr int kFwdSmem = kTileBytes * (1 + 2 + 2) + 2 * kTileBytes + 128;  // Q, K[2], V[2], P(32 KB), barriers

__global__ void __launch_bounds__(kFwdThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, bf16* __restrict__ out, float* __restrict__ lse, int S,
                int H, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;          // 2 stages
  uint8_t* sV = smem + 3 * kTileBytes;      // 2 stages
  uint8_t* sP = smem + 5 * kTileBytes;      // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]   K and V have separate rings: a K stage is released as soon as its
  uint64_t* k_empty = bars + 3;   // [2]   S = Q K^T MMA has run (long before the P V MMA of the same tile), which
  uint64_t* v_full = bars + 5;    // [2]   gives the next K tile's TMA two softmax periods of lead time
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* s_free = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* o_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (S + kTile - 1) / kTile;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tm_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_4d(sQ, &tm_qkv, q_full, 0, h, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t par = ((uint32_t)(j >> 1) & 1u) ^ 1u;
        mbar_wait(&k_empty[st], par);
        mbar_expect_tx(&k_full[st], kTileBytes);
        tma_load_4d(sK + st * kTileBytes, &tm_qkv, &k_full[st], 0, H + h, j * kTile, b);
        mbar_wait(&v_empty[st], par);
        mbar_expect_tx(&v_full[st], kTileBytes);
        tma_load_4d(sV + st * kTileBytes, &tm_qkv, &v_full[st], 0, 2 * H + h, j * kTile, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 0, 0, 128);
      constexpr uint32_t idesc_o = umma_idesc_bf16(64, 0, 1, 128);
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tS, desc_k(aQ, k), desc_k(smem_u32(sK), k), idesc_s, k > 0);
      umma_commit(s_full);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          const int st = (j + 1) & 1;
          mbar_wait(&k_full[st], (uint32_t)((j + 1) >> 1) & 1u);
          mbar_wait(s_free, (uint32_t)j & 1u);
          tc_fence_after();
          const uint32_t aK = smem_u32(sK + st * kTileBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tS, desc_k(aQ, k), desc_k(aK, k), idesc_s, k > 0);
          umma_commit(s_full);
          umma_commit(&k_empty[st]);
        }
        mbar_wait(&v_full[j & 1], (uint32_t)(j >> 1) & 1u);
        mbar_wait(p_full, (uint32_t)j & 1u);
        tc_fence_after();
        const uint32_t aV = smem_u32(sV + (j & 1) * kTileBytes);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_ss(tO, desc_k(aP + (k >> 2) * kTileBytes, k & 3), desc_mn(aV, k, 8192), idesc_o, (j > 0 || k > 0));
        umma_commit(o_full);
        umma_commit(&v_empty[j & 1]);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      const int valid = S - j * kTile;  // columns >= valid are past the end of the sequence
      // pass 1: row maximum of the raw scores, two 32-column chunks of S in registers at a time (keeping all 128
      // scores live spilled ~340 B per thread at the 168-register cap that 2 CTAs/SM imposes)
      float mx = -INFINITY;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t sv[2][32];
        tmem_ld_32x32b_x32(tS + lane_base + hh * 64, sv[0]);
        tmem_ld_32x32b_x32(tS + lane_base + hh * 64 + 32, sv[1]);
        tmem_ld_wait();
        if (valid >= kTile) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; i += 2) mx = fmax3(mx, __uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1]));
        } else {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (hh * 64 + c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(sv[c][i]));
        }
      }
      mx *= scale_log2;  // raw scores -> log2 domain (scale_log2 > 0)
      const bool need = mx > m_used + 8.0f;
      float alpha = 1.0f;
      if (need) {
        alpha = exp2f(m_used - mx);  // 0 on the first tile
        m_used = mx;
      }
      if (j > 0) {
        mbar_wait(o_full, (uint32_t)(j - 1) & 1u);  // PV_{j-1} done: P smem reusable, O stable
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32b_x32(tO + lane_base + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x32b_x32(tO + lane_base + c * 32, ov);
          }
          tmem_st_wait();
        }
      }
      l *= alpha;
      // pass 2: P = exp2(S * scale_log2 - m), row sum, bf16 P into the swizzled A-operand tile
      const uint64_t c2 = pack2(scale_log2, scale_log2), m2 = pack2(-m_used, -m_used);
      uint64_t lsum2 = pack2(0.f, 0.f);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t sv[2][32];
        tmem_ld_32x32b_x32(tS + lane_base + hh * 64, sv[0]);
        tmem_ld_32x32b_x32(tS + lane_base + hh * 64 + 32, sv[1]);
        tmem_ld_wait();
        if (hh == 1) {  // S fully consumed: the MMA warp may overwrite it with the next tile's scores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float p[8];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              const int jj = g * 8 + i;
              const uint64_t x2 = ffma2(pack2(__uint_as_float(sv[c][jj]), __uint_as_float(sv[c][jj + 1])), c2, m2);
              float a, b;
              unpack2(x2, a, b);
              p[i] = exp2f(a);
              p[i + 1] = exp2f(b);
              if (valid < kTile) {  // warp-uniform: last K/V tile only
                const int col = hh * 64 + c * 32 + jj;
                if (col >= valid) p[i] = 0.f;
                if (col + 1 >= valid) p[i + 1] = 0.f;
              }
              lsum2 = fadd2(lsum2, pack2(p[i], p[i + 1]));
            }
            uint4 pk;
            pk.x = pack_bf16x2(p[0], p[1]); pk.y = pack_bf16x2(p[2], p[3]);
            pk.z = pack_bf16x2(p[4], p[5]); pk.w = pack_bf16x2(p[6], p[7]);
            *reinterpret_cast<uint4*>(sP + ptile_chunk_off(row, hh * 8 + c * 4 + g)) = pk;
          }
      }
      float lsum, lsum_hi;
      unpack2(lsum2, lsum, lsum_hi);
      lsum += lsum_hi;
      l += lsum;
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_full, (uint32_t)(n_kv - 1) & 1u);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int qrow = q0 + row;
    bf16* orow = out + ((long long)(b * (long long)S + qrow) * H + h) * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t ov[32];
      tmem_ld_32x32b_x32(tO + lane_base + c * 32, ov);
      tmem_ld_wait();
      if (qrow < S) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(ov[g * 8 + 0]) * inv_l, __uint_as_float(ov[g * 8 + 1]) * inv_l);
          pk.y = pack_bf16x2(__uint_as_float(ov[g * 8 + 2]) * inv_l, __uint_as_float(ov[g * 8 + 3]) * inv_l);
          pk.z = pack_bf16x2(__uint_as_float(ov[g * 8 + 4]) * inv_l, __uint_as_float(ov[g * 8 + 5]) * inv_l);
          pk.w = pack_bf16x2(__uint_as_float(ov[g * 8 + 6]) * inv_l, __uint_as_float(ov[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = pk;
        }
      }
    }
    if (qrow < S) lse[((long long)b * H + h) * S + qrow] = (m_used + log2f(l)) * 0.6931471805599453f;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ================================================================================================ backward
// delta[b,h,s] = sum_d dO[b,s,h,d] * O[b,s,h,d]   (8 lanes per (b,s,h) row of 64, 8 rows per warp step, loads batched)
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout,
                                                         float* __restrict__ delta, long long rows, int S, int H) {
  const int lane = threadIdx.x & 31;
  const int sub = lane >> 3, part = lane & 7;
  const long long warp_id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * (blockDim.x >> 5) * 8;
  for (long long r0 = warp_id * 8; r0 < rows; r0 += stride) {
    uint4 a[2], g[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long r = min(r0 + u * 4 + sub, rows - 1);
      a[u] = ldv_u4(o + r * 64 + part * 8);
      g[u] = ldv_u4(dout + r * 64 + part * 8);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t aw[4] = {a[u].x, a[u].y, a[u].z, a[u].w}, gw[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
      float sacc = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sacc += __uint_as_float(aw[j] << 16) * __uint_as_float(gw[j] << 16) +
                __uint_as_float(aw[j] & 0xffff0000u) * __uint_as_float(gw[j] & 0xffff0000u);
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
      const long long r = r0 + u * 4 + sub;
      if (part == 0 && r < rows) {
        const int hh = (int)(r % H);
        const long long bs = r / H;
        const int ss = (int)(bs % S);
        const long long bb = bs / S;
        delta[(bb * H + hh) * S + ss] = sacc;
      }
    }
  }
}

constexpr int kBwdThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 compute
// resident pair (2 tiles) + stream ring (4 stages x 2 tiles: the streamed tiles' TMA latency was the critical path
// with 2 stages -- ncu: 30 % of the compute warps' samples waiting for S/dP) + P (32 KB) + dS (32 KB) + barriers
constexpr int kBwdStages = 4;
constexpr int kBwdSmem = kTileBytes * (2 + 2 * kBwdStages) + 4 * kTileBytes + 256;

template <int MODE_KV>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dqkv, int S, int H,
                float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sR0 = smem;                     // MODE_KV: K_j    | MODE_Q: Q_i
  uint8_t* sR1 = smem + kTileBytes;        // MODE_KV: V_j    | MODE_Q: dO_i
  uint8_t* sX = smem + 2 * kTileBytes;                     // [stages] MODE_KV: Q_i  | MODE_Q: K_j
  uint8_t* sY = smem + (2 + kBwdStages) * kTileBytes;      // [stages] MODE_KV: dO_i | MODE_Q: V_j
  uint8_t* sP = smem + (2 + 2 * kBwdStages) * kTileBytes;  // 32 KB (MODE_KV only)
  uint8_t* sD = sP + 2 * kTileBytes;                       // 32 KB dS
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 2 * kTileBytes);
  uint64_t* r_full = bars + 0;
  uint64_t* st_full = bars + 1;                  // [stages]
  uint64_t* st_empty = bars + 1 + kBwdStages;    // [stages]
  uint64_t* sdp_full = bars + 1 + 2 * kBwdStages;
  uint64_t* sdp_free = sdp_full + 1;
  uint64_t* pds_full = sdp_full + 2;
  uint64_t* pds_free = sdp_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sdp_full + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int own0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
  const int n_it = (S + kTile - 1) / kTile;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(r_full, 1);
    for (int i = 0; i < kBwdStages; ++i) {
      mbar_init(&st_full[i], 1);
      mbar_init(&st_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 8);
    mbar_init(pds_full, 8);
    mbar_init(pds_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tA0 = tmem_base + 256, tA1 = tmem_base + 320;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(r_full, 2 * kTileBytes);
      if (MODE_KV) {
        tma_load_4d(sR0, &tm_qkv, r_full, 0, H + h, own0, b);
        tma_load_4d(sR1, &tm_qkv, r_full, 0, 2 * H + h, own0, b);
      } else {
        tma_load_4d(sR0, &tm_qkv, r_full, 0, h, own0, b);
        tma_load_4d(sR1, &tm_do, r_full, 0, h, own0, b);
      }
      for (int i = 0; i < n_it; ++i) {
        const int st = i % kBwdStages;
        mbar_wait(&st_empty[st], ((uint32_t)(i / kBwdStages) & 1u) ^ 1u);
        mbar_expect_tx(&st_full[st], 2 * kTileBytes);
        if (MODE_KV) {
          tma_load_4d(sX + st * kTileBytes, &tm_qkv, &st_full[st], 0, h, i * kTile, b);
          tma_load_4d(sY + st * kTileBytes, &tm_do, &st_full[st], 0, h, i * kTile, b);
        } else {
          tma_load_4d(sX + st * kTileBytes, &tm_qkv, &st_full[st], 0, H + h, i * kTile, b);
          tma_load_4d(sY + st * kTileBytes, &tm_qkv, &st_full[st], 0, 2 * H + h, i * kTile, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 0, 0, 128);
      constexpr uint32_t idesc_tt = umma_idesc_bf16(64, 1, 1, 128);  // A MN-major, B MN-major (dV, dK)
      constexpr uint32_t idesc_q = umma_idesc_bf16(64, 0, 1, 128);   // A K-major, B MN-major (dQ)
      const uint32_t aR0 = smem_u32(sR0), aR1 = smem_u32(sR1);
      auto issue_s_dp = [&](int st) {
        const uint32_t aX = smem_u32(sX + st * kTileBytes), aY = smem_u32(sY + st * kTileBytes);
        // S = Q K^T, dP = dO V^T  (rows = q, cols = kv)
        const uint32_t q_t = MODE_KV ? aX : aR0, k_t = MODE_KV ? aR0 : aX;
        const uint32_t do_t = MODE_KV ? aY : aR1, v_t = MODE_KV ? aR1 : aY;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tS, desc_k(q_t, k), desc_k(k_t, k), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tDP, desc_k(do_t, k), desc_k(v_t, k), idesc_s, k > 0);
      };
      mbar_wait(r_full, 0);
      mbar_wait(&st_full[0], 0);
      tc_fence_after();
      issue_s_dp(0);
      umma_commit(sdp_full);
      for (int i = 0; i < n_it; ++i) {
        if (i + 1 < n_it) {
          const int st = (i + 1) % kBwdStages;
          mbar_wait(&st_full[st], (uint32_t)((i + 1) / kBwdStages) & 1u);
          mbar_wait(sdp_free, (uint32_t)i & 1u);
          tc_fence_after();
          issue_s_dp(st);
          umma_commit(sdp_full);
        }
        mbar_wait(pds_full, (uint32_t)i & 1u);
        tc_fence_after();
        const int cst = i % kBwdStages;
        const uint32_t aX = smem_u32(sX + cst * kTileBytes), aY = smem_u32(sY + cst * kTileBytes);
        const uint32_t aP = smem_u32(sP), aD = smem_u32(sD);
        if (MODE_KV) {
          // dV[kv, d] += P^T dO : A = P (MN-major: m = kv, k = q), B = dO_i (MN-major: n = d, k = q)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16_ss(tA0, desc_mn(aP, k, kTileBytes), desc_mn(aY, k, 8192), idesc_tt, (i > 0 || k > 0));
          // dK[kv, d] += dS^T Q
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16_ss(tA1, desc_mn(aD, k, kTileBytes), desc_mn(aX, k, 8192), idesc_tt, (i > 0 || k > 0));
        } else {
          // dQ[q, d] += dS K : A = dS (K-major: m = q, k = kv), B = K_j (MN-major: n = d, k = kv)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16_ss(tA0, desc_k(aD + (k >> 2) * kTileBytes, k & 3), desc_mn(aX, k, 8192), idesc_q,
                         (i > 0 || k > 0));
        }
        umma_commit(pds_free);
        umma_commit(&st_empty[cst]);
      }
    }
  } else {
    const int e = warp - 2;
    const int q = warp & 3;
    const int half = e >> 2;
    const int row = q * 32 + lane;  // S / dP row = query index within the tile
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float c_log2 = scale * kLog2e;
    const float* lse_bh = lse + ((long long)b * H + h) * S;
    const float* dl_bh = delta + ((long long)b * H + h) * S;
    // row statistics of query row `qg`: lse in the log2 domain and delta pre-multiplied by the softmax scale
    // row statistics of query row `qg`: only the loads here (clamped address); masking and scaling happen at use, so a
    // prefetch issued one tile ahead never stalls on its own result
    auto load_stats = [&](int qg, float& l_raw, float& d_raw) {
      const int qc = min(qg, S - 1);
      l_raw = __ldg(lse_bh + qc);
      d_raw = __ldg(dl_bh + qc);
    };
    float lse_r, dl_r, lse_n = 0.f, dl_n = 0.f;
    load_stats(MODE_KV ? row : own0 + row, lse_r, dl_r);
    const uint64_t cl2 = pack2(c_log2, c_log2), sc2 = pack2(scale, scale);
    for (int i = 0; i < n_it; ++i) {
      const int kv_valid = MODE_KV ? S - own0 : S - i * kTile;
      const bool q_ok = (MODE_KV ? i * kTile + row : own0 + row) < S;
      mbar_wait(sdp_full, (uint32_t)i & 1u);
      tc_fence_after();
      uint32_t sv[2][32], dv[2][32];
#pragma unroll
      for (int c = 0; c < 2; ++c) tmem_ld_32x32b_x32(tS + lane_base + half * 64 + c * 32, sv[c]);
#pragma unroll
      for (int c = 0; c < 2; ++c) tmem_ld_32x32b_x32(tDP + lane_base + half * 64 + c * 32, dv[c]);
      if (MODE_KV && i + 1 < n_it) load_stats((i + 1) * kTile + row, lse_n, dl_n);  // in flight during this tile
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_free);
      const float lse2 = q_ok ? lse_r * kLog2e : 0.f, dls = q_ok ? dl_r * scale : 0.f;
      const uint64_t nl2 = pack2(-lse2, -lse2), nd2 = pack2(-dls, -dls);
      const bool tail = kv_valid < kTile;
      // all of this tile's P / dS into registers (packed bf16) first: the math overlaps the previous tile's
      // dV/dK/dQ MMAs, which are still reading the shared-memory P / dS tiles
      uint4 pk[8], dk[8];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float p[8], ds[8];
#pragma unroll
          for (int t = 0; t < 8; t += 2) {
            const int j = g * 8 + t;
            const uint64_t x2 = ffma2(pack2(__uint_as_float(sv[c][j]), __uint_as_float(sv[c][j + 1])), cl2, nl2);
            float a0, a1;
            unpack2(x2, a0, a1);
            p[t] = exp2f(a0);
            p[t + 1] = exp2f(a1);
            if (tail) {  // warp-uniform: only the last K/V tile of the sequence has columns past the end
              const int col = half * 64 + c * 32 + j;
              if (col >= kv_valid) p[t] = 0.f;
              if (col + 1 >= kv_valid) p[t + 1] = 0.f;
            }
            // dS = P * (dP - delta) * scale
            const uint64_t y2 = ffma2(pack2(__uint_as_float(dv[c][j]), __uint_as_float(dv[c][j + 1])), sc2, nd2);
            unpack2(fmul2(pack2(p[t], p[t + 1]), y2), ds[t], ds[t + 1]);
          }
          if (MODE_KV) {
            pk[c * 4 + g].x = pack_bf16x2(p[0], p[1]); pk[c * 4 + g].y = pack_bf16x2(p[2], p[3]);
            pk[c * 4 + g].z = pack_bf16x2(p[4], p[5]); pk[c * 4 + g].w = pack_bf16x2(p[6], p[7]);
          }
          dk[c * 4 + g].x = pack_bf16x2(ds[0], ds[1]); dk[c * 4 + g].y = pack_bf16x2(ds[2], ds[3]);
          dk[c * 4 + g].z = pack_bf16x2(ds[4], ds[5]); dk[c * 4 + g].w = pack_bf16x2(ds[6], ds[7]);
        }
      if (i > 0) mbar_wait(pds_free, (uint32_t)(i - 1) & 1u);  // previous dV/dK/dQ MMAs finished reading P / dS
#pragma unroll
      for (int cg = 0; cg < 8; ++cg) {
        const uint32_t off = ptile_chunk_off(row, half * 8 + cg);
        if (MODE_KV) *reinterpret_cast<uint4*>(sP + off) = pk[cg];
        *reinterpret_cast<uint4*>(sD + off) = dk[cg];
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (MODE_KV) {
        lse_r = lse_n;
        dl_r = dl_n;
      }
    }
    // epilogue: the accumulators (rows = owned tile rows, 64 cols); this warp writes 32 of the 64 columns
    mbar_wait(pds_free, (uint32_t)(n_it - 1) & 1u);
    tc_fence_after();
    const int rg = own0 + row;
    const long long tok = (long long)b * S + rg;
    auto store32 = [&](uint32_t tacc, int which) {
      uint32_t ov[32];
      tmem_ld_32x32b_x32(tacc + lane_base + half * 32, ov);
      tmem_ld_wait();
      if (rg < S) {
        bf16* dst = dqkv + ((tok * 3 + which) * H + h) * 64 + half * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(ov[g * 8 + 0]), __uint_as_float(ov[g * 8 + 1]));
          pk.y = pack_bf16x2(__uint_as_float(ov[g * 8 + 2]), __uint_as_float(ov[g * 8 + 3]));
          pk.z = pack_bf16x2(__uint_as_float(ov[g * 8 + 4]), __uint_as_float(ov[g * 8 + 5]));
          pk.w = pack_bf16x2(__uint_as_float(ov[g * 8 + 6]), __uint_as_float(ov[g * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + g * 8) = pk;
        }
      }
    };
    if (MODE_KV) {
      store32(tA0, 2);  // dV
      store32(tA1, 1);  // dK
    } else {
      store32(tA0, 0);  // dQ
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int make_head_tmap(CUtensorMap* tm, const void* base, int heads_total, int S, int B) {
  // [B, S, heads_total, 64] bf16 viewed as a 4-D tensor; box = 64 x 1 x 128 x 1 -> one [128 rows][64] tile
  const uint64_t dims[4] = {64, (uint64_t)heads_total, (uint64_t)S, (uint64_t)B};
  const uint64_t strides[3] = {128, (uint64_t)heads_total * 128, (uint64_t)S * heads_total * 128};
  const uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace bvc

using namespace bvc;

extern "C" int bvc_attn_fwd(const void* qkv, int32_t B, int32_t S, int32_t H, float scale, void* out, float* lse,
                            void* stream) {
  BVC_CHECK_ARG(qkv && out && lse && B > 0 && S > 0 && H > 0);
  BVC_CHECK_ARG((((uintptr_t)qkv) & 15) == 0 && (((uintptr_t)out) & 15) == 0);
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem) != cudaSuccess)
      return BVC_ERR_LAUNCH;
    attr_done = true;
  }
  CUtensorMap tm;
  int rc = make_head_tmap(&tm, qkv, 3 * H, S, B);
  if (rc) return rc;
  dim3 grid((S + kTile - 1) / kTile, H, B);
  attn_fwd_kernel<<<grid, kFwdThreads, kFwdSmem, (cudaStream_t)stream>>>(tm, (bf16*)out, lse, S, H, scale * kLog2e);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int32_t B, int32_t S,
                            int32_t H, float scale, float* delta, void* dqkv, void* stream) {
  BVC_CHECK_ARG(qkv && out && dout && lse && delta && dqkv && B > 0 && S > 0 && H > 0);
  BVC_CHECK_ARG((((uintptr_t)qkv) & 15) == 0 && (((uintptr_t)dout) & 15) == 0 && (((uintptr_t)dqkv) & 15) == 0);
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem) != cudaSuccess)
      return BVC_ERR_LAUNCH;
    attr_done = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)B * S * H;
  long long g = (rows + 63) / 64;
  if (g > (long long)num_sms() * 8) g = (long long)num_sms() * 8;
  attn_delta_kernel<<<(int)g, 256, 0, st>>>((const bf16*)out, (const bf16*)dout, delta, rows, S, H);
  BVC_CHECK_LAUNCH();
  CUtensorMap tq, td;
  int rc = make_head_tmap(&tq, qkv, 3 * H, S, B);
  if (rc) return rc;
  rc = make_head_tmap(&td, dout, H, S, B);
  if (rc) return rc;
  dim3 grid((S + kTile - 1) / kTile, H, B);
  attn_bwd_kernel<1><<<grid, kBwdThreads, kBwdSmem, st>>>(tq, td, lse, delta, (bf16*)dqkv, S, H, scale);
  BVC_CHECK_LAUNCH();
  attn_bwd_kernel<0><<<grid, kBwdThreads, kBwdSmem, st>>>(tq, td, lse, delta, (bf16*)dqkv, S, H, scale);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}
