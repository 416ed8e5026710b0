// attn_small.cu -- attention forward for SHORT sequences (S <= 192: the ViT encoder's 160 visible tokens at mask ratio
// 0.9), head_dim 64.  Same contract as attn_fwd_kernel (attn.cu); bvc_attn_fwd dispatches here.
//
// Why a second kernel: at S = 160 the general kernel walks two ragged 128-wide K/V tiles per query tile (39 % of the
// score math is useful) through a pipeline that is refilled every two tiles.  Here the WHOLE sequence of one (clip,
// head) is resident: K and V are two contiguous [128][64] tiles, so ONE tcgen05.mma with N = S16 = roundup16(S) gives a
// query tile's full score row (the 128B-swizzled 8-row groups of the second tile continue the first tile's), the
// softmax is a plain two-pass row softmax (no running maximum, no rescale), P goes back into the TMEM columns its
// scores came from (packed bf16, chunk c of P lands on columns of score chunks <= c, which are already in registers)
// and feeds O = P V as an A-from-TMEM operand with S16 / 16 k-steps.
//
// Persistent, one CTA per SM over (clip, head) items, shared memory double-buffered across items:
//   warp 0 TMA loads (Q0, Q1, K0|K1, V0|V1 of item k + 1 while item k computes), warp 1 MMA issue, warp 2 TMA stores,
//   warps 4-7 / 8-11 softmax + epilogue of query tile 0 / 1 (thread = one query row).
// TMEM (512 columns): S0 / P0 at 0, S1 / P1 at 192, O0 at 384, O1 at 448.
#include "attn_common.cuh"

namespace bvc {

constexpr int kSmallMaxS = 192;
constexpr int kSmThreads = 384;
constexpr int kSmItemBytes = 6 * kTileBytes;               // Q0 Q1 | K0 K1 | V0 V1
constexpr int kSmSmem = 2 * kSmItemBytes + 256;

__global__ void __launch_bounds__(kSmThreads, 1)
attn_small_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out,
                      float* __restrict__ lse, int S, int H, int n_work, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kSmItemBytes);
  uint64_t* in_full = bars + 0;    // [2] TMA -> MMA: the item's six tiles landed
  uint64_t* in_empty = bars + 2;   // [2] store warp -> TMA: the item's buffer (Q tiles doubled as O staging) is free
  uint64_t* s_full = bars + 4;     // [2 query tiles] MMA -> softmax group
  uint64_t* p_full = bars + 6;     // [2] softmax group -> MMA: P is in TMEM
  uint64_t* o_full = bars + 8;     // [2] MMA -> softmax group: O = P V complete
  uint64_t* o_free = bars + 10;    // [2] softmax group -> MMA: O (and with it S / P) of the item has been read
  uint64_t* epi_full = bars + 12;  // softmax groups -> store warp: O tiles staged
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S16 = (S + 15) & ~15;
  const int n_q = S > kTile ? 2 : 1;  // query tiles (= softmax groups at work)
  const int n_my = (n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto item_bh = [&](int k) { return (int)blockIdx.x + k * (int)gridDim.x; };

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_out);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&in_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 4);
    }
    mbar_init(epi_full, 4 * n_q);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int k = 0; k < n_my; ++k) {
        const int bh = item_bh(k), h = bh % H, b = bh / H, buf = k & 1;
        uint8_t* base = smem + buf * kSmItemBytes;
        mbar_wait(&in_empty[buf], ((uint32_t)(k >> 1) & 1u) ^ 1u);
        mbar_expect_tx(&in_full[buf], (n_q == 2 ? 6 : 3) * kTileBytes);
        tma_load_4d(base, &tm_qkv, &in_full[buf], 0, h, 0, b);
        tma_load_4d(base + 2 * kTileBytes, &tm_qkv, &in_full[buf], 0, H + h, 0, b);
        tma_load_4d(base + 4 * kTileBytes, &tm_qkv, &in_full[buf], 0, 2 * H + h, 0, b);
        if (n_q == 2) {  // rows 128 .. 255: zero-filled past S
          tma_load_4d(base + kTileBytes, &tm_qkv, &in_full[buf], 0, h, kTile, b);
          tma_load_4d(base + 3 * kTileBytes, &tm_qkv, &in_full[buf], 0, H + h, kTile, b);
          tma_load_4d(base + 5 * kTileBytes, &tm_qkv, &in_full[buf], 0, 2 * H + h, kTile, b);
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform control flow, one elected lane issues (see elect_one in bvc_ptx.cuh)
    const uint32_t idesc_s = umma_idesc_bf16(S16, 0, 0, 128);
    constexpr uint32_t idesc_o = umma_idesc_bf16(64, 0, 1, 128);
    const int ksteps = S16 >> 4;
    for (int k = 0; k < n_my; ++k) {
      const int buf = k & 1;
      const uint32_t base = smem_u32(smem + buf * kSmItemBytes);
      mbar_wait(&in_full[buf], (uint32_t)(k >> 1) & 1u);
      for (int t = 0; t < n_q; ++t) {
        // S_t overwrites the previous item's S / P columns: its P V must be complete (the O columns are separate, so
        // the scores of item k are computed while the group still drains O of item k - 1)
        if (k > 0) mbar_wait(&o_full[t], (uint32_t)(k - 1) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dQ = desc_k(base + t * kTileBytes, 0), dK = desc_k(base + 2 * kTileBytes, 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ss(tmem_base + (uint32_t)t * 192, dQ + 2 * kk, dK + 2 * kk, idesc_s, kk > 0);
          umma_commit(&s_full[t]);
        }
        __syncwarp();
      }
      for (int t = 0; t < n_q; ++t) {
        mbar_wait(&p_full[t], (uint32_t)k & 1u);
        if (k > 0) mbar_wait(&o_free[t], (uint32_t)(k - 1) & 1u);  // P V overwrites O_t: the group has read item k - 1's
        tc_fence_after();
        if (elect_one()) {
          // O_t = P_t V: A = P (TMEM, packed bf16, 8 columns per K = 16 step), B = V (MN-major: n = d, k = key; the
          // second V tile continues the first one's 16-row k-steps)
          const uint64_t dV = desc_mn(base + 4 * kTileBytes, 0, 8192);
          const uint32_t aP = tmem_base + (uint32_t)t * 192, dO = tmem_base + 384 + (uint32_t)t * 64;
          for (int kk = 0; kk < ksteps; ++kk) umma_bf16_ts(dO, aP + kk * 8, dV + 128 * kk, idesc_o, kk > 0);
          umma_commit(&o_full[t]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    for (int k = 0; k < n_my; ++k) {
      const int bh = item_bh(k), h = bh % H, b = bh / H, buf = k & 1;
      uint8_t* base = smem + buf * kSmItemBytes;
      mbar_wait(epi_full, (uint32_t)k & 1u);
      if (lane == 0) {
        tma_store_4d(&tm_out, base, 0, h, 0, b);  // the store clips the rows past the end of the sequence
        if (n_q == 2) tma_store_4d(&tm_out, base + kTileBytes, 0, h, kTile, b);
        tma_store_commit();
        tma_store_wait_read0();
        mbar_arrive(&in_empty[buf]);
      }
      __syncwarp();
    }
    if (lane == 0) tma_store_wait0();
  } else if (warp >= 4) {
    const int q4 = warp & 3;
    const int t = (warp - 4) >> 2;
    if (t < n_q) {
      const int row = q4 * 32 + lane;
      const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
      const uint32_t mS = tmem_base + (uint32_t)t * 192 + lane_base, mO = tmem_base + 384 + (uint32_t)t * 64 + lane_base;
      const uint64_t c2 = pack2(scale_log2, scale_log2);
      const int n32 = S16 >> 5;          // full 32-column chunks of the score row
      const bool tail16 = (S16 & 16) != 0;  // plus one 16-column chunk
      // a warp whose 32 rows all lie past the end of the sequence (S = 160: three of the second tile's four warps)
      // only keeps the barrier protocol alive: its rows of P / O are garbage that no valid row depends on and that the
      // TMA store clips, and it no longer competes for the MUFU pipe with the warps that have work
      const bool live = t * kTile + q4 * 32 < S;
      for (int k = 0; k < n_my; ++k) {
        const int bh = item_bh(k), buf = k & 1;
        mbar_wait(&s_full[t], (uint32_t)k & 1u);
        if (!live) {
          if (lane == 0) mbar_arrive(&p_full[t]);
          mbar_wait(&o_full[t], (uint32_t)k & 1u);
          if (lane == 0) {
            mbar_arrive(&o_free[t]);
            mbar_arrive(epi_full);
          }
          continue;
        }
        tc_fence_after();
        // pass 1: row maximum of the raw scores (columns >= S are padding: K rows zero-filled by TMA)
        float mx = -INFINITY;
        for (int c = 0; c < n32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(mS + c * 32, v);
          tmem_ld_wait();
          tmem_ld_pin(v);
          if (c * 32 + 32 <= S) {
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[i]));
            mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < S) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
        if (tail16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(mS + n32 * 32, v);
          tmem_ld_wait();
          tmem_ld_pin(v);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n32 * 32 + i < S) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        const float m = mx * scale_log2;  // log2 domain (scale_log2 > 0)
        const uint64_t m2 = pack2(-m, -m);
        // pass 2: P = exp2(S * scale_log2 - m), row sum, packed bf16 P over the score columns already consumed
        float l = 0.f;
        uint64_t l2 = pack2(0.f, 0.f);
        for (int c = 0; c < n32; ++c) {
          uint32_t v[32], pk[16];
          tmem_ld_32x32b_x32(mS + c * 32, v);
          tmem_ld_wait();
          tmem_ld_pin(v);
          if (c * 32 + 32 <= S) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const uint64_t e2 = exp2_mufu2(ffma2(pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2, m2));
              float p0, p1;
              unpack2(e2, p0, p1);
              l2 = fadd2(l2, e2);
              pk[i >> 1] = pack_bf16x2(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float p0, p1;
              unpack2(exp2_mufu2(ffma2(pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2, m2)), p0, p1);
              if (c * 32 + i >= S) p0 = 0.f;
              if (c * 32 + i + 1 >= S) p1 = 0.f;
              l += p0 + p1;
              pk[i >> 1] = pack_bf16x2(p0, p1);
            }
          }
          tmem_st_32x32b_x16(mS + c * 16, pk);
        }
        {
          float la, lb;
          unpack2(l2, la, lb);
          l += la + lb;
        }
        if (tail16) {
          uint32_t v[16], pk[8];
          tmem_ld_32x32b_x16(mS + n32 * 32, v);
          tmem_ld_wait();
          tmem_ld_pin(v);
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float p0, p1;
            unpack2(exp2_mufu2(ffma2(pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2, m2)), p0, p1);
            if (n32 * 32 + i >= S) p0 = 0.f;
            if (n32 * 32 + i + 1 >= S) p1 = 0.f;
            l += p0 + p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x32b_x8(mS + n32 * 16, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        // epilogue: O / l -> bf16 -> the item's Q tile (no MMA reads it any more), 128B-swizzled -> TMA store
        mbar_wait(&o_full[t], (uint32_t)k & 1u);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        uint8_t* stage = smem + buf * kSmItemBytes + t * kTileBytes;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t ov[32];
          tmem_ld_32x32b_x32(mO + c * 32, ov);
          tmem_ld_wait();
          tmem_ld_pin(ov);
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            uint4 o4;
            o4.x = pack_bf16x2(__uint_as_float(ov[gq * 8 + 0]) * inv_l, __uint_as_float(ov[gq * 8 + 1]) * inv_l);
            o4.y = pack_bf16x2(__uint_as_float(ov[gq * 8 + 2]) * inv_l, __uint_as_float(ov[gq * 8 + 3]) * inv_l);
            o4.z = pack_bf16x2(__uint_as_float(ov[gq * 8 + 4]) * inv_l, __uint_as_float(ov[gq * 8 + 5]) * inv_l);
            o4.w = pack_bf16x2(__uint_as_float(ov[gq * 8 + 6]) * inv_l, __uint_as_float(ov[gq * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(stage + row * 128 + (((c * 4 + gq) ^ (row & 7)) << 4)) = o4;
          }
        }
        const int qrow = t * kTile + row;
        if (qrow < S) lse[(long long)bh * S + qrow] = (m + log2f(l)) * 0.6931471805599453f;
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&o_free[t]);
          mbar_arrive(epi_full);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================ backward, S <= 160
// ONE pass per (clip, head) with everything resident (the general kernels need two passes, each recomputing the
// scores, and walk two ragged 128-tiles per dimension at S = 160: 7 MMAs and 2 exponentials per score where 5 and 1
// suffice, plus a pipeline refill every two tiles).  Rows = queries:
//   per query tile t:  S_t = Q_t K^T, dP_t = dO_t V^T (N = S16)  ->  P = exp2(S c - lse), dS = P (dP - delta) scale
//                      written as bf16 to shared memory in [query row][key] order, one 64-key atom after the other;
//                      dQ_t = dS_t K reads it K-major (A operand, M = queries),
//   once per item:     dV = P^T dO and dK = dS^T Q read THE SAME tiles MN-major (A operand, M = keys, K = queries):
//                      the UMMA descriptor's major bit transposes P / dS for free, nothing is stored twice.
// TMEM: S [0,160), dP [160,320), dQ_0 [320,384), dQ_1 [384,448); when the last tile's scores have been read, dV_0, dV_1,
// dK_0, dK_1 take [0,256).  The six output tiles leave through the (then dead) P / dS buffers and TMA stores, which clip
// the rows past the end of the sequence.  The second key tile (keys 128-255) of dV / dK reads a fourth "atom" that is
// simply the memory after the third one (finite bf16 data): it only feeds accumulator rows >= 192 that no one stores.
// Shared memory: P, dS 3 x 20 KB each | Q, dO, K, V 20 KB each (160 rows: a 128-row and a 32-row TMA box).
constexpr int kSbMaxS = 160;
constexpr int kSbOperand = 160 * 128;          // [160 rows][64] bf16, 128B-swizzled
constexpr int kSbAtom = kSbOperand;            // one 64-key atom of P / dS: [160 query rows][64 keys]
constexpr int kSbP = 0, kSbDS = 3 * kSbAtom, kSbQ = 6 * kSbAtom, kSbDO = kSbQ + kSbOperand, kSbK = kSbDO + kSbOperand,
              kSbV = kSbK + kSbOperand, kSbBars = kSbV + kSbOperand;
constexpr int kSbSmem = kSbBars + 256;
static_assert(kSbSmem <= 232448, "small backward: shared memory");

__global__ void __launch_bounds__(kSmThreads, 1)
attn_small_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_qkv32,
                      const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_do32,
                      const __grid_constant__ CUtensorMap tm_dqkv, const float* __restrict__ lse,
                      const float* __restrict__ delta, int S, int H, int n_work, float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSbBars);
  uint64_t* in_full = bars + 0;      // TMA -> MMA
  uint64_t* k_free = bars + 1;       // MMA (commit) -> TMA: K's last reader (dQ of the last tile, the item's last MMAs) is done
  uint64_t* sdp_full = bars + 2;     // [2] MMA -> compute: S_t / dP_t
  uint64_t* sdp_free = bars + 4;     // [2] compute -> MMA: tile t's scores are in registers
  uint64_t* pds_full = bars + 6;     // [2] compute -> MMA: P_t / dS_t are in shared memory
  uint64_t* acc_full = bars + 8;     // MMA -> compute: dQ, dV, dK complete
  uint64_t* acc_free = bars + 9;     // compute -> MMA: accumulators read
  uint64_t* epi_full = bars + 10;    // compute -> store warp: output tiles staged
  uint64_t* stage_free = bars + 11;  // store warp -> compute: staging (= P / dS buffers) read by the TMA stores
  uint64_t* v_free = bars + 12;      // MMA (commit) -> TMA: V's last reader (dP of the last tile) is done
  uint64_t* qdo_free = bars + 13;    // MMA (commit) -> TMA: Q's and dO's last readers (dK, dV) are done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S16 = (S + 15) & ~15;
  const int n_q = S > kTile ? 2 : 1;   // query tiles = key tiles
  const int ksteps = S16 >> 4;
  const int n_my = (n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto item_bh = [&](int k) { return (int)blockIdx.x + k * (int)gridDim.x; };
  const uint32_t sbase = smem_u32(smem);

  if (threadIdx.x == 0) {
    if ((sbase & 1023u) != 0) __trap();
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_qkv32);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_do32);
    tma_prefetch_desc(&tm_dqkv);
    mbar_init(in_full, 1);
    mbar_init(k_free, 1);
    mbar_init(v_free, 1);
    mbar_init(qdo_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1);
      mbar_init(&sdp_free[i], 8);
      mbar_init(&pds_full[i], 8);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 8);
    mbar_init(epi_full, 8);
    mbar_init(stage_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 160, tDQ = tmem_base + 320;  // dQ_t at tDQ + 64 t
  // end of item: dV_j at tmem_base + 64 j, dK_j at tmem_base + 128 + 64 j

  if (warp == 0) {
    if (lane == 0) {
      for (int k = 0; k < n_my; ++k) {
        const int bh = item_bh(k), h = bh % H, b = bh / H;
        // each operand of item k is fetched as soon as its last reader of item k - 1 has finished -- V after the last
        // dP, Q / dO after dK / dV, K after the last dQ -- instead of all four after the item's last MMA: the loads of
        // the 148 CTAs, which run in lockstep, no longer hit L2 as one 12 MB burst that nothing overlaps
        const uint32_t ph = (uint32_t)(k - 1) & 1u;
        if (k > 0) mbar_wait(v_free, ph);
        mbar_expect_tx(in_full, 4 * kTileBytes + (n_q == 2 ? 4 * 32 * 128 : 0));
        tma_load_4d(smem + kSbV, &tm_qkv, in_full, 0, 2 * H + h, 0, b);
        if (n_q == 2) tma_load_4d(smem + kSbV + kTileBytes, &tm_qkv32, in_full, 0, 2 * H + h, kTile, b);
        if (k > 0) mbar_wait(qdo_free, ph);
        tma_load_4d(smem + kSbQ, &tm_qkv, in_full, 0, h, 0, b);
        tma_load_4d(smem + kSbDO, &tm_do, in_full, 0, h, 0, b);
        if (n_q == 2) {  // rows 128 .. 159 (zero-filled past S)
          tma_load_4d(smem + kSbQ + kTileBytes, &tm_qkv32, in_full, 0, h, kTile, b);
          tma_load_4d(smem + kSbDO + kTileBytes, &tm_do32, in_full, 0, h, kTile, b);
        }
        if (k > 0) mbar_wait(k_free, ph);
        tma_load_4d(smem + kSbK, &tm_qkv, in_full, 0, H + h, 0, b);
        if (n_q == 2) tma_load_4d(smem + kSbK + kTileBytes, &tm_qkv32, in_full, 0, H + h, kTile, b);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = umma_idesc_bf16(S16, 0, 0, 128);           // S, dP: A, B K-major, N = S16
    constexpr uint32_t idesc_dq = umma_idesc_bf16(64, 0, 1, 128);       // dQ: A = dS K-major, B = K MN-major
    constexpr uint32_t idesc_dkv = umma_idesc_bf16(64, 1, 1, 128);      // dV, dK: A = P^T / dS^T MN-major, B MN-major
    for (int k = 0; k < n_my; ++k) {
      mbar_wait(in_full, (uint32_t)k & 1u);
      if (k > 0) mbar_wait(acc_free, (uint32_t)(k - 1) & 1u);  // S / dP overwrite the previous item's dV / dK columns
      for (int t = 0; t < n_q; ++t) {
        if (t > 0) mbar_wait(&sdp_free[t - 1], (uint32_t)k & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dQ = desc_k(sbase + kSbQ + t * kTileBytes, 0), dK = desc_k(sbase + kSbK, 0);
          const uint64_t dO = desc_k(sbase + kSbDO + t * kTileBytes, 0), dV = desc_k(sbase + kSbV, 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tS, dQ + 2 * kk, dK + 2 * kk, idesc_s, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tDP, dO + 2 * kk, dV + 2 * kk, idesc_s, kk > 0);
          umma_commit(&sdp_full[t]);
          if (t == n_q - 1) umma_commit(v_free);  // V's last reader
        }
        __syncwarp();
      }
      // dQ_t = dS_t K: A rows = tile t's queries, k-step kk = keys 16 kk .. : atom kk / 4, 32 bytes per step inside it
      auto issue_dq = [&](int t) {
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t dA = umma_smem_desc(sbase + kSbDS + (kk >> 2) * kSbAtom + t * kTileBytes + (kk & 3) * 32, 1024, 16);
          const uint64_t dB = umma_smem_desc(sbase + kSbK + kk * 2048, 1024, 8192);
          umma_bf16_ss(tDQ + (uint32_t)t * 64, dA, dB, idesc_dq, kk > 0);
        }
      };
      for (int t = 0; t < n_q; ++t) {
        mbar_wait(&pds_full[t], (uint32_t)k & 1u);
        tc_fence_after();
        if (t < n_q - 1) {  // the first tile's dQ runs under the second tile's math
          if (elect_one()) issue_dq(t);
          __syncwarp();
        }
      }
      // every tile's scores have been read (pds_full follows the loads): dV_j / dK_j may take the S / dP columns; the
      // last tile's dQ goes after them, so that Q / dO (two of the four operands) are released before the item's end
      if (elect_one()) {
        for (int j = 0; j < n_q; ++j) {
          for (int kk = 0; kk < ksteps; ++kk) {  // k = queries, 16 rows = 2048 bytes per step
            const uint64_t aP = umma_smem_desc(sbase + kSbP + 2 * j * kSbAtom + kk * 2048, 1024, kSbAtom);
            const uint64_t bO = umma_smem_desc(sbase + kSbDO + kk * 2048, 1024, 8192);
            umma_bf16_ss(tmem_base + (uint32_t)j * 64, aP, bO, idesc_dkv, kk > 0);
          }
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t aS = umma_smem_desc(sbase + kSbDS + 2 * j * kSbAtom + kk * 2048, 1024, kSbAtom);
            const uint64_t bQ = umma_smem_desc(sbase + kSbQ + kk * 2048, 1024, 8192);
            umma_bf16_ss(tmem_base + 128 + (uint32_t)j * 64, aS, bQ, idesc_dkv, kk > 0);
          }
        }
        umma_commit(qdo_free);
        issue_dq(n_q - 1);
        umma_commit(acc_full);
        umma_commit(k_free);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    for (int k = 0; k < n_my; ++k) {
      const int bh = item_bh(k), h = bh % H, b = bh / H;
      mbar_wait(epi_full, (uint32_t)k & 1u);
      if (lane == 0) {
        for (int j = 0; j < n_q; ++j) {  // staging tiles: dQ_j, dV_j, dK_j at 3 j, 3 j + 1, 3 j + 2
          tma_store_4d(&tm_dqkv, smem + (3 * j + 0) * kTileBytes, 0, h, j * kTile, b);
          tma_store_4d(&tm_dqkv, smem + (3 * j + 1) * kTileBytes, 0, 2 * H + h, j * kTile, b);
          tma_store_4d(&tm_dqkv, smem + (3 * j + 2) * kTileBytes, 0, H + h, j * kTile, b);
        }
        tma_store_commit();
        tma_store_wait_read0();
        mbar_arrive(stage_free);
      }
      __syncwarp();
    }
    if (lane == 0) tma_store_wait0();
  } else if (warp >= 4) {
    const int q4 = warp & 3, hh = (warp - 4) >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const int Wh = S16 >> 1;  // columns per half (multiple of 8)
    const float c_log2 = scale * kLog2e;
    const uint64_t cl2 = pack2(c_log2, c_log2), sc2 = pack2(scale, scale);
    for (int k = 0; k < n_my; ++k) {
      const int bh = item_bh(k);
      if (k > 0) mbar_wait(stage_free, (uint32_t)(k - 1) & 1u);  // the previous item's stores have read P / dS space
      for (int t = 0; t < n_q; ++t) {
        const bool live = t * kTile + q4 * 32 < S;
        const int r = t * kTile + row;  // query row = row of the P / dS atoms
        float nl = 0.f, nd = 0.f;
        if (live) {
          const long long o = (long long)bh * S + min(r, S - 1);
          nl = -__ldg(lse + o) * kLog2e;
          nd = -__ldg(delta + o) * scale;
        }
        const uint64_t nl2 = pack2(nl, nl), nd2 = pack2(nd, nd);
        mbar_wait(&sdp_full[t], (uint32_t)k & 1u);
        tc_fence_after();
        if (live) {
          // groups of 16 columns (a last group may have 8); the TMEM loads of group g + 1 are in flight while group g
          // is computed and stored (two named register sets: a dynamically indexed one would live in local memory)
          const int c_end = (hh + 1) * Wh;
          uint32_t svA[2][8], dvA[2][8], svB[2][8], dvB[2][8];
          auto issue = [&](int c0, uint32_t (&sv)[2][8], uint32_t (&dv)[2][8]) {
            tmem_ld_32x32b_x8(tS + lane_base + c0, sv[0]);
            tmem_ld_32x32b_x8(tDP + lane_base + c0, dv[0]);
            if (c0 + 16 <= c_end) {
              tmem_ld_32x32b_x8(tS + lane_base + c0 + 8, sv[1]);
              tmem_ld_32x32b_x8(tDP + lane_base + c0 + 8, dv[1]);
            }
          };
          auto compute = [&](int c0, uint32_t (&sv)[2][8], uint32_t (&dv)[2][8]) {
            const bool two = c0 + 16 <= c_end;
            tmem_ld_pin(sv[0]);
            tmem_ld_pin(dv[0]);
            if (two) {
              tmem_ld_pin(sv[1]);
              tmem_ld_pin(dv[1]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (u == 1 && !two) break;
              uint32_t pk[4], dk[4];
#pragma unroll
              for (int i = 0; i < 8; i += 2) {
                const uint64_t s2 = pack2(__uint_as_float(sv[u][i]), __uint_as_float(sv[u][i + 1]));
                const uint64_t p2 = pack2(__uint_as_float(dv[u][i]), __uint_as_float(dv[u][i + 1]));
                const uint64_t e2 = exp2_mufu2(ffma2(s2, cl2, nl2));
                float p0, p1, d0, d1;
                unpack2(e2, p0, p1);
                unpack2(fmul2(e2, ffma2(p2, sc2, nd2)), d0, d1);
                pk[i >> 1] = pack_bf16x2(p0, p1);
                dk[i >> 1] = pack_bf16x2(d0, d1);
              }
              const int col = c0 + 8 * u;  // 8 keys = one 16-byte chunk of the 128B-swizzled row
              const uint32_t off = (uint32_t)((col >> 6) * kSbAtom + r * 128 + ((((col & 63) >> 3) ^ (r & 7)) << 4));
              sts_u4(sbase + kSbP + off, pk[0], pk[1], pk[2], pk[3]);
              sts_u4(sbase + kSbDS + off, dk[0], dk[1], dk[2], dk[3]);
            }
          };
          issue(hh * Wh, svA, dvA);
          for (int c0 = hh * Wh; c0 < c_end; c0 += 32) {
            tmem_ld_wait();
            if (c0 + 16 < c_end) issue(c0 + 16, svB, dvB);
            compute(c0, svA, dvA);
            if (c0 + 16 < c_end) {
              tmem_ld_wait();
              if (c0 + 32 < c_end) issue(c0 + 32, svA, dvA);
              compute(c0 + 16, svB, dvB);
            }
          }
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&sdp_free[t]);
          mbar_arrive(&pds_full[t]);
        }
      }
      // drain: dQ_j, dV_j, dK_j -> bf16 -> staging tiles 3 j, 3 j + 1, 3 j + 2 (over the dead P / dS buffers)
      mbar_wait(acc_full, (uint32_t)k & 1u);
      tc_fence_after();
      for (int j = 0; j < n_q; ++j) {
        if (j * kTile + q4 * 32 >= S) continue;  // rows the stores clip
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const uint32_t tacc = a == 0 ? tDQ + (uint32_t)j * 64
                                       : (a == 1 ? tmem_base + (uint32_t)j * 64 : tmem_base + 128 + (uint32_t)j * 64);
          uint32_t ov[32];
          tmem_ld_32x32b_x32(tacc + lane_base + hh * 32, ov);
          tmem_ld_wait();
          tmem_ld_pin(ov);
          const uint32_t dst = sbase + (uint32_t)(3 * j + a) * kTileBytes + row * 128;
#pragma unroll
          for (int gq = 0; gq < 4; ++gq)
            sts_u4(dst + (((hh * 4 + gq) ^ (row & 7)) << 4),
                   pack_bf16x2(__uint_as_float(ov[gq * 8 + 0]), __uint_as_float(ov[gq * 8 + 1])),
                   pack_bf16x2(__uint_as_float(ov[gq * 8 + 2]), __uint_as_float(ov[gq * 8 + 3])),
                   pack_bf16x2(__uint_as_float(ov[gq * 8 + 4]), __uint_as_float(ov[gq * 8 + 5])),
                   pack_bf16x2(__uint_as_float(ov[gq * 8 + 6]), __uint_as_float(ov[gq * 8 + 7])));
        }
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc_free);
        mbar_arrive(epi_full);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int attn_small_bwd_launch(const void* qkv, const void* dout, const float* lse, const float* delta, int B, int S, int H,
                          float scale, void* dqkv, cudaStream_t st) {
  static const bool attr_ok = !(cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSbSmem) != cudaSuccess);  // once, thread-safe (C++11 static initialisation)
  if (!attr_ok) return BVC_ERR_LAUNCH;
  BVC_CHECK_ARG(S <= kSbMaxS);
  CUtensorMap tq, tq32, td, td32, tdq;
  if (make_head_tmap(&tq, qkv, 3 * H, S, B) || make_head_tmap(&tq32, qkv, 3 * H, S, B, 32) ||
      make_head_tmap(&td, dout, H, S, B) || make_head_tmap(&td32, dout, H, S, B, 32) ||
      make_head_tmap(&tdq, dqkv, 3 * H, S, B))
    return BVC_ERR_DRIVER;
  const long long n_work = (long long)B * H;
  BVC_CHECK_ARG(n_work < (1ll << 30));
  const int grid = (int)(n_work < num_sms() ? n_work : num_sms());
  attn_small_bwd_kernel<<<grid, kSmThreads, kSbSmem, st>>>(tq, tq32, td, td32, tdq, lse, delta, S, H, (int)n_work, scale);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

int attn_small_fwd_launch(const void* qkv, int B, int S, int H, float scale, void* out, float* lse, cudaStream_t st) {
  static const bool attr_ok = !(cudaFuncSetAttribute(attn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmSmem) != cudaSuccess);  // once, thread-safe (C++11 static initialisation)
  if (!attr_ok) return BVC_ERR_LAUNCH;
  CUtensorMap tm, to;
  int rc = make_head_tmap(&tm, qkv, 3 * H, S, B);
  if (rc) return rc;
  rc = make_head_tmap(&to, out, H, S, B);
  if (rc) return rc;
  const long long n_work = (long long)B * H;
  BVC_CHECK_ARG(n_work < (1ll << 30) && S <= kSmallMaxS);
  const int grid = (int)(n_work < num_sms() ? n_work : num_sms());
  attn_small_fwd_kernel<<<grid, kSmThreads, kSmSmem, st>>>(tm, to, lse, S, H, (int)n_work, scale * kLog2e);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

}  // namespace bvc
