// attn.cu -- multi-head self-attention (head_dim 64, no mask, no dropout) on tcgen05, forward and backward.
// Replaces HF VideoMAESelfAttention (HF:236-266) + sdpa/eager attention (HF:181-206) and its autograd backward.
//
// Layouts: qkv bf16 [B, S, 3, H, 64] (the fused-QKV GEMM output), out/dout bf16 [B, S, H*64], lse/delta fp32 [B,H,S].
// All operand tiles are 128 rows x 64 bf16 (128-byte rows, 128B swizzle) fetched by one 4-D TMA box; rows past the
// end of the sequence are zero-filled by TMA, columns past the end are masked in registers.
//
// Forward (persistent, one CTA per SM over (q-tile pair, head, clip) items; see the kernel):
//   warp 0 TMA loads, warp 1 MMA issuer, warp 3 TMA stores, warps 4-7 / 8-11 softmax of the pair's two query tiles.
//   S = Q K^T -> TMEM (128 cols); softmax threads read S, exp2 with a lazily updated running max (rescale O in TMEM
//   only when the max grows by > 8 in log2 units), write P (bf16, packed) back into TMEM as the A operand of
//   O += P V (tcgen05.mma with A from TMEM): P never touches shared memory, whose bandwidth (128 B/clk/SM, shared
//   by the UMMA operand reads, the TMA fills and st.shared) is what bounded the shared-memory-P version.
// Backward (two passes, no atomics, deterministic):
//   MODE_KV: CTA owns a K/V tile, streams Q/dO tiles, works on the TRANSPOSED tile (rows = keys):  S^T = K Q^T,
//            dP^T = V dO^T, P^T = exp2(S^T c - lse_q), dS^T = P^T (dP^T - delta_q) scale,  dV += P^T dO, dK += dS^T Q
//   MODE_Q : CTA owns a Q/dO tile, streams K/V tiles:  S = Q K^T, dP = dO V^T, same P / dS;  dQ += dS K.
//   In both passes P / dS go back to TMEM (packed bf16) and feed the accumulating MMAs as A-from-TMEM operands.
#include <stdlib.h>
#include <type_traits>

#include "attn_common.cuh"

namespace bvc {

// -DBVC_TRACE (tools/gpu_attn_trace.py builds a separate libbvc_trace.so): per-phase SM clock stamps of CTA 0's
// backward pipeline, read back through bvc_debug_trace_copy.  Not compiled into libbvc.so.
#ifdef BVC_TRACE
constexpr int kTraceRoles = 3, kTraceTiles = 256, kTracePoints = 8;
__device__ long long g_trace[kTraceRoles * kTraceTiles * kTracePoints];
#define BVC_TR(role, g, pt)                                                                              \
  do {                                                                                                   \
    if (blockIdx.x == 0 && (g) < kTraceTiles) g_trace[((role) * kTraceTiles + (g)) * kTracePoints + (pt)] = clock64(); \
  } while (0)
#else
#define BVC_TR(role, g, pt) \
  do {                      \
  } while (0)
#endif

// ================================================================================================ forward
// PERSISTENT, two query tiles in flight per CTA (one CTA per SM walks (query-tile PAIR, head, clip) work items):
//   warpgroup 0: warp 0 TMA loads (the pair's two Q tiles, double-buffered across items; K/V ring shared by both
//                tiles), warp 1 MMA issue, warp 3 TMA stores
//   warps 4-7  : softmax of query tile 0 of the pair,  warps 8-11: softmax of query tile 1
//                (thread = one query row x all 128 key columns of the K/V tile; TMEM lane quadrant = warp % 4)
// The two softmax groups are independent pipelines (own S, O, P TMEM regions and barriers) that share the tensor
// pipe and the K/V tiles in shared memory: while one group is in its MUFU-bound exponentials the other loads scores /
// stores probabilities, which is what keeps the MUFU pipe (the bound at head_dim 64: 1024 clk per 128x128 tile) busy.
// TMEM (512 columns): S0, S1 (128 each), O0, O1 (64 each), packed-bf16 P0, P1 (64 each; A operand of O += P V straight
// from TMEM: P never touches shared memory).  The running maximum is updated lazily (O is rescaled in TMEM only when
// the maximum grows by > 8 in log2 units).  O leaves through the finished item's Q buffers (128B-swizzled) and one
// TMA store per tile, which also clips the rows past the end of the sequence.
constexpr int kFwdThreads = 384;
constexpr int kFwdStages = 4;
constexpr int kFwdRegsCtl = 88, kFwdRegsCompute = 208;  // setmaxnreg split, as in the backward kernel
constexpr int kFwdSmem = kTileBytes * (4 + 2 * kFwdStages) + 256;

__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out,
                float* __restrict__ lse, int S, int H, int n_work, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                    // [2 item buffers][2 tiles of the pair]
  uint8_t* sK = smem + 4 * kTileBytes;                   // [stages]
  uint8_t* sV = smem + (4 + kFwdStages) * kTileBytes;    // [stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (4 + 2 * kFwdStages) * kTileBytes);
  uint64_t* q_full = bars + 0;                    // [2]
  uint64_t* q_empty = bars + 2;                   // [2]
  uint64_t* kv_full = bars + 4;                   // [stages]
  uint64_t* kv_empty = bars + 4 + kFwdStages;     // [stages]
  uint64_t* s_full = bars + 4 + 2 * kFwdStages;   // [2 groups]
  uint64_t* s_free = s_full + 2;                  // [2]
  uint64_t* p_full = s_full + 4;                  // [2]
  uint64_t* p_free = s_full + 6;                  // [2]  P V MMA of the group's tile complete (P reusable, O updated)
  uint64_t* epi_full = s_full + 8;
  uint64_t* o_free = s_full + 9;                  // [2] softmax group -> MMA warp: the item's O has been read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (S + kTile - 1) / kTile;       // K/V tiles per sequence
  const int n_pairs = (n_kv + 1) / 2;             // query-tile pairs per (clip, head)
  const int n_my = (n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_glob = n_my * n_kv;
  auto item_pair = [&](int k) { return ((int)blockIdx.x + k * (int)gridDim.x) % n_pairs; };
  auto item_bh = [&](int k) { return ((int)blockIdx.x + k * (int)gridDim.x) / n_pairs; };  // = b * H + h

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_out);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&p_free[i], 1);
    }
    for (int i = 0; i < kFwdStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(epi_full, 8);
    mbar_init(&o_free[0], 4);
    mbar_init(&o_free[1], 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // group t (query tile t of the pair): S at 128 t, O at 256 + 64 t, P at 384 + 64 t
  const uint32_t tS = tmem_base, tO = tmem_base + 256, tP = tmem_base + 384;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kFwdRegsCtl));
    if (warp == 0) {
      if (lane == 0) {
        int g = 0;
        for (int k = 0; k < n_my; ++k) {
          const int q0 = item_pair(k) * 2 * kTile, bh = item_bh(k), h = bh % H, b = bh / H;
          uint8_t* qb = sQ + (k & 1) * 2 * kTileBytes;
          mbar_wait(&q_empty[k & 1], ((uint32_t)(k >> 1) & 1u) ^ 1u);
          mbar_expect_tx(&q_full[k & 1], 2 * kTileBytes);
          tma_load_4d(qb, &tm_qkv, &q_full[k & 1], 0, h, q0, b);
          tma_load_4d(qb + kTileBytes, &tm_qkv, &q_full[k & 1], 0, h, q0 + kTile, b);  // may be entirely past S: zero fill
          for (int j = 0; j < n_kv; ++j, ++g) {
            const int st = g % kFwdStages;
            mbar_wait(&kv_empty[st], ((uint32_t)(g / kFwdStages) & 1u) ^ 1u);
            mbar_expect_tx(&kv_full[st], 2 * kTileBytes);
            tma_load_4d(sK + st * kTileBytes, &tm_qkv, &kv_full[st], 0, H + h, j * kTile, b);
            tma_load_4d(sV + st * kTileBytes, &tm_qkv, &kv_full[st], 0, 2 * H + h, j * kTile, b);
          }
        }
      }
    } else if (warp == 1) {
      // MMA issuer: warp-uniform control flow, one elected lane issues (see elect_one in bvc_ptx.cuh).  Flat loop over
      // the CTA's K/V tiles g (item k = g / n_kv, tile j = g % n_kv); per tile, for both groups: S of tile g + 1, then
      // P V of tile g.
      constexpr uint32_t idesc_o = umma_idesc_bf16(64, 0, 1, 128);
      // the last K/V tile of a sequence is ragged (1568 = 12 x 128 + 32; 160 = 128 + 32): size its MMAs to the valid
      // rows rounded up to 16 (N of S = Q K^T, K-steps of O += P V) instead of paying for a full 128
      auto valid16 = [&](int j) { return min(kTile, (S - j * kTile + 15) & ~15); };
      auto issue_s = [&](int k, int j, int st, int t) {
        const uint64_t dQ = desc_k(smem_u32(sQ + ((k & 1) * 2 + t) * kTileBytes), 0);
        const uint64_t dK = desc_k(smem_u32(sK + st * kTileBytes), 0);
        const uint32_t idesc_s = umma_idesc_bf16(valid16(j), 0, 0, 128);
        const uint32_t d = tS + (uint32_t)t * 128;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(d, dQ + 2 * kk, dK + 2 * kk, idesc_s, kk > 0);
        umma_commit(&s_full[t]);
      };
      if (n_glob > 0) {
        mbar_wait(&q_full[0], 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        if (elect_one()) {
          issue_s(0, 0, 0, 0);
          issue_s(0, 0, 0, 1);
        }
        __syncwarp();
      }
      auto issue_pv = [&](int j, int cst, int t) {
        // O_t += P_t V : A = P (TMEM, packed bf16, 8 columns per K = 16 step), B = V_j (MN-major: n = d, k = kv)
        const int ksteps = valid16(j) >> 4;
        const uint64_t dV = desc_mn(smem_u32(sV + cst * kTileBytes), 0, 8192);
        const uint32_t aP = tP + (uint32_t)t * 64, dO = tO + (uint32_t)t * 64;
        const uint32_t acc0 = j > 0;
        if (ksteps == 8) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_bf16_ts(dO, aP + kk * 8, dV + 128 * kk, idesc_o, acc0 | (kk > 0));
        } else {
          for (int kk = 0; kk < ksteps; ++kk) umma_bf16_ts(dO, aP + kk * 8, dV + 128 * kk, idesc_o, acc0 | (kk > 0));
        }
        umma_commit(&p_free[t]);
        if (t == 1) umma_commit(&kv_empty[cst]);  // group 1's P V is the last reader of the stage
      };
      // Issue order per K/V tile g: S0(g+1), PV0(g), S1(g+1), PV1(g).  (The rotated order S0, PV1(g-1), S1, PV0 -- the
      // arrival order of the events when the groups run half a tile out of phase -- measured 8 % slower; one issuing
      // warp PER GROUP, so that neither group's P V waits behind the other group's barriers: 406 us against 383-404.)
      int k = 0, j = 0;
      for (int g = 0; g < n_glob; ++g) {
        int k1 = k, j1 = j + 1;
        if (j1 == n_kv) {
          j1 = 0;
          ++k1;
        }
        const int st1 = (g + 1) % kFwdStages, cst = g % kFwdStages;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (g + 1 < n_glob) {
            if (t == 0) {
              if (j1 == 0) mbar_wait(&q_full[k1 & 1], (uint32_t)(k1 >> 1) & 1u);
              mbar_wait(&kv_full[st1], (uint32_t)((g + 1) / kFwdStages) & 1u);
            }
            mbar_wait(&s_free[t], (uint32_t)g & 1u);  // group t has its scores of tile g in registers
            tc_fence_after();
            if (elect_one()) issue_s(k1, j1, st1, t);
            __syncwarp();
          }
          mbar_wait(&p_full[t], (uint32_t)g & 1u);
          // a new item's first P V overwrites O: the group's epilogue must have read the previous item's O
          if (j == 0 && k > 0) mbar_wait(&o_free[t], (uint32_t)(k - 1) & 1u);
          tc_fence_after();
          if (elect_one()) issue_pv(j, cst, t);
          __syncwarp();
        }
        k = k1;
        j = j1;
      }
    } else if (warp == 3) {
      // store warp: one TMA store per staged O tile, then the item's Q buffers go back to the TMA warp
      for (int k = 0; k < n_my; ++k) {
        const int q0 = item_pair(k) * 2 * kTile, bh = item_bh(k), h = bh % H, b = bh / H;
        uint8_t* qb = sQ + (k & 1) * 2 * kTileBytes;
        mbar_wait(epi_full, (uint32_t)k & 1u);
        if (lane == 0) {
          tma_store_4d(&tm_out, qb, 0, h, q0, b);
          if (q0 + kTile < S) tma_store_4d(&tm_out, qb + kTileBytes, 0, h, q0 + kTile, b);
          tma_store_commit();
          tma_store_wait_read0();
          mbar_arrive(&q_empty[k & 1]);
        }
        __syncwarp();
      }
      if (lane == 0) tma_store_wait0();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kFwdRegsCompute));
    const int q4 = warp & 3;
    const int t = (warp - 4) >> 2;  // which query tile of the pair this group owns
    const int row = q4 * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const uint32_t mS = tS + (uint32_t)t * 128 + lane_base, mO = tO + (uint32_t)t * 64 + lane_base;
    const uint32_t mP = tP + (uint32_t)t * 64 + lane_base;
    const uint64_t c2 = pack2(scale_log2, scale_log2);
    int g = 0;
    for (int k = 0; k < n_my; ++k) {
      const int q0 = (item_pair(k) * 2 + t) * kTile, bh = item_bh(k);
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < n_kv; ++j, ++g) {
        const int valid = S - j * kTile;  // columns >= valid are past the end of the sequence (ragged last tile)
        const bool tr = lane == 0 && q4 == 0;
        if (tr) BVC_TR(t, g, 0);
        mbar_wait(&s_full[t], (uint32_t)g & 1u);
        if (tr) BVC_TR(t, g, 1);
        tc_fence_after();
        uint32_t sv[4][32];  // the whole score row of this tile
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(mS + c * 32, sv[c]);
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_wait_pin(sv[c]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);  // the MMA warp may overwrite S with the next tile's scores
        if (tr) BVC_TR(t, g, 2);
        float mx = -INFINITY;
        if (valid >= kTile) {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(sv[c][i]));
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(sv[c][i]));
        }
        mx *= scale_log2;  // raw scores -> log2 domain (scale_log2 > 0)
        const bool need = mx > m_used + 8.0f;
        float alpha = 1.0f;
        if (need) {
          alpha = exp2f(m_used - mx);  // 0 on the first tile
          m_used = mx;
        }
        if (j > 0) {
          // P V of the previous tile: P is reusable and O is up to date (needed for the rescale)
          mbar_wait(&p_free[t], (uint32_t)(g - 1) & 1u);
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t ov[32];
              tmem_ld_32x32b_x32(mO + c * 32, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
              tmem_st_32x32b_x32(mO + c * 32, ov);
            }
          }
        }
        l *= alpha;
        if (tr) BVC_TR(t, g, 3);
        // P = exp2(S * scale_log2 - m), row sum, packed bf16 P back into TMEM (32 columns of P per 64 scores)
        const uint64_t m2 = pack2(-m_used, -m_used);
        uint64_t lsum2 = pack2(0.f, 0.f);
        auto softmax_row = [&](auto tail_tag) {
          constexpr bool kTail = decltype(tail_tag)::value;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t pk[32];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const uint32_t(&s32)[32] = sv[hh * 2 + c];
                const uint64_t x2 = ffma2(pack2(__uint_as_float(s32[i]), __uint_as_float(s32[i + 1])), c2, m2);
                float p0, p1;
                unpack2(exp2_mufu2(x2), p0, p1);
                if (kTail) {
                  const int col = hh * 64 + c * 32 + i;
                  if (col >= valid) p0 = 0.f;
                  if (col + 1 >= valid) p1 = 0.f;
                }
                lsum2 = fadd2(lsum2, pack2(p0, p1));
                pk[c * 16 + (i >> 1)] = pack_bf16x2(p0, p1);
              }
            tmem_st_32x32b_x32(mP + hh * 32, pk);
          }
        };
        if (valid >= kTile) softmax_row(std::false_type{});
        else softmax_row(std::true_type{});
        float ls0, ls1;
        unpack2(lsum2, ls0, ls1);
        l += ls0 + ls1;
        if (tr) BVC_TR(t, g, 4);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        if (tr) BVC_TR(t, g, 5);
      }
      // item epilogue: O / l -> bf16 -> this group's Q buffer of the item (128B-swizzled) -> TMA store (store warp)
      mbar_wait(&p_free[t], (uint32_t)(g - 1) & 1u);  // the item's last P V
      if (lane == 0 && q4 == 0) BVC_TR(t, g - 1, 6);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      uint8_t* stage = sQ + ((k & 1) * 2 + t) * kTileBytes;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(mO + c * 32, ov);
        tmem_ld_wait();
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
      if (q0 + row < S) lse[(long long)bh * S + q0 + row] = (m_used + log2f(l)) * 0.6931471805599453f;
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&o_free[t]);
        mbar_arrive(epi_full);
      }
      if (lane == 0 && q4 == 0) BVC_TR(t, g - 1, 7);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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

// warpgroup 0: warp 0 TMA loads, warp 1 MMA issue, warp 2 statistics (MODE_KV), warp 3 TMA stores; warpgroups 1-2
// (warps 4-11): compute.  Launched at the 168
// registers/thread that 384 threads allow, then warpgroup 0 shrinks to 88 and the compute warpgroups grow to 208
// (setmaxnreg): the compute threads hold a 64-wide slice of S and dP plus their packed P / dS outputs in registers.
constexpr int kBwdThreads = 384;
constexpr int kBwdRegsCtl = 88, kBwdRegsCompute = 208;
// PERSISTENT: one CTA per SM walks a list of work items (owned tile, head, clip).  Everything is pipelined ACROSS
// items through the same barriers that pipeline the tiles of one item -- the TMA warp runs ahead into the next item's
// resident and streamed tiles, the MMA warp issues the next item's first S / dP while the compute warps store the
// previous item's accumulators -- so the per-CTA costs of the one-CTA-per-item version (launch, barrier init, TMEM
// allocation, the DRAM latency of the first tiles, the epilogue; ~8.7k clk per item, measured: 25 % of an S = 1568 item
// and 80 % of an S = 160 item) are paid once per SM or hidden.
// smem: resident pair x 2 buffers + stream ring (4 stages x 2 tiles) + one "augmentation" tile + barriers.
// P / dS never touch shared memory: they go back into TMEM (packed bf16) as the A operands of the accumulating MMAs.
//
// MODE_KV works on the transposed tile, where the softmax statistics (lse, delta) vary along the COLUMNS a thread
// holds.  Fetching them per column (staged in shared memory, one barrier per tile) doubled the math phase (trace:
// 2021 clk vs 1041 without), so the subtraction is folded into the score MMAs instead: one extra K = 16 step with
//   A = [1 1 1 0 ...]  (rows = keys)      B = three-way bf16 split of -lse/scale (resp. -delta)  (rows = queries)
// makes the tensor core deliver S - lse/scale and dP - delta directly (the split carries 24 mantissa bits, the
// products with 1.0 are exact, accumulation is fp32).  The three [128][16] operands live in k-steps 0 / 1 / 2 of one
// 128B-swizzled [128][64] tile; the compute warps rewrite the two statistic columns once per streamed tile.
constexpr int kBwdStages = 4;
constexpr int kBwdSmem = kTileBytes * (4 + 2 * kBwdStages + 2) + 256;

template <int MODE_KV>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dqkv, const float* __restrict__ lse,
                const float* __restrict__ delta, int S, int H, int n_work, float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sR = smem;                                      // [2 buffers][R0 | R1]; MODE_KV: K_j, V_j | MODE_Q: Q_i, dO_i
  uint8_t* sX = smem + 4 * kTileBytes;                     // [stages] MODE_KV: Q_i  | MODE_Q: K_j
  uint8_t* sY = smem + (4 + kBwdStages) * kTileBytes;      // [stages] MODE_KV: dO_i | MODE_Q: V_j
  uint8_t* sAug = smem + (4 + 2 * kBwdStages) * kTileBytes;  // MODE_KV: [2] x (ones | -lse/scale | -delta in k-steps 0, 1, 2)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (6 + 2 * kBwdStages) * kTileBytes);
  uint64_t* r_full = bars + 0;                   // [2]
  uint64_t* r_empty = bars + 2;                  // [2]
  uint64_t* st_full = bars + 4;                  // [stages]
  uint64_t* st_empty = bars + 4 + kBwdStages;    // [stages]
  uint64_t* sdp_full = bars + 4 + 2 * kBwdStages;
  uint64_t* sdp_free = sdp_full + 1;
  uint64_t* pds_full = sdp_full + 2;
  uint64_t* pds_free = sdp_full + 3;
  uint64_t* aug_full = sdp_full + 4;             // [2]  statistics warp -> MMA warp
  uint64_t* aug_empty = sdp_full + 6;            // [2]
  uint64_t* epi_full = sdp_full + 8;             // compute warps -> store warp: an item's accumulators are staged
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sdp_full + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_it = (S + kTile - 1) / kTile;  // tiles per sequence = owned tiles per (clip, head) = streamed tiles per item
  // work item w -> (owned tile, head, clip), owned tile fastest: the CTAs running at any moment share a few (clip,
  // head) pairs, so the streamed tiles are served from L2
  const int n_my = (n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_glob = n_my * n_it;            // streamed tiles this CTA processes, over all its items
  auto item_tile = [&](int k) { return ((int)blockIdx.x + k * (int)gridDim.x) % n_it; };
  auto item_bh = [&](int k) { return ((int)blockIdx.x + k * (int)gridDim.x) / n_it; };  // = b * H + h

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dqkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_empty[i], 1);
    }
    for (int i = 0; i < kBwdStages; ++i) {
      mbar_init(&st_full[i], 1);
      mbar_init(&st_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 8);
    mbar_init(pds_full, 8);
    mbar_init(pds_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&aug_full[i], 1);
      mbar_init(&aug_empty[i], 1);
    }
    mbar_init(epi_full, 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // fp32 [128 lanes][128]: scores and dP; fp32 [128][64]: accumulators; packed bf16 [128][128] = 64 columns: P, dS
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tA0 = tmem_base + 256, tA1 = tmem_base + 320;
  const uint32_t tP = tmem_base + 384, tDS = tmem_base + 448;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kBwdRegsCtl));
    if (warp == 0) {
      if (lane == 0) {
        int g = 0;
        for (int k = 0; k < n_my; ++k) {
          const int own0 = item_tile(k) * kTile, bh = item_bh(k), h = bh % H, b = bh / H;
          uint8_t* r0 = sR + (k & 1) * 2 * kTileBytes;
          mbar_wait(&r_empty[k & 1], ((uint32_t)(k >> 1) & 1u) ^ 1u);
          mbar_expect_tx(&r_full[k & 1], 2 * kTileBytes);
          if (MODE_KV) {
            tma_load_4d(r0, &tm_qkv, &r_full[k & 1], 0, H + h, own0, b);
            tma_load_4d(r0 + kTileBytes, &tm_qkv, &r_full[k & 1], 0, 2 * H + h, own0, b);
          } else {
            tma_load_4d(r0, &tm_qkv, &r_full[k & 1], 0, h, own0, b);
            tma_load_4d(r0 + kTileBytes, &tm_do, &r_full[k & 1], 0, h, own0, b);
          }
          for (int i = 0; i < n_it; ++i, ++g) {
            const int st = g % kBwdStages;
            mbar_wait(&st_empty[st], ((uint32_t)(g / kBwdStages) & 1u) ^ 1u);
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
      }
    } else if (warp == 1) {
      // MMA issuer: the whole warp walks the (uniform) control flow and polls the barriers, one elected lane issues.
      // One flat loop over the CTA's streamed tiles g (item k = g / n_it, tile i = g % n_it): S / dP of tile g + 1 --
      // possibly the next item's first tile -- is issued before the accumulating MMAs of tile g.
      constexpr uint32_t idesc_acc = umma_idesc_bf16(64, 0, 1, 128);  // A from TMEM (K-major), B MN-major, N = 64
      // ragged last tile of the STREAMED operand: it is the N of the score MMAs in both modes (MODE_KV streams the
      // query tiles and computes S^T = K Q^T; MODE_Q streams the K/V tiles and computes S = Q K^T) and the reduction
      // dimension of the accumulating MMAs, so both shrink to its valid rows rounded up to 16
      auto valid16 = [&](int i) { return min(kTile, (S - i * kTile + 15) & ~15); };
      // k = 0 descriptors; a K = 16 step moves the start address by 32 B (K-major, +2 in the descriptor's 16-byte
      // units) or by 16 rows = 2048 B (MN-major, +128)
      auto issue_s_dp = [&](int k, int i, int st, int gg) {
        const uint64_t dAug = desc_k(smem_u32(sAug + (gg & 1) * kTileBytes), 0);
        const uint32_t idesc_s = umma_idesc_bf16(valid16(i), 0, 0, 128);
        const uint32_t aR = smem_u32(sR + (k & 1) * 2 * kTileBytes);
        const uint64_t dR0 = desc_k(aR, 0), dR1 = desc_k(aR + kTileBytes, 0);
        const uint64_t dX = desc_k(smem_u32(sX + st * kTileBytes), 0), dY = desc_k(smem_u32(sY + st * kTileBytes), 0);
        // MODE_KV: S^T = K Q^T, dP^T = V dO^T (rows = kv, cols = q);  MODE_Q: S = Q K^T, dP = dO V^T (rows = q)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tS, dR0 + 2 * kk, dX + 2 * kk, idesc_s, kk > 0);
        if (MODE_KV) umma_bf16_ss(tS, dAug, dAug + 2, idesc_s, 1);       // S^T - lse_q / scale
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tDP, dR1 + 2 * kk, dY + 2 * kk, idesc_s, kk > 0);
        if (MODE_KV) {
          umma_bf16_ss(tDP, dAug, dAug + 4, idesc_s, 1);                 // dP^T - delta_q
          umma_commit(&aug_empty[gg & 1]);
        }
      };
      if (n_glob > 0) {
        mbar_wait(&r_full[0], 0);
        mbar_wait(&st_full[0], 0);
        if (MODE_KV) mbar_wait(&aug_full[0], 0);
        tc_fence_after();
        if (elect_one()) {
          issue_s_dp(0, 0, 0, 0);
          umma_commit(sdp_full);
        }
        __syncwarp();
      }
      int k = 0, i = 0;  // item / tile of g
      for (int g = 0; g < n_glob; ++g) {
        int k1 = k, i1 = i + 1;  // item / tile of g + 1
        if (i1 == n_it) {
          i1 = 0;
          ++k1;
        }
        if (lane == 0) BVC_TR(2, g, 0);
        if (g + 1 < n_glob) {
          const int st1 = (g + 1) % kBwdStages;
          if (i1 == 0) mbar_wait(&r_full[k1 & 1], (uint32_t)(k1 >> 1) & 1u);
          mbar_wait(&st_full[st1], (uint32_t)((g + 1) / kBwdStages) & 1u);
          if (MODE_KV) mbar_wait(&aug_full[(g + 1) & 1], (uint32_t)((g + 1) >> 1) & 1u);
          if (lane == 0) BVC_TR(2, g, 1);
          mbar_wait(sdp_free, (uint32_t)g & 1u);
          if (lane == 0) BVC_TR(2, g, 2);
          tc_fence_after();
          if (elect_one()) {
            issue_s_dp(k1, i1, st1, g + 1);
            umma_commit(sdp_full);
          }
          __syncwarp();
        }
        if (lane == 0) BVC_TR(2, g, 3);
        mbar_wait(pds_full, (uint32_t)g & 1u);
        if (lane == 0) BVC_TR(2, g, 4);
        tc_fence_after();
        const int cst = g % kBwdStages;
        const int ksteps = valid16(i) >> 4;  // streamed tile = the reduction dimension in both modes
        if (elect_one()) {
          const uint64_t dX = desc_mn(smem_u32(sX + cst * kTileBytes), 0, 8192);
          const uint64_t dY = desc_mn(smem_u32(sY + cst * kTileBytes), 0, 8192);
          const uint32_t acc0 = i > 0;
          if (MODE_KV) {
            // dV[kv, d] += P^T dO : A = P^T (TMEM, m = kv, k = q), B = dO_i (MN-major: n = d, k = q)
            // dK[kv, d] += dS^T Q : A = dS^T (TMEM), B = Q_i (MN-major)
            if (ksteps == 8) {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) umma_bf16_ts(tA0, tP + kk * 8, dY + 128 * kk, idesc_acc, acc0 | (kk > 0));
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) umma_bf16_ts(tA1, tDS + kk * 8, dX + 128 * kk, idesc_acc, acc0 | (kk > 0));
            } else {
              for (int kk = 0; kk < ksteps; ++kk)
                umma_bf16_ts(tA0, tP + kk * 8, dY + 128 * kk, idesc_acc, acc0 | (kk > 0));
              for (int kk = 0; kk < ksteps; ++kk)
                umma_bf16_ts(tA1, tDS + kk * 8, dX + 128 * kk, idesc_acc, acc0 | (kk > 0));
            }
          } else {
            // dQ[q, d] += dS K : A = dS (TMEM, m = q, k = kv), B = K_j (MN-major: n = d, k = kv)
            if (ksteps == 8) {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) umma_bf16_ts(tA0, tDS + kk * 8, dX + 128 * kk, idesc_acc, acc0 | (kk > 0));
            } else {
              for (int kk = 0; kk < ksteps; ++kk)
                umma_bf16_ts(tA0, tDS + kk * 8, dX + 128 * kk, idesc_acc, acc0 | (kk > 0));
            }
          }
          umma_commit(pds_free);
          umma_commit(&st_empty[cst]);
        }
        __syncwarp();
        if (lane == 0) BVC_TR(2, g, 5);
        k = k1;
        i = i1;
      }
    } else if (warp == 2) {
      // statistics warp (MODE_KV): per streamed tile, the three-way bf16 split of -lse/scale and -delta of its 128
      // queries into k-steps 1 / 2 (16-byte chunks 2 / 4 of the 128B-swizzled rows) of the augmentation tile g & 1;
      // k-step 0 holds the constant ones.  Lane l owns queries 4l .. 4l+3.
      if (MODE_KV) {
        const uint4 ones = make_uint4(pack_bf16x2(1.f, 1.f), pack_bf16x2(1.f, 0.f), 0u, 0u);
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        for (int r = lane; r < 2 * kTile; r += 32) {  // both tiles are contiguous: 256 rows of 128 bytes
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c != 2 && c != 4) *reinterpret_cast<uint4*>(sAug + r * 128 + ((c ^ (r & 7)) << 4)) = c == 0 ? ones : zero;
        }
        float cur[8], nxt[8];
        auto load8 = [&](int kk, int ii, float (&v)[8]) {
          const float* lb = lse + (long long)item_bh(kk) * S;
          const float* db = delta + (long long)item_bh(kk) * S;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int qg = min(ii * kTile + lane * 4 + j, S - 1);
            v[j] = __ldg(lb + qg);
            v[4 + j] = __ldg(db + qg);
          }
        };
        int k = 0, i = 0;
        if (n_glob > 0) load8(0, 0, cur);
        for (int g = 0; g < n_glob; ++g) {
          int k1 = k, i1 = i + 1;
          if (i1 == n_it) {
            i1 = 0;
            ++k1;
          }
          if (g + 1 < n_glob) load8(k1, i1, nxt);  // in flight while this tile is written
          mbar_wait(&aug_empty[g & 1], ((uint32_t)(g >> 1) & 1u) ^ 1u);
          uint8_t* tile = sAug + (g & 1) * kTileBytes;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = lane * 4 + (j & 3);
            const float v = j < 4 ? -cur[j] / scale : -cur[j];
            const float hi = __bfloat162float(__float2bfloat16_rn(v));
            const float r1 = v - hi;
            const float mid = __bfloat162float(__float2bfloat16_rn(r1));
            const float lo = r1 - mid;
            *reinterpret_cast<uint4*>(tile + r * 128 + (((j < 4 ? 2 : 4) ^ (r & 7)) << 4)) =
                make_uint4(pack_bf16x2(hi, mid), pack_bf16x2(lo, 0.f), 0u, 0u);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&aug_full[g & 1]);
#pragma unroll
          for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
          k = k1;
          i = i1;
        }
      }
    } else {
      // store warp: one TMA store per staged accumulator (it also clips the rows past the end of the sequence), then
      // the resident-tile buffer that served as staging goes back to the TMA warp
      for (int k = 0; k < n_my; ++k) {
        const int own0 = item_tile(k) * kTile, bh = item_bh(k), h = bh % H, b = bh / H;
        uint8_t* stage = sR + (k & 1) * 2 * kTileBytes;
        mbar_wait(epi_full, (uint32_t)k & 1u);
        if (lane == 0) {
          if (MODE_KV) {
            tma_store_4d(&tm_dqkv, stage, 0, 2 * H + h, own0, b);               // dV
            tma_store_4d(&tm_dqkv, stage + kTileBytes, 0, H + h, own0, b);      // dK
          } else {
            tma_store_4d(&tm_dqkv, stage, 0, h, own0, b);                       // dQ
          }
          tma_store_commit();
          tma_store_wait_read0();
          mbar_arrive(&r_empty[k & 1]);  // the TMA warp may refill this buffer with the resident tiles of item k + 2
        }
        __syncwarp();
      }
      if (lane == 0) tma_store_wait0();  // global writes complete before the CTA exits
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kBwdRegsCompute));
    // 8 compute warps: TMEM lane quadrant q4 = warp % 4 (rows q4*32 .. +31 of the score tile), column half = (warp-4)/4.
    // MODE_Q : row = query, columns = keys;   statistics (lse, delta) are per ROW: two registers per thread.
    // MODE_KV: row = key,   columns = queries; the statistics were subtracted by the augmented score MMAs.
    // No masking anywhere: rows / columns past the end of the sequence meet zero-filled operand rows in every MMA
    // that consumes them (or land in accumulator rows that the TMA store clips); they only have to stay finite,
    // which the clamped statistics loads guarantee.
    const int e = warp - 4;
    const int q4 = warp & 3;
    const int half = e >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const float c_log2 = scale * kLog2e;
    const uint64_t cl2 = pack2(c_log2, c_log2), sc2 = pack2(scale, scale);
    // MODE_Q: this thread's row statistics of item kk
    auto load_row = [&](int kk, float& l, float& d) {
      const int qc = min(item_tile(kk) * kTile + row, S - 1);
      const long long o = (long long)item_bh(kk) * S + qc;
      l = __ldg(lse + o);
      d = __ldg(delta + o);
    };
    float l_cur = 0.f, d_cur = 0.f;
    if (!MODE_KV && n_glob > 0) load_row(0, l_cur, d_cur);
    int g = 0;
    for (int k = 0; k < n_my; ++k) {
      uint64_t nl2 = 0, nd2 = 0;
      float l_nxt = 0.f, d_nxt = 0.f;
      if (!MODE_KV) {
        const float l = -l_cur * kLog2e, d = -d_cur * scale;
        nl2 = pack2(l, l);
        nd2 = pack2(d, d);
        if (k + 1 < n_my) load_row(k + 1, l_nxt, d_nxt);  // in flight during this whole item
      }
      for (int i = 0; i < n_it; ++i, ++g) {
        const bool tr = lane == 0 && q4 == 0;
        if (tr) BVC_TR(half, g, 0);
        mbar_wait(sdp_full, (uint32_t)g & 1u);
        if (tr) BVC_TR(half, g, 1);
        tc_fence_after();
        uint32_t sv[2][32], dv[2][32];
#pragma unroll
        for (int c = 0; c < 2; ++c) tmem_ld_32x32b_x32(tS + lane_base + half * 64 + c * 32, sv[c]);
#pragma unroll
        for (int c = 0; c < 2; ++c) tmem_ld_32x32b_x32(tDP + lane_base + half * 64 + c * 32, dv[c]);
        // all four loads landed before the TMEM columns are handed back (the math below must not be hoisted between
        // the loads: that delays sdp_free, and with it the next tile's S / dP MMAs, by half a tile of MUFU work)
        tmem_ld_wait_pin(sv[0]);
        tmem_ld_wait_pin(sv[1]);
        tmem_ld_wait_pin(dv[0]);
        tmem_ld_wait_pin(dv[1]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_free);
        if (tr) BVC_TR(half, g, 2);
        // this thread's 64 P / dS values of the tile, packed bf16x2; the math overlaps the previous tile's
        // accumulating MMAs, which are still reading the TMEM P / dS operands
        uint32_t pk[32], dk[32];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const uint64_t s2 = pack2(__uint_as_float(sv[c][j]), __uint_as_float(sv[c][j + 1]));
            const uint64_t p2 = pack2(__uint_as_float(dv[c][j]), __uint_as_float(dv[c][j + 1]));
            // P = exp2((S - lse/scale) * scale * log2e);  dS = P * (dP - delta) * scale
            const uint64_t x2 = MODE_KV ? fmul2(s2, cl2) : ffma2(s2, cl2, nl2);
            const uint64_t y2 = MODE_KV ? fmul2(p2, sc2) : ffma2(p2, sc2, nd2);
            const uint64_t e2 = exp2_mufu2(x2);
            float p0, p1, d0, d1;
            unpack2(e2, p0, p1);
            unpack2(fmul2(e2, y2), d0, d1);
            if (MODE_KV) pk[c * 16 + (j >> 1)] = pack_bf16x2(p0, p1);
            dk[c * 16 + (j >> 1)] = pack_bf16x2(d0, d1);
          }
        if (tr) BVC_TR(half, g, 3);
        if (g > 0) {
          mbar_wait(pds_free, (uint32_t)(g - 1) & 1u);  // previous dV/dK/dQ MMAs finished reading P / dS
          tc_fence_after();
        }
        if (tr) BVC_TR(half, g, 4);
        if (MODE_KV) tmem_st_32x32b_x32(tP + lane_base + half * 32, pk);
        tmem_st_32x32b_x32(tDS + lane_base + half * 32, dk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_full);
        if (tr) BVC_TR(half, g, 5);
      }
      // item epilogue: accumulators (rows = owned tile rows, 64 columns) -> bf16 -> the finished item's resident-tile
      // buffer (no MMA reads it any more) in the 128B-swizzled layout -> ONE TMA store per accumulator, which also
      // clips the rows past the end of the sequence.  (Direct st.global from the one-row-per-lane TMEM layout touches
      // 32 cache lines per instruction: trace 2500-3400 clk per item.)  The MMA warp is already issuing the next item's
      // first S / dP; its first accumulating MMA (which overwrites the accumulators) waits for this warp's next
      // pds_full arrival, i.e. until after these loads.
      mbar_wait(pds_free, (uint32_t)(g - 1) & 1u);
      if (lane == 0 && q4 == 0) BVC_TR(half, g - 1, 6);
      tc_fence_after();
      uint8_t* stage = sR + (k & 1) * 2 * kTileBytes;
      auto stage32 = [&](uint32_t tacc, uint8_t* dst_tile) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(tacc + lane_base + half * 32, ov);
        tmem_ld_wait();
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          uint4 o4;
          o4.x = pack_bf16x2(__uint_as_float(ov[gq * 8 + 0]), __uint_as_float(ov[gq * 8 + 1]));
          o4.y = pack_bf16x2(__uint_as_float(ov[gq * 8 + 2]), __uint_as_float(ov[gq * 8 + 3]));
          o4.z = pack_bf16x2(__uint_as_float(ov[gq * 8 + 4]), __uint_as_float(ov[gq * 8 + 5]));
          o4.w = pack_bf16x2(__uint_as_float(ov[gq * 8 + 6]), __uint_as_float(ov[gq * 8 + 7]));
          *reinterpret_cast<uint4*>(dst_tile + row * 128 + (((half * 4 + gq) ^ (row & 7)) << 4)) = o4;
        }
      };
      stage32(tA0, stage);
      if (MODE_KV) stage32(tA1, stage + kTileBytes);
      tc_fence_before();  // the accumulator loads are ordered before this warp's next pds_full arrival
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(epi_full);  // the store warp takes it from here
      if (lane == 0 && q4 == 0) BVC_TR(half, g - 1, 7);
      l_cur = l_nxt;
      d_cur = d_nxt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ================================================================================================ backward, ONE pass
// attn_bwd1_kernel: the backward in a single pass over the (key tile, query tile) pairs -- 5 MMAs and ONE exponential
// per score where the two-pass kernels above execute 7 and 2 (they recompute S and dP in each pass).  The CTA owns a
// K/V tile like MODE_KV (transposed tile: rows = keys; statistics folded into the score MMAs through the augmentation
// tile) and streams the Q / dO tiles; in addition to dV += P^T dO and dK += dS^T Q every streamed tile yields this
// pair's contribution  dQ_i = dS K_j,  which leaves through shared memory and a TMA REDUCE (fp32 add in L2) into an
// fp32 accumulation buffer [B, S, H, 64]; bvc_attn_bwd zero-fills that buffer before and converts it into the q slot of
// dqkv after the kernel.
//   dS^T (bf16) is written ONCE, to shared memory, as two 128B-swizzled [128 keys][64 queries] atoms, and read with
//   both operand majors: K-major as the A operand of dK += dS^T Q (M = keys, K = queries) and MN-major -- transposed
//   by the UMMA descriptor -- as the A operand of dQ = dS K (M = queries, K = keys).  P^T still goes back to TMEM.
// TMEM (512 columns): S^T 128 | dP^T 128 | dV 64 | dK 64 | P^T (packed bf16) 64 | dQ 64.
// Shared memory (14 tiles of 16 KB): K_j, V_j x 2 item buffers | Q_i, dO_i x 3 stages (with two the load of tile g + 2
// could only start when tile g's accumulating MMAs had completed, and its ~1300 clk latency sat on the critical path of
// every tile) | ONE augmentation tile (16-byte chunks of a row: 0 = the A operand of the S^T step [1,1,1,0..], 2 = the A
// operand of the dP^T step [0,0,0,1,1,1,0,0], 4 / 6 = the B operand [lse hi,mid,lo, delta hi,mid,lo, 0,0] of stage 0 / 1;
// odd chunks stay zero) | dS^T (2 atoms) | dQ staging (ONE fp32 [128][32] tile, used twice per streamed tile: columns
// 0-31 by compute warps 0-3, then columns 32-63 by warps 4-7 once the first round's reduce has read the tile).
// (Measured alternatives: staging the dQ tile in its own, dead, Q / dO stage frees the stage too late -- the next
// load's latency is exposed every other tile, 1030 us; deferring the second round into the next tile's iteration makes
// the next first round wait for it, 885 us; staging columns 0-31 before the P^T / dS^T stores puts it on the MMA warp's
// critical path, 796-836 us against 790.)
// Per streamed tile the compute warps: read S^T / dP^T (TMEM), exp + dS math, drain the PREVIOUS tile's dQ accumulator
// into the staging buffer (the store warp reduces it into global memory), write P^T (TMEM) and dS^T (smem).
constexpr int kB1Stages = 3;

constexpr int kB1Smem = kTileBytes * 14 + 256;
static_assert(kB1Smem <= 232448, "single-pass backward: shared memory");

__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                                                  int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd1_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                 const __grid_constant__ CUtensorMap tm_dqkv, const __grid_constant__ CUtensorMap tm_dq,
                 const float* __restrict__ lse, const float* __restrict__ delta, int S, int H, int n_work, float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sR = smem;                                   // [2 buffers][K_j | V_j]
  uint8_t* sX = smem + 4 * kTileBytes;                  // [stages] Q_i
  uint8_t* sY = smem + (4 + kB1Stages) * kTileBytes;    // [stages] dO_i
  uint8_t* sAug = smem + (4 + 2 * kB1Stages) * kTileBytes;      // one tile, both stages (chunk layout above)
  uint8_t* sDS = smem + (5 + 2 * kB1Stages) * kTileBytes;       // dS^T: 2 atoms of [128 keys][64 queries]
  uint8_t* sDQ = smem + (7 + 2 * kB1Stages) * kTileBytes;       // dQ staging: fp32 [128 queries][32], two rounds per tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (8 + 2 * kB1Stages) * kTileBytes);
  uint64_t* r_full = bars + 0;                   // [2]
  uint64_t* r_empty = bars + 2;                  // [2]
  uint64_t* st_full = bars + 4;                  // [stages]
  uint64_t* st_empty = bars + 4 + kB1Stages;     // [stages]
  uint64_t* sdp_full = bars + 4 + 2 * kB1Stages;
  uint64_t* sdp_free = sdp_full + 1;
  uint64_t* pds_full = sdp_full + 2;             // compute -> MMA: P^T in TMEM, dS^T in smem, previous dQ drained
  uint64_t* acc_done = sdp_full + 3;             // MMA (commit) -> compute: dV / dK / dQ MMAs of a tile complete
  uint64_t* aug_full = sdp_full + 4;             // [2]
  uint64_t* aug_empty = sdp_full + 6;            // [2]
  uint64_t* epi_full = sdp_full + 8;             // compute -> store warp: an item's dV / dK are staged
  uint64_t* dq_staged = sdp_full + 9;            // compute -> store warp: a round of a tile's dQ contribution is staged
  // store warp -> compute: the TMA reduce has read the staging buffer.  [0]: after a round of columns 32-63 (the next
  // writer is a columns-0-31 warp), [1]: after a round of columns 0-31.  One barrier per writer group, so that every
  // waiter sees CONSECUTIVE phases of its barrier (with a single barrier advancing twice per tile a warp could find
  // it two phases behind and take the stale parity for its own).
  uint64_t* dq_stage_free = sdp_full + 10;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sdp_full + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_it = (S + kTile - 1) / kTile;
  const int n_my = (n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_glob = n_my * n_it;
  auto item_tile = [&](int k) { return ((int)blockIdx.x + k * (int)gridDim.x) % n_it; };
  auto item_bh = [&](int k) { return ((int)blockIdx.x + k * (int)gridDim.x) / n_it; };  // = b * H + h

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dqkv);
    tma_prefetch_desc(&tm_dq);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_empty[i], 1);
      mbar_init(&aug_full[i], 1);
      mbar_init(&aug_empty[i], 1);
    }
    for (int i = 0; i < kB1Stages; ++i) {
      mbar_init(&st_full[i], 1);
      mbar_init(&st_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 8);
    mbar_init(pds_full, 8);
    mbar_init(acc_done, 1);
    mbar_init(epi_full, 8);
    mbar_init(dq_staged, 4);   // one round = the four warps of a column half
    mbar_init(&dq_stage_free[0], 1);
    mbar_init(&dq_stage_free[1], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320;
  const uint32_t tP = tmem_base + 384, tDQ = tmem_base + 448;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kBwdRegsCtl));
    if (warp == 0) {
      if (lane == 0) {
        int g = 0;
        for (int k = 0; k < n_my; ++k) {
          const int own0 = item_tile(k) * kTile, bh = item_bh(k), h = bh % H, b = bh / H;
          uint8_t* r0 = sR + (k & 1) * 2 * kTileBytes;
          mbar_wait(&r_empty[k & 1], ((uint32_t)(k >> 1) & 1u) ^ 1u);
          mbar_expect_tx(&r_full[k & 1], 2 * kTileBytes);
          tma_load_4d(r0, &tm_qkv, &r_full[k & 1], 0, H + h, own0, b);
          tma_load_4d(r0 + kTileBytes, &tm_qkv, &r_full[k & 1], 0, 2 * H + h, own0, b);
          for (int i = 0; i < n_it; ++i, ++g) {
            const int st = g % kB1Stages;
            mbar_wait(&st_empty[st], ((uint32_t)(g / kB1Stages) & 1u) ^ 1u);
            mbar_expect_tx(&st_full[st], 2 * kTileBytes);
            tma_load_4d(sX + st * kTileBytes, &tm_qkv, &st_full[st], 0, h, i * kTile, b);
            tma_load_4d(sY + st * kTileBytes, &tm_do, &st_full[st], 0, h, i * kTile, b);
          }
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc_acc = umma_idesc_bf16(64, 0, 1, 128);  // A K-major (TMEM or smem), B MN-major, N = 64
      constexpr uint32_t idesc_dq = umma_idesc_bf16(64, 1, 1, 128);   // A = dS^T read MN-major, B = K_j MN-major
      auto valid16 = [&](int i) { return min(kTile, (S - i * kTile + 15) & ~15); };
      auto issue_s_dp = [&](int k, int i, int st, int gg) {
        const uint64_t dAug = desc_k(smem_u32(sAug), 0);
        const uint32_t idesc_s = umma_idesc_bf16(valid16(i), 0, 0, 128);
        const uint32_t aR = smem_u32(sR + (k & 1) * 2 * kTileBytes);
        const uint64_t dR0 = desc_k(aR, 0), dR1 = desc_k(aR + kTileBytes, 0);
        const uint64_t dX = desc_k(smem_u32(sX + st * kTileBytes), 0), dY = desc_k(smem_u32(sY + st * kTileBytes), 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tS, dR0 + 2 * kk, dX + 2 * kk, idesc_s, kk > 0);
        umma_bf16_ss(tS, dAug, dAug + 4 + 2 * (gg & 1), idesc_s, 1);       // S^T - lse_q / scale
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tDP, dR1 + 2 * kk, dY + 2 * kk, idesc_s, kk > 0);
        umma_bf16_ss(tDP, dAug + 2, dAug + 4 + 2 * (gg & 1), idesc_s, 1);  // dP^T - delta_q
        umma_commit(&aug_empty[gg & 1]);
      };
      // ONE wait per issue: the statistics warp arrives on aug_full only after it has itself seen the tile's operands
      // land (r_full / st_full) and the previous scores leave TMEM (sdp_free).  tcgen05.mma issue is close to
      // synchronous (the queue holds a couple of MMAs), so every mbarrier round trip of THIS warp (~100 clk even when
      // the phase has long completed) is tensor-pipe idle time: five waits per tile cost ~650 of ~2900 clk.
      if (n_glob > 0) {
        mbar_wait(&aug_full[0], 0);
        tc_fence_after();
        if (elect_one()) {
          issue_s_dp(0, 0, 0, 0);
          umma_commit(sdp_full);
        }
        __syncwarp();
      }
      int k = 0, i = 0;
      for (int g = 0; g < n_glob; ++g) {
        int k1 = k, i1 = i + 1;
        if (i1 == n_it) {
          i1 = 0;
          ++k1;
        }
        if (lane == 0) BVC_TR(2, g, 0);
        if (g + 1 < n_glob) {
          const int st1 = (g + 1) % kB1Stages;
          if (lane == 0) BVC_TR(2, g, 1);
          mbar_wait(&aug_full[(g + 1) & 1], (uint32_t)((g + 1) >> 1) & 1u);
          if (lane == 0) BVC_TR(2, g, 2);
          tc_fence_after();
          if (elect_one()) {
            issue_s_dp(k1, i1, st1, g + 1);
            umma_commit(sdp_full);
          }
          __syncwarp();
        }
        if (lane == 0) BVC_TR(2, g, 3);
        mbar_wait(pds_full, (uint32_t)g & 1u);
        if (lane == 0) BVC_TR(2, g, 4);
        tc_fence_after();
        const int cst = g % kB1Stages;
        const int ksteps = valid16(i) >> 4;               // queries of the streamed tile: reduction of dV / dK
        const int ksteps_kv = valid16(item_tile(k)) >> 4; // keys of the owned tile: reduction of dQ
        if (elect_one()) {
          const uint64_t dX = desc_mn(smem_u32(sX + cst * kTileBytes), 0, 8192);   // Q_i : n = d, k = query
          const uint64_t dY = desc_mn(smem_u32(sY + cst * kTileBytes), 0, 8192);   // dO_i
          const uint64_t dKj = desc_mn(smem_u32(sR + (k & 1) * 2 * kTileBytes), 0, 8192);  // K_j : n = d, k = key
          const uint32_t ds_base = smem_u32(sDS);
          const uint32_t acc0 = i > 0;
          // dV[key, d] += P^T dO   (A = P^T from TMEM)
          for (int kk = 0; kk < ksteps; ++kk) umma_bf16_ts(tDV, tP + kk * 8, dY + 128 * kk, idesc_acc, acc0 | (kk > 0));
          // dK[key, d] += dS^T Q   (A = dS^T from smem, K-major: queries 16 kk .. inside atom kk / 4)
          for (int kk = 0; kk < ksteps; ++kk)
            umma_bf16_ss(tDK, desc_k(ds_base + (kk >> 2) * kTileBytes, kk & 3), dX + 128 * kk, idesc_acc, acc0 | (kk > 0));
          // dQ_i[query, d] = dS K_j (A = dS^T read MN-major: m = query, k = key; 16 keys = 2048 bytes per k-step, the
          // second 64 queries are the second atom)
          for (int kk = 0; kk < ksteps_kv; ++kk)
            umma_bf16_ss(tDQ, umma_smem_desc(ds_base + kk * 2048, 1024, kTileBytes), dKj + 128 * kk, idesc_dq, kk > 0);
          umma_commit(acc_done);
          umma_commit(&st_empty[cst]);
        }
        __syncwarp();
        if (lane == 0) BVC_TR(2, g, 5);
        k = k1;
        i = i1;
      }
    } else if (warp == 2) {
      // statistics warp: as in attn_bwd_kernel<1>
      const uint4 ones_s = make_uint4(pack_bf16x2(1.f, 1.f), pack_bf16x2(1.f, 0.f), 0u, 0u);
      const uint4 ones_dp = make_uint4(0u, pack_bf16x2(0.f, 1.f), pack_bf16x2(1.f, 1.f), 0u);
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      for (int r = lane; r < kTile; r += 32) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c != 4 && c != 6)
            *reinterpret_cast<uint4*>(sAug + r * 128 + ((c ^ (r & 7)) << 4)) = c == 0 ? ones_s : (c == 2 ? ones_dp : zero);
      }
      float cur[8], nxt[8];
      auto load8 = [&](int kk, int ii, float (&v)[8]) {
        const float* lb = lse + (long long)item_bh(kk) * S;
        const float* db = delta + (long long)item_bh(kk) * S;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int qg = min(ii * kTile + lane * 4 + j, S - 1);
          v[j] = __ldg(lb + qg);
          v[4 + j] = __ldg(db + qg);
        }
      };
      int k = 0, i = 0;
      if (n_glob > 0) load8(0, 0, cur);
      for (int g = 0; g < n_glob; ++g) {
        int k1 = k, i1 = i + 1;
        if (i1 == n_it) {
          i1 = 0;
          ++k1;
        }
        if (g + 1 < n_glob) load8(k1, i1, nxt);
        mbar_wait(&aug_empty[g & 1], ((uint32_t)(g >> 1) & 1u) ^ 1u);
        const int chunk = 4 + 2 * (g & 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = lane * 4 + j;
          float h3[2], m3[2], l3[2];
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const float v = w == 0 ? -cur[j] / scale : -cur[4 + j];
            h3[w] = __bfloat162float(__float2bfloat16_rn(v));
            const float r1 = v - h3[w];
            m3[w] = __bfloat162float(__float2bfloat16_rn(r1));
            l3[w] = r1 - m3[w];
          }
          *reinterpret_cast<uint4*>(sAug + r * 128 + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(pack_bf16x2(h3[0], m3[0]), pack_bf16x2(l3[0], h3[1]), pack_bf16x2(m3[1], l3[1]), 0u);
        }
        fence_async_smem();
        // relay for the MMA warp (see there): operands of tile g landed, scores of tile g - 1 read out of TMEM
        if (i == 0) mbar_wait(&r_full[k & 1], (uint32_t)(k >> 1) & 1u);
        mbar_wait(&st_full[g % kB1Stages], (uint32_t)(g / kB1Stages) & 1u);
        if (g > 0) mbar_wait(sdp_free, (uint32_t)(g - 1) & 1u);
        __syncwarp();
        if (lane == 0) mbar_arrive(&aug_full[g & 1]);
#pragma unroll
        for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
        k = k1;
        i = i1;
      }
    } else {
      // store warp: per streamed tile (from the second on) the previous tile's dQ contribution -> TMA reduce-add into
      // the fp32 accumulation buffer; per item dV / dK -> TMA stores.  Same event order as the compute warps produce.
      int n_dq = 0;   // rounds so far (two per tile: columns 0-31, then 32-63)
      auto reduce_dq = [&](int gp) {
        const int kp = gp / n_it, ip = gp % n_it, bh = item_bh(kp), h = bh % H, b = bh / H;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh, ++n_dq) {
          mbar_wait(dq_staged, (uint32_t)n_dq & 1u);
          if (lane == 0) {
            tma_reduce_add_4d(&tm_dq, sDQ, hh * 32, h, ip * kTile, b);
            tma_store_commit();
            tma_store_wait_read0();
            mbar_arrive(&dq_stage_free[hh ^ 1]);
          }
          __syncwarp();
        }
      };
      int g = 0;
      for (int k = 0; k < n_my; ++k) {
        const int own0 = item_tile(k) * kTile, bh = item_bh(k), h = bh % H, b = bh / H;
        uint8_t* stage = sR + (k & 1) * 2 * kTileBytes;
        for (int i = 0; i < n_it; ++i, ++g)
          if (g > 0) reduce_dq(g - 1);
        mbar_wait(epi_full, (uint32_t)k & 1u);
        if (lane == 0) {
          tma_store_4d(&tm_dqkv, stage, 0, 2 * H + h, own0, b);               // dV
          tma_store_4d(&tm_dqkv, stage + kTileBytes, 0, H + h, own0, b);      // dK
          tma_store_commit();
          tma_store_wait_read0();
          mbar_arrive(&r_empty[k & 1]);
        }
        __syncwarp();
      }
      if (n_glob > 0) reduce_dq(n_glob - 1);
      if (lane == 0) tma_store_wait0();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kBwdRegsCompute));
    const int e = warp - 4;
    const int q4 = warp & 3;
    const int half = e >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const float c_log2 = scale * kLog2e;
    const uint64_t cl2 = pack2(c_log2, c_log2), sc2 = pack2(scale, scale);
    const uint32_t ds_row = smem_u32(sDS) + (uint32_t)half * kTileBytes + (uint32_t)row * 128;   // this thread's dS^T row
    const uint32_t dq_row = smem_u32(sDQ) + (uint32_t)row * 128;                                 // ... dQ staging row
    const int sw = row & 7;
    int n_dq = 0;   // tiles drained so far
    uint32_t qv[32];
    // registers (a drained dQ accumulator) -> the staging tile: columns 0-31 after the reduce of the previous tile's
    // columns 32-63 has read it, columns 32-63 after this tile's columns 0-31
    auto stage_dq = [&]() {
      if (half == 1) mbar_wait(&dq_stage_free[1], (uint32_t)n_dq & 1u);
      else if (n_dq > 0) mbar_wait(&dq_stage_free[0], (uint32_t)(n_dq - 1) & 1u);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        sts_u4(dq_row + ((c ^ sw) << 4), qv[4 * c], qv[4 * c + 1], qv[4 * c + 2], qv[4 * c + 3]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_staged);
      ++n_dq;
    };
    int g = 0;
    for (int k = 0; k < n_my; ++k) {
      for (int i = 0; i < n_it; ++i, ++g) {
        const bool tr = lane == 0 && q4 == 0;
        if (tr) BVC_TR(half, g, 0);
        mbar_wait(sdp_full, (uint32_t)g & 1u);
        if (tr) BVC_TR(half, g, 1);
        tc_fence_after();
        // scores in two column chunks, the second chunk's TMEM loads in flight during the first chunk's math; the
        // S^T / dP^T columns go back to the MMA warp (next tile's score MMAs) once the second chunk is in registers
        uint32_t sv[2][32], dv[2][32];
        uint32_t pk[32], dk[32];
        auto math32 = [&](int c) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const uint64_t s2 = pack2(__uint_as_float(sv[c][j]), __uint_as_float(sv[c][j + 1]));
            const uint64_t p2 = pack2(__uint_as_float(dv[c][j]), __uint_as_float(dv[c][j + 1]));
            const uint64_t e2 = exp2_mufu2(fmul2(s2, cl2));   // P = exp2((S - lse/scale) * scale * log2e)
            float p0, p1, d0, d1;
            unpack2(e2, p0, p1);
            unpack2(fmul2(e2, fmul2(p2, sc2)), d0, d1);       // dS = P * (dP - delta) * scale
            pk[c * 16 + (j >> 1)] = pack_bf16x2(p0, p1);
            dk[c * 16 + (j >> 1)] = pack_bf16x2(d0, d1);
          }
        };
        tmem_ld_32x32b_x32(tS + lane_base + half * 64, sv[0]);
        tmem_ld_32x32b_x32(tDP + lane_base + half * 64, dv[0]);
        tmem_ld_wait_pin(sv[0]);
        tmem_ld_pin(dv[0]);
        tmem_ld_32x32b_x32(tS + lane_base + half * 64 + 32, sv[1]);
        tmem_ld_32x32b_x32(tDP + lane_base + half * 64 + 32, dv[1]);
        if (tr) BVC_TR(half, g, 2);
        math32(0);
        tmem_ld_wait_pin(sv[1]);
        tmem_ld_pin(dv[1]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_free);
        math32(1);
        if (tr) BVC_TR(half, g, 3);
        if (g > 0) {
          mbar_wait(acc_done, (uint32_t)(g - 1) & 1u);  // previous tile: P^T / dS^T consumed, dQ accumulator complete
          tc_fence_after();
          tmem_ld_32x32b_x32(tDQ + lane_base + half * 32, qv);   // in flight during the stores below
        }
        if (tr) BVC_TR(half, g, 4);
        // the tensor pipe idles from the end of the previous tile's accumulating MMAs until pds_full arrives (minus the
        // next score MMAs): only the dQ drain into registers and the P^T / dS^T stores sit on that path; staging the
        // drained tile for the store warp's reduce comes after the arrive
#pragma unroll
        for (int c = 0; c < 8; ++c) sts_u4(ds_row + ((c ^ sw) << 4), dk[4 * c], dk[4 * c + 1], dk[4 * c + 2], dk[4 * c + 3]);
        tmem_st_32x32b_x32(tP + lane_base + half * 32, pk);
        fence_async_smem();
        if (g > 0) tmem_ld_wait_pin(qv);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_full);
        if (g > 0) stage_dq();
        if (tr) BVC_TR(half, g, 5);
      }
      // item epilogue: dV / dK -> bf16 -> the finished item's resident-tile buffer -> TMA stores (store warp)
      mbar_wait(acc_done, (uint32_t)(g - 1) & 1u);
      if (lane == 0 && q4 == 0) BVC_TR(half, g - 1, 6);
      tc_fence_after();
      uint8_t* stage = sR + (k & 1) * 2 * kTileBytes;
      auto stage32 = [&](uint32_t tacc, uint8_t* dst_tile) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(tacc + lane_base + half * 32, ov);
        tmem_ld_wait();
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          uint4 o4;
          o4.x = pack_bf16x2(__uint_as_float(ov[gq * 8 + 0]), __uint_as_float(ov[gq * 8 + 1]));
          o4.y = pack_bf16x2(__uint_as_float(ov[gq * 8 + 2]), __uint_as_float(ov[gq * 8 + 3]));
          o4.z = pack_bf16x2(__uint_as_float(ov[gq * 8 + 4]), __uint_as_float(ov[gq * 8 + 5]));
          o4.w = pack_bf16x2(__uint_as_float(ov[gq * 8 + 6]), __uint_as_float(ov[gq * 8 + 7]));
          *reinterpret_cast<uint4*>(dst_tile + row * 128 + (((half * 4 + gq) ^ (row & 7)) << 4)) = o4;
        }
      };
      stage32(tDV, stage);
      stage32(tDK, stage + kTileBytes);
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(epi_full);
      if (lane == 0 && q4 == 0) BVC_TR(half, g - 1, 7);
    }
    // the last tile's dQ contribution (acc_done of the last tile was awaited by the last item epilogue)
    if (n_glob > 0) {
      tc_fence_after();
      tmem_ld_32x32b_x32(tDQ + lane_base + half * 32, qv);
      tmem_ld_wait_pin(qv);
      stage_dq();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dqkv[b, s, 0, h, :] = bf16(dq_accum[b, s, h, :])   (8 floats -> one 16-byte store per thread step)
__global__ void __launch_bounds__(256) attn_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dqkv,
                                                              long long n8, int H) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = ldv_f4(acc + i * 8), b = ldv_f4(acc + i * 8 + 4);
    const long long row = i >> 3;            // (b * S + s) * H + h
    const int c8 = (int)(i & 7);
    const long long bs = row / H;
    const int h = (int)(row - bs * H);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y);
    o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y);
    o.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(dqkv + (bs * 3 * H + h) * 64 + c8 * 8) = o;
  }
}

}  // namespace bvc

using namespace bvc;

namespace bvc {
int attn_small_fwd_launch(const void* qkv, int B, int S, int H, float scale, void* out, float* lse, cudaStream_t st);
int attn_small_bwd_launch(const void* qkv, const void* dout, const float* lse, const float* delta, int B, int S, int H,
                          float scale, void* dqkv, cudaStream_t st);
// BVC_ATTN_SMALL=0 forces the general kernels for short sequences too (A/B measurements)
static bool small_path_enabled() {
  static const bool on = []() {
    const char* e = getenv("BVC_ATTN_SMALL");
    return !(e && e[0] == '0');
  }();
  return on;
}
}  // namespace bvc

extern "C" int bvc_attn_fwd(const void* qkv, int32_t B, int32_t S, int32_t H, float scale, void* out, float* lse,
                            void* stream) {
  BVC_CHECK_ARG(qkv && out && lse && B > 0 && S > 0 && H > 0);
  BVC_CHECK_ARG((((uintptr_t)qkv) & 15) == 0 && (((uintptr_t)out) & 15) == 0);
  // short sequences (the encoder's 160 visible tokens): whole-sequence-resident kernel, attn_small.cu
  if (S <= 192 && small_path_enabled()) return attn_small_fwd_launch(qkv, B, S, H, scale, out, lse, (cudaStream_t)stream);
  static const bool attr_ok = !(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem) != cudaSuccess);  // once, thread-safe (C++11 static initialisation)
  if (!attr_ok) return BVC_ERR_LAUNCH;
  CUtensorMap tm, to;
  int rc = make_head_tmap(&tm, qkv, 3 * H, S, B);
  if (rc) return rc;
  rc = make_head_tmap(&to, out, H, S, B);
  if (rc) return rc;
  const int n_kv = (S + kTile - 1) / kTile;
  const long long n_work = (long long)((n_kv + 1) / 2) * H * B;  // (query-tile pair, head, clip)
  BVC_CHECK_ARG(n_work < (1ll << 30));
  const int grid = (int)(n_work < num_sms() ? n_work : num_sms());  // persistent: one CTA per SM
  attn_fwd_kernel<<<grid, kFwdThreads, kFwdSmem, (cudaStream_t)stream>>>(tm, to, lse, S, H, (int)n_work, scale * kLog2e);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int32_t B, int32_t S,
                            int32_t H, float scale, float* delta, void* dqkv, float* dq_accum, int32_t dq_accum_zeroed,
                            void* stream) {
  BVC_CHECK_ARG(qkv && out && dout && lse && delta && dqkv && B > 0 && S > 0 && H > 0);
  BVC_CHECK_ARG((((uintptr_t)qkv) & 15) == 0 && (((uintptr_t)dout) & 15) == 0 && (((uintptr_t)dqkv) & 15) == 0);
  static const bool attr_ok = !(cudaFuncSetAttribute(attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attn_bwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kB1Smem) != cudaSuccess);
  if (!attr_ok) return BVC_ERR_LAUNCH;  // once, thread-safe (C++11 static initialisation)
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)B * S * H;
  long long g = (rows + 63) / 64;
  if (g > (long long)num_sms() * 8) g = (long long)num_sms() * 8;
  attn_delta_kernel<<<(int)g, 256, 0, st>>>((const bf16*)out, (const bf16*)dout, delta, rows, S, H);
  BVC_CHECK_LAUNCH();
  // short sequences: single-pass, whole-sequence-resident kernel (attn_small.cu); BVC_ATTN_SMALL_BWD=0 switches it off
  static const bool small_bwd = []() {
    const char* e = getenv("BVC_ATTN_SMALL_BWD");
    return !(e && e[0] == '0');
  }();
  if (S <= 160 && small_path_enabled() && small_bwd) return attn_small_bwd_launch(qkv, dout, lse, delta, B, S, H, scale, dqkv, st);
  CUtensorMap tq, td;
  int rc = make_head_tmap(&tq, qkv, 3 * H, S, B);
  if (rc) return rc;
  rc = make_head_tmap(&td, dout, H, S, B);
  if (rc) return rc;
  const long long n_work = (long long)((S + kTile - 1) / kTile) * H * B;
  BVC_CHECK_ARG(n_work < (1ll << 30));
  const int grid = (int)(n_work < num_sms() ? n_work : num_sms());  // persistent: one CTA per SM
  CUtensorMap tdq;
  rc = make_head_tmap(&tdq, dqkv, 3 * H, S, B);
  if (rc) return rc;
  // general sequences: ONE pass (attn_bwd1_kernel) when the caller provides the fp32 dQ accumulation workspace
  // [B, S, H, 64]; BVC_ATTN_BWD1=0 (or a null workspace) selects the two-pass kernels (A/B measurements)
  static const bool one_pass = []() {
    const char* e = getenv("BVC_ATTN_BWD1");
    return !(e && e[0] == '0');
  }();
  if (dq_accum != nullptr && one_pass) {
    BVC_CHECK_ARG((((uintptr_t)dq_accum) & 15) == 0);
    CUtensorMap tacc;
    {
      const uint64_t dims[4] = {64, (uint64_t)H, (uint64_t)S, (uint64_t)B};
      const uint64_t strides[3] = {256, (uint64_t)H * 256, (uint64_t)S * H * 256};
      const uint32_t box[4] = {32, 1, 128, 1};
      rc = make_tmap(&tacc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dq_accum, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    if (!dq_accum_zeroed && cudaMemsetAsync(dq_accum, 0, (size_t)rows * 64 * sizeof(float), st) != cudaSuccess)
      return BVC_ERR_LAUNCH;
    attn_bwd1_kernel<<<grid, kBwdThreads, kB1Smem, st>>>(tq, td, tdq, tacc, lse, delta, S, H, (int)n_work, scale);
    BVC_CHECK_LAUNCH();
    const long long n8 = rows * 8;
    long long gc = (n8 + 255) / 256;
    if (gc > (long long)num_sms() * 16) gc = (long long)num_sms() * 16;
    attn_dq_convert_kernel<<<(int)gc, 256, 0, st>>>(dq_accum, (bf16*)dqkv, n8, H);
    BVC_CHECK_LAUNCH();
    return BVC_OK;
  }
  attn_bwd_kernel<1><<<grid, kBwdThreads, kBwdSmem, st>>>(tq, td, tdq, lse, delta, S, H, (int)n_work, scale);
  BVC_CHECK_LAUNCH();
  attn_bwd_kernel<0><<<grid, kBwdThreads, kBwdSmem, st>>>(tq, td, tdq, lse, delta, S, H, (int)n_work, scale);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

#ifdef BVC_TRACE
// which: 0 = only the Q pass runs afterwards ... the trace buffer holds whatever pass ran last on CTA 0
extern "C" int bvc_debug_trace_copy(long long* dst_host, int n) {
  if (n > kTraceRoles * kTraceTiles * kTracePoints) n = kTraceRoles * kTraceTiles * kTracePoints;
  return (int)cudaMemcpyFromSymbol(dst_host, g_trace, sizeof(long long) * n);
}
extern "C" int bvc_debug_attn_fwd(const void* qkv, int32_t B, int32_t S, int32_t H, float scale, void* out, float* lse,
                                  void* stream) {
  return bvc_attn_fwd(qkv, B, S, H, scale, out, lse, stream);
}
extern "C" int bvc_debug_attn_bwd_pass(const void* qkv, const void* dout, const float* lse, const float* delta,
                                       int32_t B, int32_t S, int32_t H, float scale, void* dqkv, int32_t mode_kv,
                                       void* stream) {
  cudaFuncSetAttribute(attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
  cudaFuncSetAttribute(attn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
  CUtensorMap tq, td, tdq;
  if (make_head_tmap(&tq, qkv, 3 * H, S, B) || make_head_tmap(&td, dout, H, S, B) ||
      make_head_tmap(&tdq, dqkv, 3 * H, S, B))
    return -1;
  const long long n_work = (long long)((S + kTile - 1) / kTile) * H * B;
  const int grid = (int)(n_work < num_sms() ? n_work : num_sms());
  if (mode_kv == 3) {  // the one-pass kernel; dQ contributions go to a scratch the caller passes in place of nothing:
    // the trace only needs the timing, so the accumulation buffer is the static one below (sized for the trace shapes)
    static float* acc = nullptr;
    static size_t acc_bytes = 0;
    const size_t need = (size_t)B * S * H * 64 * sizeof(float);
    if (need > acc_bytes) {
      if (acc) cudaFree(acc);
      if (cudaMalloc(&acc, need) != cudaSuccess) return -2;
      acc_bytes = need;
    }
    cudaFuncSetAttribute(attn_bwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kB1Smem);
    CUtensorMap tacc;
    const uint64_t dims[4] = {64, (uint64_t)H, (uint64_t)S, (uint64_t)B};
    const uint64_t strides[3] = {256, (uint64_t)H * 256, (uint64_t)S * H * 256};
    const uint32_t box[4] = {32, 1, 128, 1};
    if (make_tmap(&tacc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, acc, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return -3;
    attn_bwd1_kernel<<<grid, kBwdThreads, kB1Smem, (cudaStream_t)stream>>>(tq, td, tdq, tacc, lse, delta, S, H, (int)n_work, scale);
  } else if (mode_kv)
    attn_bwd_kernel<1><<<grid, kBwdThreads, kBwdSmem, (cudaStream_t)stream>>>(tq, td, tdq, lse, delta, S, H,
                                                                             (int)n_work, scale);
  else
    attn_bwd_kernel<0><<<grid, kBwdThreads, kBwdSmem, (cudaStream_t)stream>>>(tq, td, tdq, lse, delta, S, H,
                                                                             (int)n_work, scale);
  return (int)cudaGetLastError();
}
#endif
