// patchify.cu -- tube-mask indexing and the one-pass tubelet patchify + normalised-pixel target kernel.
//
//   bvc_mask_count / bvc_mask_to_index : one warp per clip, ballot + popc prefix; bit-exact replacement of the
//       boolean-index ops (HF:121-122, 587-588, 669-670).
//   bvc_patchify_target : persistent CTAs over (clip, temporal slot, row of patches, column split).  A 5-D TMA box
//       (W x 16 rows x 3 channels x ts frames) stages the row in shared memory; each warp then owns whole
//       tubelets: visible ones are emitted as bf16 GEMM rows in Conv3d K-order (c,t,ph,pw), masked ones get their
//       per-channel mean / unbiased variance over the ts*256 pixels (two-pass, from registers) and are emitted as
//       fp32 target rows in (t,ph,pw,c) order with coalesced 16-byte stores.  Pixels are read from HBM once.
//       Replaces HF:175-176 (im2col + cast of the visible tokens) and HF:598-670.
#include "../../include/bvc.h"
#include "bvc_host.h"
#include "bvc_ptx.cuh"

namespace bvc {

// ------------------------------------------------------------------------------------------------ mask kernels
__global__ void mask_count_kernel(const uint8_t* __restrict__ mask, int B, int N, int* __restrict__ n_visible) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  int cnt = 0;
  for (int n = lane; n < N; n += 32) cnt += (mask[(long long)b * N + n] == 0);
  cnt = (int)warp_sum((float)cnt);  // exact: counts < 2^24
  if (lane == 0) n_visible[b] = cnt;
}

__global__ void mask_to_index_kernel(const uint8_t* __restrict__ mask, int B, int N, int nv, int* __restrict__ vis_idx,
                                     int* __restrict__ msk_idx, int* __restrict__ slot, int* __restrict__ status) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int nm = N - nv;
  int vbase = 0, mbase = 0;
  for (int n0 = 0; n0 < N; n0 += 32) {
    const int n = n0 + lane;
    const bool in = n < N;
    const bool masked = in && mask[(long long)b * N + n] != 0;
    const bool visible = in && !masked;
    const unsigned vb = __ballot_sync(0xffffffffu, visible);
    const unsigned mb = __ballot_sync(0xffffffffu, masked);
    const unsigned lt = (1u << lane) - 1u;
    if (visible) {
      const int r = vbase + __popc(vb & lt);
      if (r < nv) vis_idx[(long long)b * nv + r] = n;
      slot[(long long)b * N + n] = r < nv ? r : -1;  // -1: row violates the equal-count contract (status bit set)
    } else if (masked) {
      const int r = mbase + __popc(mb & lt);
      if (r < nm) msk_idx[(long long)b * nm + r] = n;
      slot[(long long)b * N + n] = r < nm ? nv + r : -1;
    }
    vbase += __popc(vb);
    mbase += __popc(mb);
  }
  if (lane == 0 && vbase != nv) atomicOr(status, 1);
}

// ------------------------------------------------------------------------------------------------ patchify + target
__device__ __constant__ float kInStd[3] = {0.229f, 0.224f, 0.225f};
__device__ __constant__ float kInMean[3] = {0.485f, 0.456f, 0.406f};

// Persistent, TMA-staged: one CTA per SM walks work items (clip, temporal slot, row of patches, column split).  Warp 0
// streams 5-D TMA boxes  [TS frames][3 channels][16 rows][BW pixels] fp32  through a ring of shared-memory stages
// (mbarrier full / empty), so ~130 KB of loads are in flight per SM at all times; 7 consumer warps take whole tubelets
// out of a landed box.  A lane holds 4 consecutive pixels of one image row for every (frame, channel) plane -- which
// is exactly 3 consecutive float4 of the (t, ph, pw, c)-ordered target row per (frame, row) -- so the normalised target
// is written straight from registers (no shared-memory re-read, no index arithmetic), and the visible tubelets as
// bf16 rows in Conv3d K-order.  The first version (one CTA per box, load -> compute -> store in sequence, targets
// re-read element-wise from shared memory) reached 3.5 TB/s; HBM: 18.78 MB per clip.
constexpr int kPatchConsumers = 7;
constexpr int kPatchThreads = 32 * (1 + kPatchConsumers);

// U8: the clip arrives as uint8 frames (4x fewer bytes over PCIe and from HBM) and the dataset's ToTensor + Normalize
// (homeview.py:218-231: x / 255, then (x - mean) / std per channel) is applied here, in registers, with IEEE
// divisions in torchvision's operation order -- bit-identical to normalising on the host.
struct PixelNorm {
  float mean[3], std[3];
};

template <int TS, bool U8>
__global__ void __launch_bounds__(kPatchThreads, 1) patchify_target_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                           const int* __restrict__ slot, int Tg, int Hg,
                                                                           int Wg, int split, int n_items, int n_stages,
                                                                           int nv, int N, bf16* __restrict__ patches_vis,
                                                                           float* __restrict__ target, int norm_pix,
                                                                           PixelNorm pn) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int TB = Wg / split;   // tubelets per box
  const int BW = TB * 16;      // pixels per box row
  const int box_bytes = TS * 3 * 16 * BW * (U8 ? 1 : 4);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)n_stages * box_bytes);
  uint64_t* empty = full + n_stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kPatchConsumers);
    }
    fence_mbar_init();
  }
  __syncthreads();

  constexpr int K = 3 * TS * 256;
  constexpr int NI = TS * 6;  // 8-row groups per tubelet: TS*3 planes x 2
  const int nm = N - nv;
  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int sp = w % split, hg = (w / split) % Hg, tg = (w / (split * Hg)) % Tg, b = w / (split * Hg * Tg);
        const int st = it % n_stages;
        mbar_wait(&empty[st], ((uint32_t)(it / n_stages) & 1u) ^ 1u);
        mbar_expect_tx(&full[st], (uint32_t)box_bytes);
        tma_load_5d(base + (size_t)st * box_bytes, &tmap, &full[st], sp * BW, hg * 16, 0, tg * TS, b);
      }
    }
  } else {
    const int cw = warp - 1;
    const int sub_row = lane >> 2, chunk = lane & 3;
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int sp = w % split, hg = (w / split) % Hg, tg = (w / (split * Hg)) % Tg, b = w / (split * Hg * Tg);
      const int st = it % n_stages;
      const uint8_t* tile_b = base + (size_t)st * box_bytes;
      const float* tile = reinterpret_cast<const float*>(tile_b);
      mbar_wait(&full[st], (uint32_t)(it / n_stages) & 1u);
      for (int wl = cw; wl < TB; wl += kPatchConsumers) {
        const int n = (tg * Hg + hg) * Wg + sp * TB + wl;
        const int s = __ldg(slot + (long long)b * N + n);
        if (s < 0 || s >= N) continue;  // malformed mask row: flagged by bvc_mask_to_index, nothing is written
        float4 v[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int row = i * 8 + sub_row;  // = (t*3 + c)*16 + ph
          if (U8) {
            const uchar4 u = *reinterpret_cast<const uchar4*>(tile_b + (long long)row * BW + wl * 16 + chunk * 4);
            const int c = (i >> 1) % 3;
            const float mu = pn.mean[c], sd = pn.std[c];
            v[i].x = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u.x, 255.f), mu), sd);
            v[i].y = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u.y, 255.f), mu), sd);
            v[i].z = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u.z, 255.f), mu), sd);
            v[i].w = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u.w, 255.f), mu), sd);
          } else {
            v[i] = *reinterpret_cast<const float4*>(tile + (long long)row * BW + wl * 16 + chunk * 4);
          }
        }
        if (s < nv) {
          // visible: bf16 row in Conv3d K-order k = ((c*TS + t)*16 + ph)*16 + pw
          bf16* dst = patches_vis + ((long long)b * nv + s) * K;
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const int plane = i >> 1;  // t*3 + c
            const int t = plane / 3, c = plane % 3;
            const int ph = (i & 1) * 8 + sub_row;
            uint2 pk;
            pk.x = pack_bf16x2(v[i].x, v[i].y);
            pk.y = pack_bf16x2(v[i].z, v[i].w);
            *reinterpret_cast<uint2*>(dst + ((c * TS + t) * 16 + ph) * 16 + chunk * 4) = pk;
          }
        } else {
          // masked: per-channel statistics over TS*256 pixels, two-pass from registers
          float mean[3], inv[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < NI; ++i)
              if ((i >> 1) % 3 == c) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            const float mu = warp_sum(sum) * (1.0f / (TS * 256));
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < NI; ++i)
              if ((i >> 1) % 3 == c) {
                const float a = v[i].x - mu, bb = v[i].y - mu, cc = v[i].z - mu, dd = v[i].w - mu;
                ss += (a * a + bb * bb) + (cc * cc + dd * dd);
              }
            const float var = warp_sum(ss) * (1.0f / (TS * 256 - 1));
            if (norm_pix) {
              // HF:598-643 on p = x*std + mean:  (p - mean_p) / (sqrt(var_p) + 1e-6) = (x - mu) * std / (std*sd + 1e-6)
              mean[c] = mu;
              inv[c] = kInStd[c] / (kInStd[c] * sqrtf(var) + 1e-6f);
            } else {
              // HF:644-667: the un-normalised frames themselves, p = x*std + mean
              mean[c] = -kInMean[c] / kInStd[c];
              inv[c] = kInStd[c];
            }
          }
          // output order f = ((t*16 + ph)*16 + pw)*3 + c: this lane's 4 pixels x 3 channels of (t, ph) are 12
          // consecutive floats = 3 float4 (16-byte aligned: (.. + chunk*4) * 3 * 4 bytes)
          float* trow = target + ((long long)b * nm + (s - nv)) * K;
#pragma unroll
          for (int t = 0; t < TS; ++t)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const float4 c0 = v[(t * 3 + 0) * 2 + hf], c1 = v[(t * 3 + 1) * 2 + hf], c2 = v[(t * 3 + 2) * 2 + hf];
              const float m0 = mean[0], m1 = mean[1], m2 = mean[2], i0 = inv[0], i1 = inv[1], i2 = inv[2];
              const int ph = hf * 8 + sub_row;
              float4* dst = reinterpret_cast<float4*>(trow + ((t * 16 + ph) * 16 + chunk * 4) * 3);
              __stcs(dst + 0, make_float4((c0.x - m0) * i0, (c1.x - m1) * i1, (c2.x - m2) * i2, (c0.y - m0) * i0));
              __stcs(dst + 1, make_float4((c1.y - m1) * i1, (c2.y - m2) * i2, (c0.z - m0) * i0, (c1.z - m1) * i1));
              __stcs(dst + 2, make_float4((c2.z - m2) * i2, (c0.w - m0) * i0, (c1.w - m1) * i1, (c2.w - m2) * i2));
            }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);  // this warp has read everything it needs from the stage
    }
  }
}

}  // namespace bvc

using namespace bvc;

extern "C" int bvc_mask_count(const uint8_t* mask, int32_t B, int32_t N, int32_t* n_visible, void* stream) {
  BVC_CHECK_ARG(mask && n_visible && B > 0 && N > 0);
  mask_count_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(mask, B, N, n_visible);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_mask_to_index(const uint8_t* mask, int32_t B, int32_t N, int32_t nv, int32_t* vis_idx,
                                 int32_t* msk_idx, int32_t* slot, int32_t* status, void* stream) {
  BVC_CHECK_ARG(mask && vis_idx && msk_idx && slot && status && B > 0 && N > 0 && nv >= 0 && nv <= N);
  mask_to_index_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(mask, B, N, nv, vis_idx, msk_idx, slot, status);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

static int patchify_launch(const void* pixels, bool u8, PixelNorm pn, const int32_t* slot, int32_t B, int32_t T,
                           int32_t C, int32_t H, int32_t W, int32_t ts, int32_t ps, int32_t nv, void* patches_vis,
                           float* target, int32_t norm_pix, void* stream) {
  BVC_CHECK_ARG(pixels && slot && patches_vis && target);
  BVC_CHECK_ARG(C == 3 && ps == 16 && (ts == 1 || ts == 2));
  BVC_CHECK_ARG(B > 0 && T % ts == 0 && H % 16 == 0 && W % 16 == 0 && W <= 256);
  BVC_CHECK_ARG((((uintptr_t)pixels) & 15) == 0);
  const int esz = u8 ? 1 : 4;
  const int Tg = T / ts, Hg = H / 16, Wg = W / 16;
  const int N = Tg * Hg * Wg;
  BVC_CHECK_ARG(nv >= 0 && nv <= N);
  // column split: the widest box (whole tubelets) that keeps a stage <= 48 KB, so several stages fit per SM
  int split = 1;
  while (split < Wg && ((size_t)ts * 3 * 16 * (W / split) * esz > 48 * 1024 || Wg % split != 0)) ++split;
  BVC_CHECK_ARG(Wg % split == 0);
  const int BW = W / split;
  const size_t box_bytes = (size_t)ts * 3 * 16 * BW * esz;
  BVC_CHECK_ARG(box_bytes % 128 == 0);  // every stage of the ring starts 128-byte aligned
  int n_stages = (int)((200 * 1024) / box_bytes);
  if (n_stages > 6) n_stages = 6;
  BVC_CHECK_ARG(n_stages >= 2);
  CUtensorMap tm;
  const uint64_t dims[5] = {(uint64_t)W, (uint64_t)H, 3, (uint64_t)T, (uint64_t)B};
  const uint64_t strides[4] = {(uint64_t)W * esz, (uint64_t)W * H * esz, (uint64_t)W * H * 3 * esz,
                               (uint64_t)W * H * 3 * T * esz};
  const uint32_t box[5] = {(uint32_t)BW, 16, 3, (uint32_t)ts, 1};
  int rc = make_tmap(&tm, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, pixels, dims, strides,
                     box, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;
  const size_t smem = (size_t)n_stages * box_bytes + 2 * n_stages * sizeof(uint64_t) + 128;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_items_ll = (long long)B * Tg * Hg * split;
  BVC_CHECK_ARG(n_items_ll < (1ll << 31));
  const int n_items = (int)n_items_ll;
  const int grid = n_items < num_sms() ? n_items : num_sms();  // persistent: one CTA per SM
  static const bool attr_ok = !(cudaFuncSetAttribute(patchify_target_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024) !=
            cudaSuccess ||
        cudaFuncSetAttribute(patchify_target_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024) !=
            cudaSuccess ||
        cudaFuncSetAttribute(patchify_target_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024) !=
            cudaSuccess ||
        cudaFuncSetAttribute(patchify_target_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024) !=
            cudaSuccess);  // once, thread-safe (C++11 static initialisation)
  if (!attr_ok) return BVC_ERR_LAUNCH;
#define BVC_PATCHIFY_LAUNCH(TS_, U8_)                                                                                  \
  patchify_target_kernel<TS_, U8_><<<grid, kPatchThreads, smem, st>>>(tm, slot, Tg, Hg, Wg, split, n_items, n_stages, \
                                                                      nv, N, (bf16*)patches_vis, target, norm_pix, pn)
  if (ts == 1) {
    if (u8) BVC_PATCHIFY_LAUNCH(1, true);
    else BVC_PATCHIFY_LAUNCH(1, false);
  } else {
    if (u8) BVC_PATCHIFY_LAUNCH(2, true);
    else BVC_PATCHIFY_LAUNCH(2, false);
  }
#undef BVC_PATCHIFY_LAUNCH
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_patchify_target(const float* pixels, const int32_t* slot, int32_t B, int32_t T, int32_t C, int32_t H,
                                   int32_t W, int32_t ts, int32_t ps, int32_t nv, void* patches_vis, float* target,
                                   int32_t norm_pix, void* stream) {
  PixelNorm pn{{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
  return patchify_launch(pixels, false, pn, slot, B, T, C, H, W, ts, ps, nv, patches_vis, target, norm_pix, stream);
}

extern "C" int bvc_patchify_target_u8(const uint8_t* pixels, const float* mean3, const float* std3, const int32_t* slot,
                                      int32_t B, int32_t T, int32_t C, int32_t H, int32_t W, int32_t ts, int32_t ps,
                                      int32_t nv, void* patches_vis, float* target, int32_t norm_pix, void* stream) {
  BVC_CHECK_ARG(mean3 && std3 && std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f);
  PixelNorm pn{{mean3[0], mean3[1], mean3[2]}, {std3[0], std3[1], std3[2]}};
  return patchify_launch(pixels, true, pn, slot, B, T, C, H, W, ts, ps, nv, patches_vis, target, norm_pix, stream);
}
