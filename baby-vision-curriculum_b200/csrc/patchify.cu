// patchify.cu -- tube-mask indexing and the one-pass tubelet patchify + normalised-pixel target kernel.
//
//   bvc_mask_count / bvc_mask_to_index : one warp per clip, ballot + popc prefix; bit-exact replacement of the
//       boolean-index ops (HF:121-122, 587-588, 669-670).
//   bvc_patchify_target : one CTA per (clip, temporal slot, row of patches).  A single 5-D TMA box
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

// smem layout of the staged box: [t][c][ph][w] fp32, w = 0..W-1
template <int TS>
__global__ void __launch_bounds__(256) patchify_target_kernel(const __grid_constant__ CUtensorMap tmap,
                                                              const int* __restrict__ slot, int Tg, int Hg, int Wg,
                                                              int W, int nv, int N, bf16* __restrict__ patches_vis,
                                                              float* __restrict__ target, int norm_pix) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* tile = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t bar;
  const int hg = blockIdx.x % Hg;
  const int tg = (blockIdx.x / Hg) % Tg;
  const int b = blockIdx.x / (Hg * Tg);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    mbar_expect_tx(&bar, (uint32_t)(TS * 3 * 16 * W * 4));
    tma_load_5d(tile, &tmap, &bar, 0, hg * 16, 0, tg * TS, b);
  }
  __syncthreads();
  mbar_wait(&bar, 0);

  constexpr int K = 3 * TS * 256;
  constexpr int NI = TS * 6;  // 8-row groups per tubelet: TS*3 planes x 2
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nm = N - nv;
  const int sub_row = lane >> 2, chunk = lane & 3;
  for (int wg = warp; wg < Wg; wg += nwarps) {
    const int n = (tg * Hg + hg) * Wg + wg;
    const int s = __ldg(slot + (long long)b * N + n);
    if (s < 0 || s >= N) continue;  // malformed mask row: flagged by bvc_mask_to_index, nothing is written
    float4 v[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int row = i * 8 + sub_row;  // = (t*3 + c)*16 + ph
      v[i] = *reinterpret_cast<const float4*>(tile + (long long)row * W + wg * 16 + chunk * 4);
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
      // output order f = ((t*16 + ph)*16 + pw)*3 + c ; lane writes float4 number lane + 32*m
      float4* dst = reinterpret_cast<float4*>(target + ((long long)b * nm + (s - nv)) * K);
      const float* tbase = tile + wg * 16;
#pragma unroll 4
      for (int m = 0; m < K / 128; ++m) {
        const int f4 = lane + 32 * m;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int f = f4 * 4 + e;
          const int c = f % 3, pix = f / 3;  // pix = (t*16 + ph)*16 + pw
          const int pw = pix & 15, ph = (pix >> 4) & 15, t = pix >> 8;
          const float xv = tbase[(long long)((t * 3 + c) * 16 + ph) * W + pw];
          const float mu = c == 0 ? mean[0] : (c == 1 ? mean[1] : mean[2]);
          const float iv = c == 0 ? inv[0] : (c == 1 ? inv[1] : inv[2]);
          o[e] = (xv - mu) * iv;
        }
        __stcs(dst + f4, make_float4(o[0], o[1], o[2], o[3]));
      }
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

extern "C" int bvc_patchify_target(const float* pixels, const int32_t* slot, int32_t B, int32_t T, int32_t C, int32_t H,
                                   int32_t W, int32_t ts, int32_t ps, int32_t nv, void* patches_vis, float* target,
                                   int32_t norm_pix, void* stream) {
  BVC_CHECK_ARG(pixels && slot && patches_vis && target);
  BVC_CHECK_ARG(C == 3 && ps == 16 && (ts == 1 || ts == 2));
  BVC_CHECK_ARG(B > 0 && T % ts == 0 && H % 16 == 0 && W % 16 == 0 && W <= 256);
  BVC_CHECK_ARG((((uintptr_t)pixels) & 15) == 0);
  const int Tg = T / ts, Hg = H / 16, Wg = W / 16;
  const int N = Tg * Hg * Wg;
  BVC_CHECK_ARG(nv >= 0 && nv <= N);
  CUtensorMap tm;
  const uint64_t dims[5] = {(uint64_t)W, (uint64_t)H, 3, (uint64_t)T, (uint64_t)B};
  const uint64_t strides[4] = {(uint64_t)W * 4, (uint64_t)W * H * 4, (uint64_t)W * H * 3 * 4,
                               (uint64_t)W * H * 3 * T * 4};
  const uint32_t box[5] = {(uint32_t)W, 16, 3, (uint32_t)ts, 1};
  int rc = make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, pixels, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;
  const size_t smem = (size_t)ts * 3 * 16 * W * 4 + 128;
  int nwarps = (Wg + 1) / 2;
  if (nwarps > 8) nwarps = 8;
  if (nwarps < 1) nwarps = 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = B * Tg * Hg;
  static bool attr1 = false, attr2 = false;
  if (ts == 1) {
    if (!attr1) {
      if (cudaFuncSetAttribute(patchify_target_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) !=
          cudaSuccess)
        return BVC_ERR_LAUNCH;
      attr1 = true;
    }
    patchify_target_kernel<1><<<grid, nwarps * 32, smem, st>>>(tm, slot, Tg, Hg, Wg, W, nv, N, (bf16*)patches_vis,
                                                               target, norm_pix);
  } else {
    if (!attr2) {
      if (cudaFuncSetAttribute(patchify_target_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) !=
          cudaSuccess)
        return BVC_ERR_LAUNCH;
      attr2 = true;
    }
    patchify_target_kernel<2><<<grid, nwarps * 32, smem, st>>>(tm, slot, Tg, Hg, Wg, W, nv, N, (bf16*)patches_vis,
                                                               target, norm_pix);
  }
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}
