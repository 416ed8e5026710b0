// nce.cu -- the SimCLR loss of the contrastive path (pretraining/contrastive/pretrain_simclr.py:114-128 `info_nce_loss`,
// masks :86-91, :284-292) on the same tensor-core GEMM as the VideoMAE step.  The reference materialises
// F.cosine_similarity(feats[:, None], feats[None]) as an n x n x D fp32 tensor (2.1 GB at n = 1024, D = 512); here
//     S = (F^ . F^T) / T        is one tcgen05 GEMM (bvc_gemm_bf16) over row-normalised features,
//     loss = logsumexp(S[neg]) - mean(S[pos])   one masked pass over S (the reference's flattened, GLOBAL log-sum-exp),
//     dF^ = (G + G^T) . F^ / T  a second GEMM with G = neg * softmax weight - pos / P, then the normalisation backward.
// Precision: the GEMM takes bf16 operands; every fp32 operand x is split into hi = bf16(x), lo = bf16(x - hi) and the
// three significant products are concatenated along K ([hi | lo | hi] . [hi | hi | lo]^T), which carries ~16 mantissa
// bits -- the similarity logits come out at fp32-like accuracy although they ride the bf16 tensor pipe.
#include "../../include/bvc.h"
#include "bvc_host.h"
#include "bvc_ptx.cuh"

namespace bvc {

__device__ __forceinline__ void split_bf16(float x, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// one warp per row: inv_norm = 1 / max(|f|, eps) (ATen cosine_similarity), f^ = f * inv_norm, split, three layouts:
//   a_split [n, 3D] = [hi | lo | hi]   b_split [n, 3D] = [hi | hi | lo]   (forward: S = a_split . b_split^T)
//   bk_split [3n, D] = [hi ; hi ; lo]  (backward: dF^ = [G_hi | G_lo | G_hi] . bk_split, bk_split read MN-major)
template <bool BF16_IN>
__global__ void __launch_bounds__(256) nce_normalize_kernel(const void* __restrict__ feats, long long ld, int n, int D,
                                                            float eps, bf16* __restrict__ a_split,
                                                            bf16* __restrict__ b_split, bf16* __restrict__ bk_split,
                                                            float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  auto ldf = [&](int c) {
    return BF16_IN ? __bfloat162float(reinterpret_cast<const bf16*>(feats)[(long long)r * ld + c])
                   : reinterpret_cast<const float*>(feats)[(long long)r * ld + c];
  };
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float v = ldf(c);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
  if (lane == 0) inv_norm[r] = inv;
  for (int c = lane; c < D; c += 32) {
    bf16 hi, lo;
    split_bf16(ldf(c) * inv, hi, lo);
    bf16* a = a_split + (long long)r * 3 * D;
    bf16* b = b_split + (long long)r * 3 * D;
    a[c] = hi; a[D + c] = lo; a[2 * D + c] = hi;
    b[c] = hi; b[D + c] = hi; b[2 * D + c] = lo;
    bk_split[(long long)r * D + c] = hi;
    bk_split[((long long)n + r) * D + c] = hi;
    bk_split[((long long)2 * n + r) * D + c] = lo;
  }
}

// online (max, sum exp) pair combine
__device__ __forceinline__ void lse_combine(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  if (mn == -INFINITY) return;  // both empty
  s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
  m = mn;
}

// masked pass over S: per-block partials {max over neg, sum exp(S - max) over neg, sum of S over pos, count of pos}
__global__ void __launch_bounds__(256) nce_reduce_kernel(const float* __restrict__ S, long long lds,
                                                         const uint8_t* __restrict__ pos,
                                                         const uint8_t* __restrict__ neg, int n,
                                                         float* __restrict__ partials) {
  float m = -INFINITY, s = 0.f, ps = 0.f, pc = 0.f;
  const long long total = (long long)n * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), j = (int)(e - (long long)i * n);
    const float v = S[(long long)i * lds + j];
    if (neg[e]) {
      if (v > m) {
        s = s * __expf(m - v) + 1.0f;
        m = v;
      } else {
        s += __expf(v - m);
      }
    }
    if (pos[e]) {
      ps += v;
      pc += 1.0f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_combine(m, s, m2, s2);
    ps += __shfl_xor_sync(0xffffffffu, ps, o);
    pc += __shfl_xor_sync(0xffffffffu, pc, o);
  }
  __shared__ float sm[8][4];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    sm[w][0] = m; sm[w][1] = s; sm[w][2] = ps; sm[w][3] = pc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      lse_combine(m, s, sm[k][0], sm[k][1]);
      ps += sm[k][2];
      pc += sm[k][3];
    }
    float* o = partials + (long long)blockIdx.x * 4;
    o[0] = m; o[1] = s; o[2] = ps; o[3] = pc;
  }
}

// out[0] = loss = lse - mean_pos, out[1] = lse, out[2] = mean_pos, out[3] = pos count
__global__ void __launch_bounds__(32) nce_finalize_kernel(const float* __restrict__ partials, int nparts,
                                                          float* __restrict__ out) {
  float m = -INFINITY, s = 0.f, ps = 0.f, pc = 0.f;
  for (int k = threadIdx.x; k < nparts; k += 32) {
    lse_combine(m, s, partials[4 * k], partials[4 * k + 1]);
    ps += partials[4 * k + 2];
    pc += partials[4 * k + 3];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_combine(m, s, m2, s2);
    ps += __shfl_xor_sync(0xffffffffu, ps, o);
    pc += __shfl_xor_sync(0xffffffffu, pc, o);
  }
  if (threadIdx.x == 0) {
    const float lse = m + __logf(s);
    const float mp = ps / pc;
    out[0] = lse - mp; out[1] = lse; out[2] = mp; out[3] = pc;
  }
}

// Gs[i, j] = grad_out * ((neg[i,j] + neg[j,i]) * exp(S[i,j] - lse) - (pos[i,j] + pos[j,i]) / P), split as [hi | lo | hi]
// (S is symmetric: the same three products in both orders), so that dF^ = Gs . F^ / T is one GEMM.
__global__ void __launch_bounds__(256) nce_grad_kernel(const float* __restrict__ S, long long lds,
                                                       const uint8_t* __restrict__ pos, const uint8_t* __restrict__ neg,
                                                       int n, const float* __restrict__ out4,
                                                       const float* __restrict__ grad_out, bf16* __restrict__ g_split) {
  const float lse = out4[1], invP = 1.0f / out4[3], go = grad_out ? *grad_out : 1.0f;
  const long long total = (long long)n * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), j = (int)(e - (long long)i * n);
    const long long et = (long long)j * n + i;
    const float wn = (float)(neg[e] + neg[et]), wp = (float)(pos[e] + pos[et]);
    float g = 0.f;
    if (wn != 0.f) g = wn * __expf(S[(long long)i * lds + j] - lse);
    g = go * (g - wp * invP);
    bf16 hi, lo;
    split_bf16(g, hi, lo);
    bf16* row = g_split + (long long)i * 3 * n;
    row[j] = hi; row[n + j] = lo; row[2 * n + j] = hi;
  }
}

// normalisation backward, one warp per row:  df = (dF^ - f^ (f^ . dF^)) * inv_norm   (f^ = f * inv_norm; rows whose
// norm fell below eps are scaled by the constant 1/eps, i.e. df = dF^ * inv_norm)
template <bool BF16_IN>
__global__ void __launch_bounds__(256) nce_normalize_bwd_kernel(const float* __restrict__ dfhat,
                                                                const void* __restrict__ feats, long long ld,
                                                                const float* __restrict__ inv_norm, int n, int D,
                                                                float eps, float* __restrict__ dfeats) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  auto ldf = [&](int c) {
    return BF16_IN ? __bfloat162float(reinterpret_cast<const bf16*>(feats)[(long long)r * ld + c])
                   : reinterpret_cast<const float*>(feats)[(long long)r * ld + c];
  };
  const float inv = inv_norm[r];
  const bool clamped = inv >= 1.0f / eps;
  float dot = 0.f;
  for (int c = lane; c < D; c += 32) dot = fmaf(ldf(c) * inv, dfhat[(long long)r * D + c], dot);
  dot = clamped ? 0.f : warp_sum(dot);
  for (int c = lane; c < D; c += 32)
    dfeats[(long long)r * D + c] = (dfhat[(long long)r * D + c] - ldf(c) * inv * dot) * inv;
}

}  // namespace bvc

using namespace bvc;

extern "C" int bvc_nce_normalize_split(const void* feats, int32_t feats_is_bf16, int64_t ld, int32_t n, int32_t D,
                                       float eps, void* a_split, void* b_split, void* bk_split, float* inv_norm,
                                       void* stream) {
  BVC_CHECK_ARG(feats && a_split && b_split && bk_split && inv_norm && n > 0 && D > 0 && ld >= D);
  const int grid = (n + 7) / 8;
  if (feats_is_bf16)
    nce_normalize_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(feats, ld, n, D, eps, (bf16*)a_split,
                                                                      (bf16*)b_split, (bf16*)bk_split, inv_norm);
  else
    nce_normalize_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(feats, ld, n, D, eps, (bf16*)a_split,
                                                                       (bf16*)b_split, (bf16*)bk_split, inv_norm);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int64_t bvc_nce_partial_slots(int32_t n) {
  long long blocks = ((long long)n * n + 256 * 8 - 1) / (256 * 8);
  const long long cap = (long long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return blocks * 4;
}

extern "C" int bvc_nce_loss(const float* S, int64_t lds, const uint8_t* pos_mask, const uint8_t* neg_mask, int32_t n,
                            float* partials, float* out4, void* stream) {
  BVC_CHECK_ARG(S && pos_mask && neg_mask && partials && out4 && n > 0 && lds >= n);
  const int blocks = (int)(bvc_nce_partial_slots(n) / 4);
  nce_reduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(S, lds, pos_mask, neg_mask, n, partials);
  BVC_CHECK_LAUNCH();
  nce_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partials, blocks, out4);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_nce_grad(const float* S, int64_t lds, const uint8_t* pos_mask, const uint8_t* neg_mask, int32_t n,
                            const float* out4, const float* grad_out, void* g_split, void* stream) {
  BVC_CHECK_ARG(S && pos_mask && neg_mask && out4 && g_split && n > 0 && lds >= n);
  long long blocks = ((long long)n * n + 256 * 4 - 1) / (256 * 4);
  if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
  nce_grad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(S, lds, pos_mask, neg_mask, n, out4, grad_out,
                                                                (bf16*)g_split);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_nce_normalize_bwd(const float* dfhat, const void* feats, int32_t feats_is_bf16, int64_t ld,
                                     const float* inv_norm, int32_t n, int32_t D, float eps, float* dfeats,
                                     void* stream) {
  BVC_CHECK_ARG(dfhat && feats && inv_norm && dfeats && n > 0 && D > 0 && ld >= D);
  const int grid = (n + 7) / 8;
  if (feats_is_bf16)
    nce_normalize_bwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(dfhat, feats, ld, inv_norm, n, D, eps,
                                                                          dfeats);
  else
    nce_normalize_bwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(dfhat, feats, ld, inv_norm, n, D, eps,
                                                                           dfeats);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}
