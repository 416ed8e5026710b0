// jepa.cu -- the pieces of the predictive (JEPA) path that sit beside its ViT blocks:
//   apply_masks (pretraining/predictive/mask.py:58-67: torch.gather of kept patch rows, one output block per mask),
//   repeat_interleave_batch (tensors.py:65-71), the target branch F.layer_norm + apply_masks + repeat_interleave_batch
//   (pretrain_jepa.py:384-392) as ONE pass, F.smooth_l1_loss forward / backward (:399-402) and the momentum (EMA)
//   update of the target encoder (:426-432) as one multi-tensor launch.
// All of it is HBM-bound row / element work: one warp per gathered row, 16-byte accesses, grids sized in multiples of the
// SM count.  Index and copy work is bit-exact; the EMA reproduces torch's three fp32 roundings (mul, mul, add; no FMA).
#include "../../include/bvc.h"
#include "bvc_host.h"
#include "bvc_ptx.cuh"

namespace bvc {

// out[((i * repeat + r) * B + b), k, :] = x[b, idx[i, b, k], :]      (rows of row_bytes, moved as 16-byte chunks)
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint4* __restrict__ x, int B, int N, int chunks,
                                                          const long long* __restrict__ idx, int n_masks, int K,
                                                          int repeat, uint4* __restrict__ out, int* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)n_masks * B * K;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = warp0; r < rows; r += nwarps) {
    const int k = (int)(r % K);
    const int b = (int)((r / K) % B);
    const int i = (int)(r / ((long long)K * B));
    long long src = __ldg(idx + r);
    const bool bad = src < 0 || src >= N;
    if (bad) {
      if (lane == 0 && status) atomicExch(status, 1);
      src = 0;
    }
    const uint4* xr = x + ((long long)b * N + src) * chunks;
    for (int c0 = 0; c0 < chunks; c0 += 128) {  // up to 4 chunks per lane in flight
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * 32 + lane;
        v[u] = c < chunks ? __ldg(xr + c) : make_uint4(0u, 0u, 0u, 0u);
        if (bad) v[u] = make_uint4(0u, 0u, 0u, 0u);
      }
      for (int rr = 0; rr < repeat; ++rr) {
        uint4* orow = out + ((((long long)i * repeat + rr) * B + b) * K + k) * chunks;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * 32 + lane;
          if (c < chunks) orow[c] = v[u];
        }
      }
    }
  }
}

// backward of one mask's gather: dx[b, idx[b, k], :] += dy[b, k, :]  (indices of one mask are unique per sample, so a
// plain read-add-write; the masks are applied by consecutive launches, last mask first = autograd's accumulation order)
template <bool F32>
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const void* __restrict__ dy, int B, int N, int D,
                                                               const long long* __restrict__ idx, int K,
                                                               void* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * K;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = warp0; r < rows; r += nwarps) {
    const int b = (int)(r / K);
    const long long src = __ldg(idx + r);
    if (src < 0 || src >= N) continue;
    if (F32) {
      const float4* g = reinterpret_cast<const float4*>(dy) + r * (D >> 2);
      float4* d = reinterpret_cast<float4*>(dx) + ((long long)b * N + src) * (D >> 2);
      for (int c = lane; c < (D >> 2); c += 32) {
        const float4 a = d[c], q = __ldg(g + c);
        d[c] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
      }
    } else {
      const uint4* g = reinterpret_cast<const uint4*>(dy) + r * (D >> 3);
      uint4* d = reinterpret_cast<uint4*>(dx) + ((long long)b * N + src) * (D >> 3);
      for (int c = lane; c < (D >> 3); c += 32) {
        const uint4 a = d[c], q = __ldg(g + c);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, qw[4] = {q.x, q.y, q.z, q.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          o[j] = pack_bf16x2(__uint_as_float(aw[j] << 16) + __uint_as_float(qw[j] << 16),
                             __uint_as_float(aw[j] & 0xffff0000u) + __uint_as_float(qw[j] & 0xffff0000u));
        d[c] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// x[i * B + b] -> out[(i * repeat + r) * B + b]   (slabs of slab_chunks 16-byte chunks)
__global__ void __launch_bounds__(256) repeat_interleave_kernel(const uint4* __restrict__ x, int B, int n_groups,
                                                                long long slab_chunks, int repeat,
                                                                uint4* __restrict__ out) {
  const long long total = (long long)n_groups * B * slab_chunks;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long c = t % slab_chunks, s = t / slab_chunks;
    const int b = (int)(s % B), i = (int)(s / B);
    const uint4 v = __ldg(x + t);
    for (int r = 0; r < repeat; ++r) out[((((long long)i * repeat + r) * B) + b) * slab_chunks + c] = v;
  }
}

// target branch in one pass: out[((i * repeat + r) * B + b), k, :] = layer_norm(h[b, idx[i, b, k], :])  (no affine,
// biased variance, fp32 statistics -- F.layer_norm(h, (D,)) under autocast runs in fp32).  One warp per gathered row,
// the row stays in registers between the statistics and the `repeat` normalised copies.
template <bool F32>
__global__ void __launch_bounds__(256) jepa_targets_kernel(const void* __restrict__ h, int B, int N, int D,
                                                           const long long* __restrict__ idx, int n_masks, int K,
                                                           int repeat, float eps, float* __restrict__ out,
                                                           int* __restrict__ status) {
  constexpr int kMaxVec = 8;  // float4 per lane: D <= 1024
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 2;
  const float inv_d = 1.0f / (float)D;
  const long long rows = (long long)n_masks * B * K;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = warp0; r < rows; r += nwarps) {
    const int k = (int)(r % K);
    const int b = (int)((r / K) % B);
    const int i = (int)(r / ((long long)K * B));
    long long src = __ldg(idx + r);
    if (src < 0 || src >= N) {
      if (lane == 0 && status) atomicExch(status, 1);
      src = 0;
    }
    const long long row_off = ((long long)b * N + src) * D;
    float4 v[kMaxVec];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < kMaxVec; ++u) {
      const int c = lane + u * 32;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nvec) {
        if (F32) {
          v[u] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(h) + row_off) + c);
        } else {
          const uint2 p = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(h) + row_off) + c);
          v[u] = make_float4(__uint_as_float(p.x << 16), __uint_as_float(p.x & 0xffff0000u), __uint_as_float(p.y << 16),
                             __uint_as_float(p.y & 0xffff0000u));
        }
      }
      s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    }
    const float mu = warp_sum(s) * inv_d;
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < kMaxVec; ++u) {
      if (lane + u * 32 < nvec) {
        const float a = v[u].x - mu, bb = v[u].y - mu, c2 = v[u].z - mu, d2 = v[u].w - mu;
        ss += (a * a + bb * bb) + (c2 * c2 + d2 * d2);
      }
    }
    const float rs = 1.0f / sqrtf(warp_sum(ss) * inv_d + eps);
#pragma unroll
    for (int u = 0; u < kMaxVec; ++u)
      v[u] = make_float4((v[u].x - mu) * rs, (v[u].y - mu) * rs, (v[u].z - mu) * rs, (v[u].w - mu) * rs);
    for (int rr = 0; rr < repeat; ++rr) {
      float4* orow = reinterpret_cast<float4*>(out + ((((long long)i * repeat + rr) * B + b) * K + k) * D);
#pragma unroll
      for (int u = 0; u < kMaxVec; ++u) {
        const int c = lane + u * 32;
        if (c < nvec) orow[c] = v[u];
      }
    }
  }
}

__device__ __forceinline__ float ld_elem(const void* p, long long i, bool f32) {
  return f32 ? __ldg(reinterpret_cast<const float*>(p) + i)
             : __bfloat162float(reinterpret_cast<const bf16*>(p)[i]);
}

// smooth-L1 (Huber, beta): per-block partial sums in a fixed order (deterministic); bvc_loss_finalize divides by numel
template <bool F32>
__global__ void __launch_bounds__(256) smooth_l1_fwd_kernel(const void* __restrict__ z, const float* __restrict__ h,
                                                            long long n, float beta, float* __restrict__ partials) {
  __shared__ float sh[8];
  const float half_over_beta = 0.5f / beta, half_beta = 0.5f * beta;
  float acc = 0.f;
  const long long nv = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nv; t += stride) {
    const float4 hv = __ldg(reinterpret_cast<const float4*>(h) + t);
    float4 zv;
    if (F32) {
      zv = __ldg(reinterpret_cast<const float4*>(z) + t);
    } else {
      const uint2 p = __ldg(reinterpret_cast<const uint2*>(z) + t);
      zv = make_float4(__uint_as_float(p.x << 16), __uint_as_float(p.x & 0xffff0000u), __uint_as_float(p.y << 16),
                       __uint_as_float(p.y & 0xffff0000u));
    }
    const float d[4] = {zv.x - hv.x, zv.y - hv.y, zv.z - hv.z, zv.w - hv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = fabsf(d[j]);
      acc += a < beta ? half_over_beta * d[j] * d[j] : a - half_beta;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long t = (nv << 2) + threadIdx.x;
    const float dd = ld_elem(z, t, F32) - h[t], a = fabsf(dd);
    acc += a < beta ? half_over_beta * dd * dd : a - half_beta;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partials[blockIdx.x] = t;
  }
}

// dz = g * (|d| < beta ? d / beta : sign(d)) / numel,  d = z - h;  written in z's dtype
template <bool F32>
__global__ void __launch_bounds__(256) smooth_l1_bwd_kernel(const void* __restrict__ z, const float* __restrict__ h,
                                                            long long n, float beta, const float* __restrict__ g_dev,
                                                            float inv_numel, void* __restrict__ dz) {
  const float g = __ldg(g_dev) * inv_numel, inv_beta = 1.0f / beta;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const float d = ld_elem(z, t, F32) - __ldg(h + t);
    const float a = fabsf(d);
    const float v = g * (a < beta ? d * inv_beta : (d > 0.f ? 1.0f : -1.0f));
    if (F32) reinterpret_cast<float*>(dz)[t] = v;
    else reinterpret_cast<bf16*>(dz)[t] = __float2bfloat16_rn(v);
  }
}

// target-encoder momentum update, every parameter in one launch:  k = fl(fl(m * k) + fl((1 - m) * q))  -- the
// reference's `param_k.mul_(m).add_((1. - m) * param_q)` rounds three times; no FMA contraction here either
struct EmaEntry {
  float* dst;        // target-encoder parameter (updated in place)
  const float* src;  // online-encoder parameter
  long long n;
};
__global__ void __launch_bounds__(256) ema_update_kernel(const EmaEntry* __restrict__ table, float m, float one_minus_m) {
  const EmaEntry en = table[blockIdx.y];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(en.dst) | reinterpret_cast<uintptr_t>(en.src)) & 15) == 0;
  const long long nv = vec ? en.n >> 2 : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const float4 k = reinterpret_cast<const float4*>(en.dst)[i];
    const float4 q = __ldg(reinterpret_cast<const float4*>(en.src) + i);
    float4 o;
    o.x = __fadd_rn(__fmul_rn(m, k.x), __fmul_rn(one_minus_m, q.x));
    o.y = __fadd_rn(__fmul_rn(m, k.y), __fmul_rn(one_minus_m, q.y));
    o.z = __fadd_rn(__fmul_rn(m, k.z), __fmul_rn(one_minus_m, q.z));
    o.w = __fadd_rn(__fmul_rn(m, k.w), __fmul_rn(one_minus_m, q.w));
    reinterpret_cast<float4*>(en.dst)[i] = o;
  }
  for (long long i = (nv << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < en.n; i += stride)
    en.dst[i] = __fadd_rn(__fmul_rn(m, en.dst[i]), __fmul_rn(one_minus_m, en.src[i]));
}

static int rows_grid(long long rows) {
  long long g = (rows + 7) / 8;
  const long long cap = (long long)num_sms() * 8;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}
static int elems_grid(long long n_threads) {
  long long g = (n_threads + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace bvc

using namespace bvc;

extern "C" int bvc_jepa_apply_masks(const void* x, int32_t elem_bytes, int32_t B, int32_t N, int32_t D,
                                    const int64_t* idx, int32_t n_masks, int32_t K, int32_t repeat, void* out,
                                    int32_t* status, void* stream) {
  BVC_CHECK_ARG(x && idx && out && B > 0 && N > 0 && D > 0 && n_masks > 0 && K > 0 && repeat > 0);
  BVC_CHECK_ARG(elem_bytes == 2 || elem_bytes == 4);
  BVC_CHECK_ARG(((long long)D * elem_bytes) % 16 == 0);
  BVC_CHECK_ARG((((uintptr_t)x) & 15) == 0 && (((uintptr_t)out) & 15) == 0);
  const int chunks = (int)((long long)D * elem_bytes / 16);
  const long long rows = (long long)n_masks * B * K;
  gather_rows_kernel<<<rows_grid(rows), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)x, B, N, chunks, (const long long*)idx, n_masks, K, repeat, (uint4*)out, status);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_jepa_apply_masks_bwd(const void* dy, int32_t elem_bytes, int32_t B, int32_t N, int32_t D,
                                        const int64_t* idx, int32_t n_masks, int32_t K, void* dx, void* stream) {
  BVC_CHECK_ARG(dy && idx && dx && B > 0 && N > 0 && D > 0 && n_masks > 0 && K > 0);
  BVC_CHECK_ARG(elem_bytes == 2 || elem_bytes == 4);
  BVC_CHECK_ARG(((long long)D * elem_bytes) % 16 == 0);
  BVC_CHECK_ARG((((uintptr_t)dy) & 15) == 0 && (((uintptr_t)dx) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(dx, 0, (size_t)B * N * D * elem_bytes, st) != cudaSuccess) return BVC_ERR_LAUNCH;
  const long long rows = (long long)B * K;
  // autograd runs the gathers' backward nodes in reverse creation order, so the LAST mask's contribution lands first
  for (int i = n_masks - 1; i >= 0; --i) {
    const char* dyi = (const char*)dy + (size_t)i * rows * D * elem_bytes;
    const long long* idxi = (const long long*)idx + (size_t)i * rows;
    if (elem_bytes == 4)
      scatter_add_rows_kernel<true><<<rows_grid(rows), 256, 0, st>>>(dyi, B, N, D, idxi, K, dx);
    else
      scatter_add_rows_kernel<false><<<rows_grid(rows), 256, 0, st>>>(dyi, B, N, D, idxi, K, dx);
    BVC_CHECK_LAUNCH();
  }
  return BVC_OK;
}

extern "C" int bvc_repeat_interleave_batch(const void* x, int64_t slab_bytes, int32_t B, int32_t n_groups,
                                           int32_t repeat, void* out, void* stream) {
  BVC_CHECK_ARG(x && out && slab_bytes > 0 && slab_bytes % 16 == 0 && B > 0 && n_groups > 0 && repeat > 0);
  BVC_CHECK_ARG((((uintptr_t)x) & 15) == 0 && (((uintptr_t)out) & 15) == 0);
  const long long chunks = slab_bytes / 16;
  repeat_interleave_kernel<<<elems_grid((long long)n_groups * B * chunks), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)x, B, n_groups, chunks, repeat, (uint4*)out);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_jepa_targets(const void* h, int32_t h_is_f32, int32_t B, int32_t N, int32_t D, const int64_t* idx,
                                int32_t n_masks, int32_t K, int32_t repeat, float eps, float* out, int32_t* status,
                                void* stream) {
  BVC_CHECK_ARG(h && idx && out && B > 0 && N > 0 && n_masks > 0 && K > 0 && repeat > 0);
  BVC_CHECK_ARG(D > 0 && D % 4 == 0 && D <= 1024);
  BVC_CHECK_ARG((((uintptr_t)h) & 15) == 0 && (((uintptr_t)out) & 15) == 0);
  const long long rows = (long long)n_masks * B * K;
  cudaStream_t st = (cudaStream_t)stream;
  if (h_is_f32)
    jepa_targets_kernel<true><<<rows_grid(rows), 256, 0, st>>>(h, B, N, D, (const long long*)idx, n_masks, K, repeat,
                                                                eps, out, status);
  else
    jepa_targets_kernel<false><<<rows_grid(rows), 256, 0, st>>>(h, B, N, D, (const long long*)idx, n_masks, K, repeat,
                                                                 eps, out, status);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int64_t bvc_smooth_l1_slots(int64_t n) { return (int64_t)elems_grid((n + 3) / 4); }

extern "C" int bvc_smooth_l1_fwd(const void* z, int32_t z_is_f32, const float* h, int64_t n, float beta,
                                 float* partials, void* stream) {
  BVC_CHECK_ARG(z && h && partials && n > 0 && beta > 0.f);
  BVC_CHECK_ARG((((uintptr_t)z) & 15) == 0 && (((uintptr_t)h) & 15) == 0);
  const int grid = elems_grid((n + 3) / 4);
  if (z_is_f32)
    smooth_l1_fwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(z, h, (long long)n, beta, partials);
  else
    smooth_l1_fwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(z, h, (long long)n, beta, partials);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_smooth_l1_bwd(const void* z, int32_t z_is_f32, const float* h, int64_t n, float beta,
                                 const float* grad_out, void* dz, void* stream) {
  BVC_CHECK_ARG(z && h && grad_out && dz && n > 0 && beta > 0.f);
  const int grid = elems_grid(n);
  const float inv_numel = (float)(1.0 / (double)n);
  if (z_is_f32)
    smooth_l1_bwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(z, h, (long long)n, beta, grad_out, inv_numel, dz);
  else
    smooth_l1_bwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(z, h, (long long)n, beta, grad_out, inv_numel, dz);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_ema_update(const void* table, int32_t n_entries, double momentum, void* stream) {
  BVC_CHECK_ARG(table && n_entries > 0 && n_entries <= 65535);
  // the reference's Python scalars: m and (1. - m) are doubles, rounded to fp32 where the fp32 kernels consume them
  const float m = (float)momentum, om = (float)(1.0 - momentum);
  dim3 grid(num_sms(), n_entries);
  ema_update_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const EmaEntry*)table, m, om);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}
