// rows.cu -- HBM-bound row kernels of the VideoMAE step: LayerNorm forward/backward (one warp per row, fp32
// statistics, 16-byte vector accesses), column sums (bias / mask_token gradients), fp32->bf16 casts and the
// mask-token half of the decoder input.  See include/bvc.h for the reference lines each one replaces.
#include "../../include/bvc.h"
#include "bvc_host.h"
#include "bvc_ptx.cuh"

namespace bvc {

struct Seg {
  int seg, stride, off;
  __device__ __forceinline__ long long row(int r) const {
    return seg > 0 ? (long long)(r / seg) * stride + (r % seg) + off : (long long)r;
  }
};

constexpr int kLnMaxVec = 8;  // float4 per lane -> d <= 1024

// ------------------------------------------------------------------------------------------------ LN forward
template <int VPL>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, long long ldx, Seg xs,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, int M, int d,
                                                            bf16* __restrict__ y, float* __restrict__ mean,
                                                            float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int nvec = d >> 2;
  const float inv_d = 1.0f / (float)d;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < M; r += gridDim.x * warps_per_block) {
    const float4* xr = reinterpret_cast<const float4*>(x + xs.row(r) * ldx);
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + i * 32;
      v[i] = (c < nvec) ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mu = warp_sum(s) * inv_d;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float a = v[i].x - mu, b = v[i].y - mu, cc = v[i].z - mu, dd = v[i].w - mu;
        ss += (a * a + b * b) + (cc * cc + dd * dd);
      }
    }
    const float rs = rsqrtf(warp_sum(ss) * inv_d + eps);
    if (lane == 0) {
      mean[r] = mu;
      rstd[r] = rs;
    }
    uint2* yr = reinterpret_cast<uint2*>(y + (long long)r * d);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
        uint2 o;
        o.x = pack_bf16x2((v[i].x - mu) * rs * g.x + b.x, (v[i].y - mu) * rs * g.y + b.y);
        o.y = pack_bf16x2((v[i].z - mu) * rs * g.z + b.z, (v[i].w - mu) * rs * g.w + b.w);
        yr[c] = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ LN backward
template <int VPL>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ x,
                                                            long long ldx, Seg xs, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ dres, int M, int d,
                                                            float* __restrict__ dx_f32, bf16* __restrict__ dx_bf16,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            float* __restrict__ dxsum) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int nvec = d >> 2;
  const float inv_d = 1.0f / (float)d;
  float4 dg[VPL], db[VPL], ds[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ds[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < M; r += gridDim.x * warps_per_block) {
    const long long pr = xs.row(r);
    const float4* xr = reinterpret_cast<const float4*>(x + pr * ldx);
    const uint2* dyr = reinterpret_cast<const uint2*>(dy + (long long)r * d);
    const float4* rr = reinterpret_cast<const float4*>(dres + pr * ldx);
    // every operand of the row first (3 x VPL independent 16-byte requests per lane in flight), then the math
    float4 xh[VPL], g[VPL], rv[VPL];
    uint2 dvv[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = min(lane + i * 32, nvec - 1);
      xh[i] = ldv_f4(xr + c);
      dvv[i] = ldv_u2(dyr + c);
      if (dres) rv[i] = ldv_f4(rr + c);
    }
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 xv = xh[i];
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const uint2 dv = dvv[i];
        const float d0 = __uint_as_float(dv.x << 16), d1 = __uint_as_float(dv.x & 0xffff0000u);
        const float d2 = __uint_as_float(dv.y << 16), d3 = __uint_as_float(dv.y & 0xffff0000u);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[i] = make_float4(d0 * gm.x, d1 * gm.y, d2 * gm.z, d3 * gm.w);
        dg[i].x += d0 * xh[i].x; dg[i].y += d1 * xh[i].y; dg[i].z += d2 * xh[i].z; dg[i].w += d3 * xh[i].w;
        db[i].x += d0; db[i].y += d1; db[i].z += d2; db[i].w += d3;
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      } else {
        xh[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        g[i] = xh[i];
      }
    }
    const float m1 = warp_sum(s1) * inv_d;
    const float m2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        float4 o = make_float4(rs * (g[i].x - m1 - xh[i].x * m2), rs * (g[i].y - m1 - xh[i].y * m2),
                               rs * (g[i].z - m1 - xh[i].z * m2), rs * (g[i].w - m1 - xh[i].w * m2));
        if (dres) {
          o.x += rv[i].x; o.y += rv[i].y; o.z += rv[i].z; o.w += rv[i].w;
        }
        ds[i].x += o.x; ds[i].y += o.y; ds[i].z += o.z; ds[i].w += o.w;
        if (dx_f32) reinterpret_cast<float4*>(dx_f32 + pr * ldx)[c] = o;
        if (dx_bf16) {
          uint2 pk;
          pk.x = pack_bf16x2(o.x, o.y);
          pk.y = pack_bf16x2(o.z, o.w);
          reinterpret_cast<uint2*>(dx_bf16 + pr * ldx)[c] = pk;
        }
      }
    }
  }
  // block reduction through shared memory ([warps][d] partials, fixed summation order), then one global atomic
  // per column per block
  extern __shared__ float sh[];
  const int wid = threadIdx.x >> 5;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : dxsum);
    if (dst == nullptr) continue;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) reinterpret_cast<float4*>(sh + wid * d)[c] = pass == 0 ? dg[i] : (pass == 1 ? db[i] : ds[i]);
    }
    __syncthreads();
    for (int col = threadIdx.x; col < d; col += blockDim.x) {
      float t = 0.f;
      for (int k = 0; k < warps_per_block; ++k) t += sh[k * d + col];
      atomicAdd(dst + col, t);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ column sums
// block (32 x 8): lane -> 8 consecutive columns (256 per block), warp -> row phase; 256 rows per block.
template <bool F32>
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ in, long long ld, Seg s, int M, int N,
                                                     float scale, const float* __restrict__ scale_dev,
                                                     float* __restrict__ out, int rows_per_block) {
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c < N) {
    for (int rb = r0 + wy; rb < r1; rb += 32) {  // 4 independent rows per step (volatile: really 4 requests in flight)
      if (F32) {
        float4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float* p = (const float*)in + s.row(min(rb + 8 * u, r1 - 1)) * ld + c;
          a[u] = ldv_f4(p);
          b[u] = ldv_f4(p + 4);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (rb + 8 * u < r1) {
            acc[0] += a[u].x; acc[1] += a[u].y; acc[2] += a[u].z; acc[3] += a[u].w;
            acc[4] += b[u].x; acc[5] += b[u].y; acc[6] += b[u].z; acc[7] += b[u].w;
          }
      } else {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ldv_u4((const bf16*)in + s.row(min(rb + 8 * u, r1 - 1)) * ld + c);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (rb + 8 * u < r1) {
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[2 * j] += __uint_as_float(w[j] << 16);
              acc[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
            }
          }
      }
    }
  }
  __shared__ float sh[8][256 + 8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[wy][lane * 8 + j] = acc[j];
  __syncthreads();
  const int col = threadIdx.x;
  if (blockIdx.x * 256 + col < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][col];
    const float sc = scale * (scale_dev ? __ldg(scale_dev) : 1.0f);
    atomicAdd(out + blockIdx.x * 256 + col, t * sc);
  }
}

// ------------------------------------------------------------------------------------------------ casts
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long nv = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const long long i = (nv << 3) + threadIdx.x;
    dst[i] = __float2bfloat16_rn(src[i]);
  }
}

// one launch for every weight of the model: entry t copies (dst fp32) or casts (dst bf16) counts[t] fp32 elements
struct CastEntry {
  const float* src;
  void* dst;
  long long n;      // elements; src and dst 16-byte aligned
  int dst_is_f32;
  int pad;
};
__global__ void __launch_bounds__(256) cast_multi_kernel(const CastEntry* __restrict__ table) {
  const CastEntry en = table[blockIdx.y];
  const long long nv = en.n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(en.src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(en.src) + 2 * i + 1);
    if (en.dst_is_f32) {
      reinterpret_cast<float4*>(en.dst)[2 * i] = a;
      reinterpret_cast<float4*>(en.dst)[2 * i + 1] = b;
    } else {
      uint4 o;
      o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
      o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
      reinterpret_cast<uint4*>(en.dst)[i] = o;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (en.n & 7)) {
    const long long i = (nv << 3) + threadIdx.x;
    if (en.dst_is_f32) reinterpret_cast<float*>(en.dst)[i] = en.src[i];
    else reinterpret_cast<bf16*>(en.dst)[i] = __float2bfloat16_rn(en.src[i]);
  }
}

__global__ void __launch_bounds__(256) rows_to_bf16_kernel(const float* __restrict__ src, long long ld, Seg s, int M,
                                                           int d, bf16* __restrict__ dst) {
  const int nvec = d >> 2;
  const long long total = (long long)M * nvec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int r = (int)(i / nvec), c = (int)(i % nvec);
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + s.row(r) * ld) + c);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(dst + (long long)r * d)[c] = o;
  }
}

__global__ void __launch_bounds__(256) decoder_mask_rows_kernel(float* __restrict__ x,
                                                                const float* __restrict__ mask_token,
                                                                const float* __restrict__ pos,
                                                                const int* __restrict__ msk_idx, int B, int N, int nv,
                                                                int d) {
  const int nvec = d >> 2;
  const int nm = N - nv;
  const long long total = (long long)B * nm * nvec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % nvec);
    const long long rj = i / nvec;
    const int j = (int)(rj % nm), b = (int)(rj / nm);
    const int n = __ldg(msk_idx + (long long)b * nm + j);
    const float4 pv = __ldg(reinterpret_cast<const float4*>(pos + (long long)n * d) + c);
    const float4 mt = __ldg(reinterpret_cast<const float4*>(mask_token) + c);
    reinterpret_cast<float4*>(x + ((long long)b * N + nv + j) * d)[c] =
        make_float4(pv.x + mt.x, pv.y + mt.y, pv.z + mt.z, pv.w + mt.w);
  }
}

static inline int grid_for(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace bvc

using namespace bvc;

#define BVC_LN_DISPATCH(VPL_EXPR, CALL)            \
  switch (VPL_EXPR) {                              \
    case 1: { constexpr int V = 1; CALL; } break;  \
    case 2: { constexpr int V = 2; CALL; } break;  \
    case 3: { constexpr int V = 3; CALL; } break;  \
    case 4: { constexpr int V = 4; CALL; } break;  \
    case 5: case 6: { constexpr int V = 6; CALL; } break; \
    default: { constexpr int V = 8; CALL; } break; \
  }

extern "C" int bvc_layernorm_fwd(const float* x, int64_t ldx, int32_t x_seg, int32_t x_seg_stride, int32_t x_seg_off,
                                 const float* gamma, const float* beta, float eps, int32_t M, int32_t d, void* y,
                                 float* mean, float* rstd, void* stream) {
  BVC_CHECK_ARG(x && gamma && beta && y && mean && rstd);
  BVC_CHECK_ARG(M > 0 && d > 0 && d % 4 == 0 && d <= 128 * kLnMaxVec && ldx % 4 == 0 && ldx >= d);
  Seg s{x_seg, x_seg_stride, x_seg_off};
  const int vpl = (d / 4 + 31) / 32;
  const int grid = grid_for(M, 8);
  cudaStream_t st = (cudaStream_t)stream;
  BVC_LN_DISPATCH(vpl, (layernorm_fwd_kernel<V><<<grid, 256, 0, st>>>(x, ldx, s, gamma, beta, eps, M, d, (bf16*)y,
                                                                       mean, rstd)));
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_layernorm_bwd(const void* dy, const float* x, int64_t ldx, int32_t x_seg, int32_t x_seg_stride,
                                 int32_t x_seg_off, const float* mean, const float* rstd, const float* gamma,
                                 const float* dres, int32_t M, int32_t d, float* dx_f32, void* dx_bf16, float* dgamma,
                                 float* dbeta, float* dxsum, void* stream) {
  BVC_CHECK_ARG(dy && x && mean && rstd && gamma && dgamma && dbeta && (dx_f32 || dx_bf16));
  BVC_CHECK_ARG(M > 0 && d > 0 && d % 4 == 0 && d <= 128 * kLnMaxVec && ldx % 4 == 0 && ldx >= d);
  Seg s{x_seg, x_seg_stride, x_seg_off};
  const int vpl = (d / 4 + 31) / 32;
  int grid = num_sms() * (vpl <= 4 ? 2 : 1);  // resident blocks only (register-limited): a 2nd wave would just add tail
  if (grid > (M + 7) / 8) grid = (M + 7) / 8;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t shm = 8 * (size_t)d * sizeof(float);  // [warps][d], <= 32 KB
  BVC_LN_DISPATCH(vpl, (layernorm_bwd_kernel<V><<<grid, 256, shm, st>>>((const bf16*)dy, x, ldx, s, mean, rstd, gamma,
                                                                         dres, M, d, dx_f32, (bf16*)dx_bf16, dgamma,
                                                                         dbeta, dxsum)));
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_colsum(const void* in, int32_t in_is_f32, int64_t ld, int32_t seg, int32_t seg_stride,
                          int32_t seg_off, int32_t M, int32_t N, float scale_host, const float* scale_dev, float* out,
                          void* stream) {
  BVC_CHECK_ARG(in && out && M > 0 && N > 0 && N % 8 == 0 && ld % 8 == 0);
  Seg s{seg, seg_stride, seg_off};
  const int gx = (N + 255) / 256;
  int rpb = 256;  // shrink the row chunk until the grid covers the machine ~3x (small M would leave SMs idle)
  while (rpb > 32 && (long long)gx * ((M + rpb - 1) / rpb) < 3LL * num_sms()) rpb >>= 1;
  dim3 grid(gx, (M + rpb - 1) / rpb);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_is_f32)
    colsum_kernel<true><<<grid, 256, 0, st>>>(in, ld, s, M, N, scale_host, scale_dev, out, rpb);
  else
    colsum_kernel<false><<<grid, 256, 0, st>>>(in, ld, s, M, N, scale_host, scale_dev, out, rpb);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  BVC_CHECK_ARG(src && dst && n > 0 && (((uintptr_t)src) & 15) == 0 && (((uintptr_t)dst) & 15) == 0);
  cast_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, (long long)n);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_cast_multi(const void* table, int32_t n_entries, void* stream) {
  BVC_CHECK_ARG(table && n_entries > 0 && n_entries <= 65535);
  dim3 grid(32, n_entries);
  cast_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const CastEntry*)table);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_rows_to_bf16(const float* src, int64_t ld, int32_t seg, int32_t seg_stride, int32_t seg_off,
                                int32_t M, int32_t d, void* dst, void* stream) {
  BVC_CHECK_ARG(src && dst && M > 0 && d > 0 && d % 4 == 0 && ld % 4 == 0);
  Seg s{seg, seg_stride, seg_off};
  rows_to_bf16_kernel<<<grid_for((long long)M * (d / 4), 256), 256, 0, (cudaStream_t)stream>>>(src, ld, s, M, d,
                                                                                              (bf16*)dst);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_decoder_mask_rows(float* x, const float* mask_token, const float* pos, const int32_t* msk_idx,
                                     int32_t B, int32_t N, int32_t nv, int32_t d, void* stream) {
  BVC_CHECK_ARG(x && mask_token && pos && msk_idx && B > 0 && N > nv && nv >= 0 && d % 4 == 0);
  decoder_mask_rows_kernel<<<grid_for((long long)B * (N - nv) * (d / 4), 256), 256, 0, (cudaStream_t)stream>>>(
      x, mask_token, pos, msk_idx, B, N, nv, d);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}
