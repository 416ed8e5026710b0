// optim.cu -- the optimizer half of the training step (pretrain_videomae.py:187-197, 312-314) as ONE multi-tensor pass:
// GradScaler unscale + skip-on-overflow, weight decay, (Nesterov) momentum, parameter update, and the bf16 shadow copy
// that the next forward's GEMMs read (what autocast would re-cast per call).  torch's foreach path makes 5-6 passes
// over the 94 M parameters (~5.6 GB of HBM traffic per step); this is one read of p/g/m and one write of p/g/m (+bf16).
#include "../../include/bvc.h"
#include "bvc_host.h"
#include "bvc_ptx.cuh"

namespace bvc {

struct SgdEntry {  // 48 bytes, mirrored by bvc_b200/optim.py
  float* p;
  float* g;
  float* m;        // momentum buffer (may be null when momentum == 0)
  void* shadow;    // bf16 (or fp32) copy of p kept for the forward pass, or null
  long long n;
  int shadow_is_f32;
  int m_uninit;    // 1: the momentum buffer holds nothing yet (torch: buf = clone(grad) on a parameter's first step)
};

struct SgdHyper {
  float lr, momentum, dampening, weight_decay;
  int nesterov;
};

__device__ __forceinline__ float sgd_one(float p, float g, float& m, const SgdHyper& h, bool first) {
  if (h.weight_decay != 0.f) g = fmaf(h.weight_decay, p, g);
  if (h.momentum != 0.f) {
    m = first ? g : fmaf(h.momentum, m, (1.f - h.dampening) * g);
    g = h.nesterov ? fmaf(h.momentum, m, g) : m;
  }
  return fmaf(-h.lr, g, p);
}

__global__ void __launch_bounds__(256) sgd_multi_kernel(const SgdEntry* __restrict__ table, SgdHyper h,
                                                        const float* __restrict__ grad_scale,
                                                        const float* __restrict__ found_inf) {
  if (found_inf != nullptr && *found_inf != 0.f) return;  // GradScaler: overflow somewhere -> the whole step is skipped
  const float inv = grad_scale != nullptr ? 1.0f / *grad_scale : 1.0f;
  const SgdEntry en = table[blockIdx.y];
  const bool has_m = h.momentum != 0.f && en.m != nullptr;
  const bool first = en.m_uninit != 0;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(en.p) | reinterpret_cast<uintptr_t>(en.g) |
                        reinterpret_cast<uintptr_t>(en.m) | reinterpret_cast<uintptr_t>(en.shadow)) & 15) == 0;
  const long long nv = vec_ok ? en.n >> 2 : 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float4 p = reinterpret_cast<const float4*>(en.p)[i];
    float4 g = reinterpret_cast<const float4*>(en.g)[i];
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_m && !first) m = reinterpret_cast<const float4*>(en.m)[i];
    g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
    p.x = sgd_one(p.x, g.x, m.x, h, first);
    p.y = sgd_one(p.y, g.y, m.y, h, first);
    p.z = sgd_one(p.z, g.z, m.z, h, first);
    p.w = sgd_one(p.w, g.w, m.w, h, first);
    reinterpret_cast<float4*>(en.p)[i] = p;
    if (grad_scale != nullptr) reinterpret_cast<float4*>(en.g)[i] = g;  // .grad holds unscaled values after step()
    if (has_m) reinterpret_cast<float4*>(en.m)[i] = m;
    if (en.shadow != nullptr) {
      if (en.shadow_is_f32) {
        reinterpret_cast<float4*>(en.shadow)[i] = p;
      } else {
        uint2 o;
        o.x = pack_bf16x2(p.x, p.y);
        o.y = pack_bf16x2(p.z, p.w);
        reinterpret_cast<uint2*>(en.shadow)[i] = o;
      }
    }
  }
  // scalar tail (n % 4, or a misaligned tensor)
  for (long long i = (nv << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < en.n; i += stride) {
    float p = en.p[i], g = en.g[i] * inv, m = (has_m && !first) ? en.m[i] : 0.f;
    p = sgd_one(p, g, m, h, first);
    en.p[i] = p;
    if (grad_scale != nullptr) en.g[i] = g;
    if (has_m) en.m[i] = m;
    if (en.shadow != nullptr) {
      if (en.shadow_is_f32) reinterpret_cast<float*>(en.shadow)[i] = p;
      else reinterpret_cast<bf16*>(en.shadow)[i] = __float2bfloat16_rn(p);
    }
  }
}

// ------------------------------------------------------------------------------------------------ Adam / AdamW
// torch.optim.AdamW(betas=(0.9, 0.95)) is the reference's other optimizer choice (pretrain_videomae.py:190-193).  Same
// table as SGD with `m` = exp_avg and a second table of exp_avg_sq pointers; op order of torch's _single_tensor_adam:
//   p *= 1 - lr wd (AdamW)  |  g += wd p (Adam);  m = lerp(m, g, 1 - b1);  v = b2 v + (1 - b2) g g;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The step count t lives on the device (one float per parameter group, shared by its parameters): a step skipped by
// GradScaler must not advance it, and the host never learns whether a step was skipped.  IEEE division / square root
// (this file is compiled without --use_fast_math).
struct AdamHyper {
  double lr, beta1, beta2, eps, weight_decay;  // torch does its scalar arithmetic on Python floats (doubles)
  int decoupled;                               // 1: AdamW
};
struct AdamScalars {  // every scalar rounded to fp32 ONCE, where torch's kernels receive it
  float w1, beta2, w2, eps, wd, decay, step_size, bc2_sqrt;
  int decoupled;
};

__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, const AdamScalars& h) {
  if (h.wd != 0.f) {
    if (h.decoupled) p = __fmul_rn(p, h.decay);
    else g = fmaf(h.wd, p, g);
  }
  m = fmaf(h.w1, g - m, m);                                       // lerp(m, g, 1 - beta1)
  v = fmaf(h.w2, __fmul_rn(g, g), __fmul_rn(v, h.beta2));         // v * beta2 + (1 - beta2) * (g * g)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), h.bc2_sqrt), h.eps);
  return fmaf(-h.step_size, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const SgdEntry* __restrict__ table,
                                                         float* const* __restrict__ exp_avg_sq, AdamHyper h,
                                                         const float* __restrict__ step, const float* __restrict__ grad_scale,
                                                         const float* __restrict__ found_inf) {
  if (found_inf != nullptr && *found_inf != 0.f) return;
  const float inv = grad_scale != nullptr ? 1.0f / *grad_scale : 1.0f;
  const SgdEntry en = table[blockIdx.y];
  float* vbuf = exp_avg_sq[blockIdx.y];
  const bool first = en.m_uninit != 0;  // state buffers hold nothing yet: exp_avg = exp_avg_sq = 0
  // scalars in double like torch's Python-float arithmetic, then one rounding to fp32
  const double t = (double)*step + 1.0;
  const double bc1 = 1.0 - pow(h.beta1, t), bc2 = 1.0 - pow(h.beta2, t);
  AdamScalars sc;
  sc.w1 = (float)(1.0 - h.beta1);
  sc.beta2 = (float)h.beta2;
  sc.w2 = (float)(1.0 - h.beta2);
  sc.eps = (float)h.eps;
  sc.wd = (float)h.weight_decay;
  sc.decay = (float)(1.0 - h.lr * h.weight_decay);
  sc.step_size = (float)(h.lr / bc1);
  sc.bc2_sqrt = (float)sqrt(bc2);
  sc.decoupled = h.decoupled;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(en.p) | reinterpret_cast<uintptr_t>(en.g) |
                        reinterpret_cast<uintptr_t>(en.m) | reinterpret_cast<uintptr_t>(vbuf) |
                        reinterpret_cast<uintptr_t>(en.shadow)) & 15) == 0;
  const long long nv = vec_ok ? en.n >> 2 : 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float4 p = reinterpret_cast<const float4*>(en.p)[i];
    float4 g = reinterpret_cast<const float4*>(en.g)[i];
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f), v = m;
    if (!first) {
      m = reinterpret_cast<const float4*>(en.m)[i];
      v = reinterpret_cast<const float4*>(vbuf)[i];
    }
    g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
    p.x = adam_one(p.x, g.x, m.x, v.x, sc);
    p.y = adam_one(p.y, g.y, m.y, v.y, sc);
    p.z = adam_one(p.z, g.z, m.z, v.z, sc);
    p.w = adam_one(p.w, g.w, m.w, v.w, sc);
    reinterpret_cast<float4*>(en.p)[i] = p;
    if (grad_scale != nullptr) reinterpret_cast<float4*>(en.g)[i] = g;
    reinterpret_cast<float4*>(en.m)[i] = m;
    reinterpret_cast<float4*>(vbuf)[i] = v;
    if (en.shadow != nullptr) {
      if (en.shadow_is_f32) {
        reinterpret_cast<float4*>(en.shadow)[i] = p;
      } else {
        uint2 o;
        o.x = pack_bf16x2(p.x, p.y);
        o.y = pack_bf16x2(p.z, p.w);
        reinterpret_cast<uint2*>(en.shadow)[i] = o;
      }
    }
  }
  for (long long i = (nv << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < en.n; i += stride) {
    float p = en.p[i], g = en.g[i] * inv, m = first ? 0.f : en.m[i], v = first ? 0.f : vbuf[i];
    p = adam_one(p, g, m, v, sc);
    en.p[i] = p;
    if (grad_scale != nullptr) en.g[i] = g;
    en.m[i] = m;
    vbuf[i] = v;
    if (en.shadow != nullptr) {
      if (en.shadow_is_f32) reinterpret_cast<float*>(en.shadow)[i] = p;
      else reinterpret_cast<bf16*>(en.shadow)[i] = __float2bfloat16_rn(p);
    }
  }
}

// after the update kernel (stream order): t += 1 unless the step was skipped
__global__ void adam_advance_kernel(float* step, const float* __restrict__ found_inf) {
  if (found_inf == nullptr || *found_inf == 0.f) *step += 1.0f;
}

// ------------------------------------------------------------------------------------------------ inf / nan check
// GradScaler's _amp_foreach_non_finite_check_and_unscale_ with inv_scale = 1 re-WRITES every gradient (5 launches,
// read + write); all the optimizer pass needs beforehand is the yes / no answer: one read-only multi-tensor launch.
__global__ void __launch_bounds__(256) nonfinite_multi_kernel(const SgdEntry* __restrict__ table,
                                                              float* __restrict__ found_inf) {
  const SgdEntry en = table[blockIdx.y];
  const bool vec_ok = (reinterpret_cast<uintptr_t>(en.g) & 15) == 0;
  const long long nv = vec_ok ? en.n >> 2 : 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  bool bad = false;
  // |x| <= FLT_MAX is false for inf and nan; four independent 16-byte loads per iteration in flight
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < nv; i += 4 * stride) {
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = ldv_f4(reinterpret_cast<const float4*>(en.g) + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      bad |= !(fabsf(a[u].x) <= 3.402823466e38f) | !(fabsf(a[u].y) <= 3.402823466e38f) |
             !(fabsf(a[u].z) <= 3.402823466e38f) | !(fabsf(a[u].w) <= 3.402823466e38f);
  }
  for (; i < nv; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(en.g)[i];
    bad |= !(fabsf(a.x) <= 3.402823466e38f) | !(fabsf(a.y) <= 3.402823466e38f) | !(fabsf(a.z) <= 3.402823466e38f) |
           !(fabsf(a.w) <= 3.402823466e38f);
  }
  for (long long j = (nv << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < en.n; j += stride)
    bad |= !(fabsf(en.g[j]) <= 3.402823466e38f);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *found_inf = 1.0f;  // benign race: every writer stores 1
}

}  // namespace bvc

using namespace bvc;

extern "C" int bvc_sgd_step(const void* table, int32_t n_entries, float lr, float momentum, float dampening,
                            float weight_decay, int32_t nesterov, const float* grad_scale,
                            const float* found_inf, void* stream) {
  BVC_CHECK_ARG(table && n_entries > 0 && n_entries <= 65535);
  SgdHyper h{lr, momentum, dampening, weight_decay, nesterov};
  dim3 grid(32, n_entries);
  sgd_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const SgdEntry*)table, h, grad_scale, found_inf);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_adam_step(const void* table, const void* exp_avg_sq_table, int32_t n_entries, double lr, double beta1,
                             double beta2, double eps, double weight_decay, int32_t decoupled, float* step,
                             const float* grad_scale, const float* found_inf, void* stream) {
  BVC_CHECK_ARG(table && exp_avg_sq_table && step && n_entries > 0 && n_entries <= 65535);
  BVC_CHECK_ARG(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0);
  AdamHyper h{lr, beta1, beta2, eps, weight_decay, decoupled};
  dim3 grid(32, n_entries);
  adam_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const SgdEntry*)table, (float* const*)exp_avg_sq_table, h,
                                                            step, grad_scale, found_inf);
  BVC_CHECK_LAUNCH();
  adam_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, found_inf);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}

extern "C" int bvc_grad_nonfinite(const void* table, int32_t n_entries, float* found_inf, void* stream) {
  BVC_CHECK_ARG(table && found_inf && n_entries > 0 && n_entries <= 65535);
  if (cudaMemsetAsync(found_inf, 0, sizeof(float), (cudaStream_t)stream) != cudaSuccess) return BVC_ERR_LAUNCH;
  dim3 grid(32, n_entries);
  nonfinite_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const SgdEntry*)table, found_inf);
  BVC_CHECK_LAUNCH();
  return BVC_OK;
}
