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
