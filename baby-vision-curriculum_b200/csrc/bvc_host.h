// bvc_host.h -- host-side helpers shared by the translation units of libbvc.so:
// error plumbing for the C-ABI (never throw, never sync, never allocate) and TMA descriptor encoding
// through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define BVC_OK 0
#define BVC_ERR_ARG (-1)      /* bad argument (shape / alignment / null pointer) */
#define BVC_ERR_DRIVER (-2)   /* driver entry point or tensor-map encode failed */
#define BVC_ERR_LAUNCH (-3)   /* kernel launch failed (cudaGetLastError) */

#define BVC_CHECK_ARG(cond)                                                                   \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      fprintf(stderr, "bvc: bad argument: %s (%s:%d)\n", #cond, __FILE__, __LINE__);           \
      return BVC_ERR_ARG;                                                                     \
    }                                                                                         \
  } while (0)

#define BVC_CHECK_LAUNCH()                                                                    \
  do {                                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess) {                                                                 \
      fprintf(stderr, "bvc: launch failed: %s (%s:%d)\n", cudaGetErrorString(e__), __FILE__,  \
              __LINE__);                                                                      \
      return BVC_ERR_LAUNCH;                                                                  \
    }                                                                                         \
  } while (0)

namespace bvc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// rank-N tiled tensor map. dims/box innermost first; strides_bytes has rank-1 entries (dims 1..rank-1).
inline int make_tmap(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    fprintf(stderr, "bvc: cuTensorMapEncodeTiled entry point unavailable\n");
    return BVC_ERR_DRIVER;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "bvc: cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu] box [%u %u %u]\n", (int)r,
            rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
            (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return BVC_ERR_DRIVER;
  }
  return BVC_OK;
}

inline int num_sms() {
  static int n = []() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    return v;
  }();
  return n;
}

}  // namespace bvc
