// bvc_host.h -- host-side helpers shared by the translation units of libbvc.so:
// error plumbing for the C-ABI (never throw, never sync, never allocate) and TMA descriptor encoding
// through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

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
// Descriptors are CACHED per (base pointer, shape, strides, box, dtype, swizzle): a training step re-launches the same
// ~700 kernels on buffers the caching allocator hands back at the same addresses, so after the first step every
// descriptor of a launch is a table hit (a 128-byte copy) instead of a driver call.  A descriptor
// depends on the address and geometry only, never on the memory's contents, so a hit can not be stale.  Thread-safe:
// the forward and the autograd thread both launch.
struct TmapKey {
  const void* base;
  uint64_t dims[5];
  uint64_t strides[4];
  uint32_t box[5];
  int32_t rank, dt, sw;
};
struct TmapSlot {
  TmapKey key;
  CUtensorMap map;
  int valid;
};
constexpr int kTmapSlots = 4096;  // direct-mapped
inline std::mutex& tmap_mutex() {
  static std::mutex mu;
  return mu;
}
inline TmapSlot* tmap_table() {
  static TmapSlot* t = static_cast<TmapSlot*>(calloc(kTmapSlots, sizeof(TmapSlot)));
  return t;
}
inline unsigned long long& tmap_hits() {
  static unsigned long long h = 0;
  return h;
}
inline unsigned long long& tmap_misses() {
  static unsigned long long m = 0;
  return m;
}

inline int make_tmap(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle sw) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base;
  key.rank = rank;
  key.dt = (int32_t)dt;
  key.sw = (int32_t)sw;
  for (int i = 0; i < rank; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides_bytes[i];
  uint64_t h = 1469598103934665603ull;  // FNV-1a over the key's words
  const uint64_t* w = reinterpret_cast<const uint64_t*>(&key);
  for (size_t i = 0; i < sizeof(key) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
  TmapSlot* slot = tmap_table() + (size_t)((h ^ (h >> 29)) % kTmapSlots);
  {
    std::lock_guard<std::mutex> lk(tmap_mutex());
    if (slot->valid && memcmp(&slot->key, &key, sizeof(key)) == 0) {
      memcpy(m, &slot->map, sizeof(CUtensorMap));
      ++tmap_hits();
      return BVC_OK;
    }
  }
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
  {
    std::lock_guard<std::mutex> lk(tmap_mutex());
    slot->key = key;
    memcpy(&slot->map, m, sizeof(CUtensorMap));
    slot->valid = 1;
    ++tmap_misses();
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
