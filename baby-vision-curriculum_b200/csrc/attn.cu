// attn.cu -- placeholder entry points while the tcgen05 attention kernels are brought up (replaced below).
#include "../../include/bvc.h"
#include "bvc_host.h"
extern "C" int bvc_attn_fwd(const void*, int32_t, int32_t, int32_t, float, void*, float*, void*) {
  fprintf(stderr, "bvc: attention kernel not built\n");
  return BVC_ERR_ARG;
}
extern "C" int bvc_attn_bwd(const void*, const void*, const void*, const float*, int32_t, int32_t, int32_t, float,
                            float*, void*, void*) {
  fprintf(stderr, "bvc: attention kernel not built\n");
  return BVC_ERR_ARG;
}
