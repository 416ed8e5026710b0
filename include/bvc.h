/* bvc.h -- C ABI of libbvc.so: the sm_100a kernels behind the VideoMAE pretraining step.
 *
 * The reference (ssheybani/baby-vision-curriculum) has no FFI: its hot path is the Python call
 *     outputs = xmodel(inputs, bool_masked_pos=bool_masked_pos)      pretraining/generative/pretrain_videomae.py:301
 * whose arithmetic lives in HuggingFace transformers 5.5.0, transformers/models/videomae/modeling_videomae.py
 * (cited below as HF:<line>).  Each entry point names the reference lines it replaces.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer owned by the caller (a torch tensor's data_ptr()); the caller keeps it
 *     alive until `stream` has passed the call;
 *   - returns 0 on success, <0 on error (BVC_ERR_*); never throws, never allocates, never synchronises;
 *   - `stream` is a cudaStream_t passed as void*; re-entrant across streams and host threads (backward runs on
 *     the autograd engine's thread);
 *   - bf16 = __nv_bfloat16 storage (uint16), row-major, leading dimensions in ELEMENTS.
 */
#ifndef BVC_H_
#define BVC_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVC_OK 0
#define BVC_ERR_ARG (-1)
#define BVC_ERR_DRIVER (-2)
#define BVC_ERR_LAUNCH (-3)

/* library / build info: returns the ABI version (bumped when a signature changes) */
int bvc_abi_version(void);

/* ------------------------------------------------------------------------------------------------------
 * Dense contraction on the tcgen05 tensor cores:   out[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T )
 * Replaces every F.linear / Conv3d-as-GEMM on the path: HF:175-176 (tubelet embedding), HF:239-242 (q,k,v),
 * HF:281-284 (attention out-proj), HF:314-317 (fc1+GELU), HF:327-331 (fc2+residual), HF:576 (encoder_to_decoder),
 * HF:510 (decoder head) and their autograd backward (dgrad / wgrad).
 *
 * Operand storage: a_mn_major == 0: A is stored [M, K] row-major (K contiguous), lda >= K.
 *                  a_mn_major == 1: A is stored [K, M] row-major (M contiguous), lda >= M  (i.e. A^T in memory).
 *                  same for B with N.   Forward = (0,0); dgrad dX = dY.W = (0,1); wgrad dW = dY^T.X = (1,1).
 * Epilogue, in this order, per output element (r, c):
 *     v = alpha_host * (alpha_dev ? *alpha_dev : 1) * acc + (bias ? bias[c] : 0)
 *     act == 1 (GELU, exact erf, HF:316):  if aux_out: aux_out[r*ld_aux+c] = bf16(v);  v = gelu(v)
 *     act == 2 (GELU backward):            v *= gelu'(aux_in[r*ld_aux+c])
 *     if res:     v += res[(res_idx ? res_idx[r] : r) * ldr + c]          (fp32 residual / position table)
 *     if target:  (masked-MSE, HF:672-673)  if logits_out: logits_out[r*ldo+c] = bf16(v);
 *                 v -= target[r*ldt+c];  *loss_acc += v*v  (double, atomically, one add per warp)
 *     row remap:  R = out_seg > 0 ? (r / out_seg) * out_seg_stride + (r % out_seg) + out_seg_off : r
 *     store:      k_splits > 1 -> atomicAdd(out_f32[R*ldo+c], v)   (caller zero-fills out_f32; no bias/act/res)
 *                 else out_f32[R*ldo+c] = v and/or out_bf16[R*ldo+c] = bf16(v)
 * ------------------------------------------------------------------------------------------------------ */
typedef struct bvc_gemm_args {
  const void* a;          /* bf16 */
  const void* b;          /* bf16 */
  int64_t lda, ldb;
  int32_t a_mn_major, b_mn_major;
  int32_t M, N, K;
  int32_t k_splits;       /* >= 1; > 1 only with out_f32 (atomic accumulate) */
  float* out_f32;
  void* out_bf16;
  int64_t ldo;
  int32_t out_seg, out_seg_stride, out_seg_off;
  float alpha_host;
  const float* alpha_dev;
  const float* bias;
  int32_t act;
  void* aux_out;          /* bf16 */
  const void* aux_in;     /* bf16 */
  int64_t ld_aux;
  const float* res;
  int64_t ldr;
  const int32_t* res_idx;
  const float* target;
  int64_t ldt;
  double* loss_acc;
  void* logits_out;       /* bf16, ld = ldo */
  int32_t block_n;        /* 0 = choose; else 64 / 128 / 192 / 256 */
} bvc_gemm_args;

int bvc_gemm_bf16(const bvc_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BVC_H_ */
