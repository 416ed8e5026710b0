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
 *   - bf16 = __nv_bfloat16 storage (uint16), row-major, leading dimensions in ELEMENTS;
 *   - "segment remap" (seg, seg_stride, seg_off): logical row r of a compact [rows, d] operand lives at physical
 *     row (r / seg) * seg_stride + (r % seg) + seg_off of a [B * seg_stride, d] buffer; seg == 0 means identity.
 *     It expresses "the Nv visible rows" / "the last Nm rows" of every clip of a [B, N, d] tensor (HF:506, HF:591).
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
#define BVC_ABI_VERSION 15

/* library / build info: returns BVC_ABI_VERSION (bumped when a signature changes) */
int bvc_abi_version(void);

/* ------------------------------------------------------------------------------------------------------
 * Tube-mask indexing.  Replaces the boolean-index ops `x[~bool_masked_pos]` / `x[bool_masked_pos]`
 * (HF:121-122, HF:587-588, HF:669-670; mask built at pretrain_videomae.py:294-298).  Bit-exact integer work.
 *   mask        uint8/bool [B, N], non-zero = masked
 *   bvc_mask_count:    n_visible[b] = #zeros of row b
 *   bvc_mask_to_index: vis_idx [B, nv] / msk_idx [B, N-nv] = ascending token ids of the visible / masked tokens;
 *                      slot [B, N] = position of token n in the decoder sequence [visible asc ; masked asc]
 *                      (slot < nv <=> visible); status[0] |= 1 if some row does not have exactly nv visible
 *                      tokens (HF's reshape raises in that case; the caller turns the flag into an error / NaN loss)
 * ------------------------------------------------------------------------------------------------------ */
int bvc_mask_count(const uint8_t* mask, int32_t B, int32_t N, int32_t* n_visible, void* stream);
int bvc_mask_to_index(const uint8_t* mask, int32_t B, int32_t N, int32_t nv, int32_t* vis_idx, int32_t* msk_idx,
                      int32_t* slot, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Tubelet patchify + normalised-pixel target in one HBM pass (TMA-staged rows of patches).
 * Replaces the Conv3d im2col/cast of HF:175-176 (visible tokens only: gather-first, HF:119-122) and the whole
 * label construction HF:598-670.
 *   pixels      fp32 [B, T, C=3, H, W] contiguous; patch 16x16, tubelet ts in {1, 2}; W <= 256
 *   slot        from bvc_mask_to_index
 *   patches_vis bf16 [B*nv, K]      K = 3*ts*256, k = ((c*ts + t)*16 + ph)*16 + pw   (Conv3d weight order)
 *   target      fp32 [B*(N-nv), K]  f = ((t*16 + ph)*16 + pw)*3 + c; if norm_pix:
 *               (p - mean) / (sqrt(var_unbiased) + 1e-6) per (token, channel) with p = x*std_IN[c] + mean_IN[c],
 *               else p itself (HF:644-667)
 * ------------------------------------------------------------------------------------------------------ */
int bvc_patchify_target(const float* pixels, const int32_t* slot, int32_t B, int32_t T, int32_t C, int32_t H,
                        int32_t W, int32_t ts, int32_t ps, int32_t nv, void* patches_vis, float* target,
                        int32_t norm_pix, void* stream);
/* Same pass on uint8 frames [B,T,C,H,W]: the dataset's ToTensor + Normalize
 * (pretraining/generative/homeview.py:218-231: x / 255, then (x - mean[c]) / std[c]) is applied in registers with IEEE
 * divisions in torchvision's operation order, so the outputs are bit-identical to normalising on the host and calling
 * bvc_patchify_target; the clip crosses PCIe and HBM at a quarter of the bytes.  mean3 / std3 are HOST pointers. */
int bvc_patchify_target_u8(const uint8_t* pixels, const float* mean3, const float* std3, const int32_t* slot, int32_t B,
                           int32_t T, int32_t C, int32_t H, int32_t W, int32_t ts, int32_t ps, int32_t nv,
                           void* patches_vis, float* target, int32_t norm_pix, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Dense contraction on the tcgen05 tensor cores:   out[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T )
 * Replaces every F.linear / Conv3d-as-GEMM on the path: HF:175-176 (tubelet embedding), HF:239-242 (q,k,v),
 * HF:281-284 (attention out-proj), HF:314-317 (fc1+GELU), HF:327-331 (fc2+residual), HF:576 (encoder_to_decoder),
 * HF:510 (decoder head) + HF:672-673 (MSE) and their autograd backward (dgrad / wgrad).
 *
 * Operand storage: a_mn_major == 0: A is stored [M, K] row-major (K contiguous), lda >= K.
 *                  a_mn_major == 1: A is stored [K, M] row-major (M contiguous), lda >= M  (i.e. A^T in memory).
 *                  same for B with N.   Forward = (0,0); dgrad dX = dY.W = (0,1); wgrad dW = dY^T.X = (1,1).
 * Requirements: lda, ldb, ldo, ld_aux multiples of 8; N multiple of 8; pointers 16-byte aligned.
 * Epilogue, in this order, per output element (r, c):
 *     v = alpha_host * (alpha_dev ? *alpha_dev : 1) * acc + (bias ? bias[c] : 0)
 *     act == 1 (GELU, exact erf, HF:316):  v = bf16(v); if aux_out: aux_out[r*ld_aux+c] = bf16(gelu'(v));  v = gelu(v)
 *     act == 2 (GELU backward):            v *= aux_in[r*ld_aux+c]      (the gelu' saved by the forward call: the
 *                                          backward epilogue carries no transcendental math)
 *     if res:     v += res[(res_idx ? res_idx[r] : r) * ldr + c]          (fp32 residual / position table)
 *     if target:  (masked MSE, HF:672-673)  if logits_out: logits_out[r*ldo+c] = bf16(v);
 *                 v -= target[r*ldt+c];  tile partial of sum(v*v) -> loss_partial (see bvc_gemm_loss_slots)
 *     if colsum:  colsum[c] += v  (fp32 atomics; caller zero-fills) -- the bias gradient of the Linear whose output
 *                 gradient this GEMM produces, without a second pass over it (not with k_splits > 1 / target)
 *     R = segment remap of r with (out_seg, out_seg_stride, out_seg_off)
 *     store:      k_splits > 1 -> atomicAdd(out_f32[R*ldo+c], v)   (caller zero-fills out_f32; no bias/act/res)
 *                 else out_f32[R*ldo+c] = v and/or out_bf16[R*ldo+c] = bf16(v)
 * ------------------------------------------------------------------------------------------------------ */
typedef struct bvc_gemm_args {
  const void* a;          /* bf16 */
  const void* b;          /* bf16 */
  int64_t lda, ldb;
  int32_t a_mn_major, b_mn_major;
  int32_t M, N, K;
  int32_t k_splits;       /* >= 1; > 1 only with out_f32 (atomic accumulate); 0 = choose (wgrad shapes) */
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
  float* loss_partial;    /* fp32 [bvc_gemm_loss_slots(M, N, block_n)], every slot written by the call */
  void* logits_out;       /* bf16, ld = ldo */
  float* colsum;          /* fp32 [N] or null */
  int32_t block_n;        /* 0 = choose; else 64 / 128 / 192 / 256 */
  int32_t cta_pair;       /* 0 = choose; 1 = one CTA per 128 x block_n tile; 2 = CTA pair (tcgen05 cta_group::2) on
                             256 x block_n tiles (block_n 128 / 256, or 192 with a K-major B; else BVC_ERR_ARG) */
} bvc_gemm_args;

int bvc_gemm_bf16(const bvc_gemm_args* args, void* stream);
/* number of fp32 partial sums a target/loss GEMM of this shape writes (block_n as passed to the call) */
int64_t bvc_gemm_loss_slots(int32_t M, int32_t N, int32_t block_n);
/* loss = sum(partials) / numel (fixed order, double accumulation: deterministic); NaN when status[0] != 0 */
int bvc_loss_finalize(const float* partials, int64_t n, double numel, const int32_t* status, float* loss,
                      void* stream);

/* ------------------------------------------------------------------------------------------------------
 * LayerNorm (HF:352, HF:360 eps 1e-12; decoder.norm HF:508 eps 1e-5), fp32 statistics, one warp per row.
 *   x fp32 [*, d] with segment remap (x_seg...) ; y bf16 [M, d] compact; mean/rstd fp32 [M].  d % 4 == 0, d <= 1024.
 * Backward: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma  (+ dres if given);
 *   dx_f32 / dx_bf16 (either may be null) are written at the remapped rows, dres read at the remapped rows;
 *   dgamma/dbeta fp32 [d] are ACCUMULATED atomically (caller zero-fills); dxsum (optional) accumulates the column
 *   sums of the produced dx the same way -- it IS the bias gradient of the Linear whose output this LayerNorm's
 *   input stream received (fc2 / attention out-proj / patch embedding), saving a separate pass.
 * ------------------------------------------------------------------------------------------------------ */
int bvc_layernorm_fwd(const float* x, int64_t ldx, int32_t x_seg, int32_t x_seg_stride, int32_t x_seg_off,
                      const float* gamma, const float* beta, float eps, int32_t M, int32_t d, void* y,
                      float* mean, float* rstd, void* stream);
int bvc_layernorm_bwd(const void* dy, const float* x, int64_t ldx, int32_t x_seg, int32_t x_seg_stride,
                      int32_t x_seg_off, const float* mean, const float* rstd, const float* gamma,
                      const float* dres, int32_t M, int32_t d, float* dx_f32, void* dx_bf16, float* dgamma,
                      float* dbeta, float* dxsum, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Column sums (bias gradients; mask_token gradient HF:528/591): out[c] += scale * sum_r in[row(r), c].
 *   in_is_f32 selects fp32 / bf16 input; segment remap on the input rows; out fp32 [N] accumulated atomically
 *   (caller zero-fills); scale = scale_host * (scale_dev ? *scale_dev : 1).  N % 8 == 0.
 * ------------------------------------------------------------------------------------------------------ */
int bvc_colsum(const void* in, int32_t in_is_f32, int64_t ld, int32_t seg, int32_t seg_stride, int32_t seg_off,
               int32_t M, int32_t N, float scale_host, const float* scale_dev, float* out, void* stream);

/* fp32 -> bf16 cast of n contiguous elements (per-step weight cast that autocast does per call, HF linear layers) */
int bvc_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
/* Every weight of the model in ONE launch: `table` is a device array of n_entries records
 *   { const float* src; void* dst; int64_t n; int32_t dst_is_f32; int32_t pad; }   (32 bytes each)
 * entry t casts (dst bf16) or copies (dst fp32, used to pack q_bias / v_bias into the fused QKV bias) n elements. */
int bvc_cast_multi(const void* table, int32_t n_entries, void* stream);
/* rows of an fp32 [*, d] buffer (segment remap) -> compact bf16 [M, d]  (d % 4 == 0) */
int bvc_rows_to_bf16(const float* src, int64_t ld, int32_t seg, int32_t seg_stride, int32_t seg_off, int32_t M,
                     int32_t d, void* dst, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Decoder input, mask-token half (HF:588-591): x[b, nv + j, :] = mask_token + pos[msk_idx[b, j], :]
 *   x fp32 [B, N, d]; pos fp32 [N, d]; msk_idx int32 [B, N-nv].  (The visible half is written by the
 *   encoder_to_decoder GEMM epilogue: res = pos, res_idx = vis_idx, out_seg = nv, out_seg_stride = N.)
 * ------------------------------------------------------------------------------------------------------ */
int bvc_decoder_mask_rows(float* x, const float* mask_token, const float* pos, const int32_t* msk_idx, int32_t B,
                          int32_t N, int32_t nv, int32_t d, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Multi-head self-attention, head_dim 64, no mask, no dropout (HF:236-266 + sdpa / HF:181-206), flash-style on
 * tcgen05: S = Q K^T and O = P V accumulate in TMEM, softmax in fp32 registers.
 *   qkv  bf16 [B, S, 3, H, 64]  (= the fused QKV GEMM output [B*S, 3*H*64])
 *   out  bf16 [B, S, H*64];  lse fp32 [B, H, S] (natural-log sum-exp of the scaled scores)
 * Backward: dqkv bf16 [B, S, 3, H, 64] from dout bf16 [B, S, H*64], recomputing P from lse.
 *   delta fp32 [B, H, S] is scratch (rowsum(dO * O)).
 *   dq_accum fp32 [B, S, H, 64] is optional scratch: with it, sequences longer than 160 run the ONE-pass backward
 *   (5 MMAs and one exponential per score; each (key tile, query tile) pair's dQ contribution is added into dq_accum
 *   by TMA reduce and converted into dqkv's q slot afterwards); with NULL the two-pass kernels run (no atomics,
 *   7 MMAs and two exponentials per score).  The call zero-fills dq_accum itself unless dq_accum_zeroed != 0 (the
 *   caller cleared it earlier, e.g. on another stream while the preceding GEMMs ran).
 * ------------------------------------------------------------------------------------------------------ */
int bvc_attn_fwd(const void* qkv, int32_t B, int32_t S, int32_t H, float scale, void* out, float* lse,
                 void* stream);
int bvc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int32_t B, int32_t S,
                 int32_t H, float scale, float* delta, void* dqkv, float* dq_accum, int32_t dq_accum_zeroed,
                 void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Optimizer step (pretrain_videomae.py:187-197 torch.optim.SGD(nesterov) under GradScaler :312-314) as one
 * multi-tensor pass.  `table` is a device array of n_entries records (48 bytes each)
 *   { float* p; float* g; float* m; void* shadow; int64_t n; int32_t shadow_is_f32; int32_t m_uninit; }
 * per entry, with g' = g / *grad_scale (grad_scale may be null):
 *   g' += weight_decay * p;  m = m_uninit ? g' : momentum * m + (1 - dampening) * g';
 *   g'' = nesterov ? g' + momentum * m : m;  p -= lr * g''          (torch/optim/sgd.py _single_tensor_sgd)
 * g is overwritten by the unscaled gradient when grad_scale is given (what loggingtools.py:107-118 reads after
 * scaler.step), shadow (bf16, or fp32 when shadow_is_f32) receives the updated parameter (the operand copy the
 * next forward's GEMMs read), m may be null when momentum == 0.  If found_inf is non-null and *found_inf != 0
 * the whole call is a no-op (GradScaler's skipped step).
 * ------------------------------------------------------------------------------------------------------ */
int bvc_sgd_step(const void* table, int32_t n_entries, float lr, float momentum, float dampening,
                 float weight_decay, int32_t nesterov, const float* grad_scale, const float* found_inf,
                 void* stream);

/* torch.optim.AdamW / Adam (pretrain_videomae.py:190-193: AdamW, betas (0.9, 0.95)) in one multi-tensor pass over
 * the same table (m = exp_avg; m_uninit: both state buffers count as zero); exp_avg_sq_table is a device array of
 * n_entries float* (exp_avg_sq of each entry).  Op order of torch/optim/adam.py _single_tensor_adam:
 *   AdamW (decoupled != 0): p *= 1 - lr wd      Adam: g' += wd p
 *   m = m + (1 - beta1)(g' - m);  v = beta2 v + (1 - beta2) g' g';   t = *step + 1
 *   p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
 * Hyper-parameters are doubles: torch derives 1 - beta, lr / (1 - beta1^t), ... in double and rounds each ONCE to fp32.
 * `step` is a device float (the group's step count before this call); a second one-thread kernel adds 1 to it after
 * the update unless *found_inf != 0, in which case the whole call is a no-op.  g, shadow, grad_scale as bvc_sgd_step. */
int bvc_adam_step(const void* table, const void* exp_avg_sq_table, int32_t n_entries, double lr, double beta1,
                  double beta2, double eps, double weight_decay, int32_t decoupled, float* step,
                  const float* grad_scale, const float* found_inf, void* stream);

/* GradScaler's inf / nan check (torch._amp_foreach_non_finite_check_and_unscale_ with inv_scale 1, which re-writes
 * every gradient) as one read-only multi-tensor launch over the g pointers of the same table:
 * *found_inf = any non-finite gradient element ? 1 : 0. */
int bvc_grad_nonfinite(const void* table, int32_t n_entries, float* found_inf, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * SimCLR loss of the contrastive path (pretraining/contrastive/pretrain_simclr.py:114-128 info_nce_loss with the
 * masks of :86-91 / :284-292).  With S = cos_sim(feats_i, feats_j) / T:
 *     loss = logsumexp(S[neg_mask]) - mean(S[pos_mask])          (ONE global log-sum-exp, as the reference computes)
 * The similarity runs on bvc_gemm_bf16; these entry points are the passes around it (the host side strings them
 * together: baby-vision-curriculum_b200/simclr.py).  n % 8 == 0, D % 8 == 0.
 *   bvc_nce_normalize_split : f^ = feats / max(|feats|, eps) per row, split into bf16 hi + lo;
 *        a_split [n,3D] = [hi|lo|hi], b_split [n,3D] = [hi|hi|lo]  ->  S = (1/T) * a_split . b_split^T (K = 3D)
 *        bk_split [3n,D] = [hi;hi;lo]  (B operand, MN-major, of the backward GEMM);  inv_norm fp32 [n]
 *   bvc_nce_loss  : masked pass over S fp32 [n,n] -> out4 = {loss, lse, mean_pos, pos_count};
 *        partials = scratch of bvc_nce_partial_slots(n) floats.  Masks are uint8 / bool [n,n], arbitrary.
 *   bvc_nce_grad  : g_split [n,3n] = [hi|lo|hi] of  grad_out * ((neg_ij + neg_ji) exp(S_ij - lse) - (pos_ij + pos_ji)/P)
 *        ->  dF^ [n,D] = (1/T) * g_split . bk_split   (bvc_gemm_bf16, b_mn_major = 1, K = 3n)
 *   bvc_nce_normalize_bwd : dfeats = (dF^ - f^ (f^ . dF^)) * inv_norm
 * ------------------------------------------------------------------------------------------------------ */
int bvc_nce_normalize_split(const void* feats, int32_t feats_is_bf16, int64_t ld, int32_t n, int32_t D, float eps,
                            void* a_split, void* b_split, void* bk_split, float* inv_norm, void* stream);
int64_t bvc_nce_partial_slots(int32_t n);
int bvc_nce_loss(const float* S, int64_t lds, const uint8_t* pos_mask, const uint8_t* neg_mask, int32_t n,
                 float* partials, float* out4, void* stream);
int bvc_nce_grad(const float* S, int64_t lds, const uint8_t* pos_mask, const uint8_t* neg_mask, int32_t n,
                 const float* out4, const float* grad_out, void* g_split, void* stream);
int bvc_nce_normalize_bwd(const float* dfhat, const void* feats, int32_t feats_is_bf16, int64_t ld,
                          const float* inv_norm, int32_t n, int32_t D, float eps, float* dfeats, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Predictive (JEPA) path, the pieces beside its ViT blocks.  idx = the list of masks
 * stacked: int64 [n_masks, B, K] indices of kept patches in [0, N) (what MaskCollator / update_masks produce,
 * pretraining/predictive/mask.py:21-38, :161-219).  An index outside [0, N) sets status[0] = 1 (if given) and yields a
 * zero row (torch.gather raises; the host mirror raises when it checks the flag).
 *   bvc_jepa_apply_masks     : mask.py:58-67 apply_masks -- out[(i*repeat + r)*B + b, k, :] = x[b, idx[i,b,k], :]
 *                              (repeat = 1 is apply_masks itself; repeat > 1 also applies tensors.py:65-71
 *                              repeat_interleave_batch(., B, repeat) to the result).  elem_bytes 2 / 4, D*elem_bytes % 16 == 0.
 *   bvc_jepa_apply_masks_bwd : its autograd backward -- dx zero-filled, then dx[b, idx[i,b,k], :] += dy[i*B + b, k, :]
 *                              mask after mask, LAST mask first (autograd's accumulation order; bit-exact, no atomics;
 *                              requires the indices of ONE mask to be unique per sample, which block masks are)
 *   bvc_repeat_interleave_batch : tensors.py:65-71 on x [n_groups*B, slab] -> out [n_groups*repeat*B, slab]
 *   bvc_jepa_targets         : pretrain_jepa.py:384-392 in one pass -- F.layer_norm(h, (D,)) (eps, no affine, fp32
 *                              statistics), apply_masks(., masks_pred), repeat_interleave_batch(., B, repeat):
 *                              out fp32 [(n_masks*repeat*B), K, D].  Only the gathered rows are normalised.  D <= 1024.
 *   bvc_smooth_l1_fwd / bwd  : F.smooth_l1_loss(z, h) (pretrain_jepa.py:399-402; mean reduction, beta) --
 *                              fwd writes bvc_smooth_l1_slots(n) partial sums, bvc_loss_finalize(partials, slots, n, ...)
 *                              makes the mean; bwd: dz = grad_out[0] * (|d| < beta ? d/beta : sign d) / n in z's dtype
 *   bvc_ema_update           : pretrain_jepa.py:426-432 for every parameter in one launch; table = device array of
 *                              { float* dst; const float* src; int64_t n; } (24 bytes each):
 *                              dst = fl(fl(m*dst) + fl((1-m)*src)) with m, (1-m) rounded to fp32 as torch does
 * ------------------------------------------------------------------------------------------------------ */
int bvc_jepa_apply_masks(const void* x, int32_t elem_bytes, int32_t B, int32_t N, int32_t D, const int64_t* idx,
                         int32_t n_masks, int32_t K, int32_t repeat, void* out, int32_t* status, void* stream);
int bvc_jepa_apply_masks_bwd(const void* dy, int32_t elem_bytes, int32_t B, int32_t N, int32_t D, const int64_t* idx,
                             int32_t n_masks, int32_t K, void* dx, void* stream);
int bvc_repeat_interleave_batch(const void* x, int64_t slab_bytes, int32_t B, int32_t n_groups, int32_t repeat,
                                void* out, void* stream);
int bvc_jepa_targets(const void* h, int32_t h_is_f32, int32_t B, int32_t N, int32_t D, const int64_t* idx,
                     int32_t n_masks, int32_t K, int32_t repeat, float eps, float* out, int32_t* status, void* stream);
int64_t bvc_smooth_l1_slots(int64_t n);
int bvc_smooth_l1_fwd(const void* z, int32_t z_is_f32, const float* h, int64_t n, float beta, float* partials,
                      void* stream);
int bvc_smooth_l1_bwd(const void* z, int32_t z_is_f32, const float* h, int64_t n, float beta, const float* grad_out,
                      void* dz, void* stream);
int bvc_ema_update(const void* table, int32_t n_entries, double momentum, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BVC_H_ */
