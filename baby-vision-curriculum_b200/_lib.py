"""ctypes binding of libbvc.so (include/bvc.h).  No fallback: if the library is missing or a call fails, raise.

Every wrapper takes torch CUDA tensors, passes raw device pointers plus the current CUDA stream, and returns nothing;
outputs are pre-allocated by the caller (torch is the allocator / stream owner, not the compute path).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbvc.so")
ABI_VERSION = 15

_lib = None


class BvcError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("lda", C.c_int64), ("ldb", C.c_int64),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("k_splits", C.c_int32),
        ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p), ("ldo", C.c_int64),
        ("out_seg", C.c_int32), ("out_seg_stride", C.c_int32), ("out_seg_off", C.c_int32),
        ("alpha_host", C.c_float), ("alpha_dev", C.c_void_p), ("bias", C.c_void_p),
        ("act", C.c_int32), ("aux_out", C.c_void_p), ("aux_in", C.c_void_p), ("ld_aux", C.c_int64),
        ("res", C.c_void_p), ("ldr", C.c_int64), ("res_idx", C.c_void_p),
        ("target", C.c_void_p), ("ldt", C.c_int64), ("loss_partial", C.c_void_p), ("logits_out", C.c_void_p),
        ("colsum", C.c_void_p), ("block_n", C.c_int32), ("cta_pair", C.c_int32),
    ]


_SIGNATURES = {
    "bvc_abi_version": (C.c_int, []),
    "bvc_mask_count": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bvc_mask_to_index": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "bvc_patchify_target": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 8 + [C.c_void_p, C.c_void_p, C.c_int32,
                                                                                  C.c_void_p]),
    "bvc_patchify_target_u8": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p] +
                               [C.c_int32] * 8 + [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "bvc_gemm_bf16": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "bvc_gemm_loss_slots": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "bvc_loss_finalize": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bvc_layernorm_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bvc_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "bvc_colsum": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                             C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bvc_cast_f32_to_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "bvc_cast_multi": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "bvc_rows_to_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_void_p]),
    "bvc_decoder_mask_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_void_p]),
    "bvc_attn_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "bvc_attn_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                               C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "bvc_nce_normalize_split": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bvc_nce_partial_slots": (C.c_int64, [C.c_int32]),
    "bvc_nce_loss": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "bvc_nce_grad": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "bvc_nce_normalize_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                        C.c_float, C.c_void_p, C.c_void_p]),
    "bvc_sgd_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "bvc_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bvc_grad_nonfinite": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "bvc_jepa_apply_masks": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bvc_jepa_apply_masks_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                           C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bvc_repeat_interleave_batch": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                              C.c_void_p]),
    "bvc_jepa_targets": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bvc_smooth_l1_slots": (C.c_int64, [C.c_int64]),
    "bvc_smooth_l1_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]),
    "bvc_smooth_l1_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "bvc_ema_update": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load():
    """Load libbvc.so once; raise BvcError (never fall back) when it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BvcError(f"{LIB_PATH} not found: build it with `python baby-vision-curriculum_b200/build.py` "
                       "(there is no CPU or PyTorch fallback for the CUDA path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == missing export
        fn.restype = res
        fn.argtypes = args
    v = lib.bvc_abi_version()
    if v != ABI_VERSION:
        raise BvcError(f"libbvc.so ABI {v} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _check(rc, what):
    if rc != 0:
        raise BvcError(f"{what} failed with code {rc} (see stderr)")


_launches = 0
_prof = None  # when a list: (kind, flops, bytes, start_event, end_event) per call, recorded on the current stream


def set_profiler(records):
    """bench.py: pass a list to time every libbvc.so call with CUDA events on its own stream; None to stop."""
    global _prof
    _prof = records


class _Timed:
    __slots__ = ("kind", "flops", "bytes", "s", "detail")

    def __init__(self, kind, flops=0.0, nbytes=0.0, detail=""):
        self.kind, self.flops, self.bytes, self.detail = kind, flops, nbytes, detail

    def __enter__(self):
        if _prof is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if _prof is not None and exc[0] is None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            _prof.append((self.kind, self.flops, self.bytes, self.s, e, self.detail))
        return False


def launch_count():
    """Number of libbvc.so kernel-launching calls made so far in this process (bench.py's gpu_launches)."""
    return _launches


def _count(n=1):
    global _launches
    _launches += n


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise BvcError("libbvc.so kernels need CUDA tensors; there is no CPU path")


# ------------------------------------------------------------------------------------------------ wrappers
def mask_count(mask_u8, n_visible):
    _cuda(mask_u8, n_visible)
    B, N = mask_u8.shape
    with _Timed("mask", 0.0, float(B * N)):
        _check(load().bvc_mask_count(_p(mask_u8), B, N, _p(n_visible), _stream()), "bvc_mask_count")
    _count()


def mask_to_index(mask_u8, nv, vis_idx, msk_idx, slot, status):
    _cuda(mask_u8, vis_idx, msk_idx, slot, status)
    B, N = mask_u8.shape
    with _Timed("mask", 0.0, float(B * N * 9)):
        _check(load().bvc_mask_to_index(_p(mask_u8), B, N, nv, _p(vis_idx), _p(msk_idx), _p(slot), _p(status), _stream()),
               "bvc_mask_to_index")
    _count()


def patchify_target(pixels, slot, ts, ps, nv, patches_vis, target, norm_pix=True, pixel_norm=None):
    """pixels fp32 [B,T,C,H,W], or uint8 with pixel_norm = (mean[3], std[3]) (the dataset's Normalize, applied in-kernel)."""
    _cuda(pixels, slot, patches_vis, target)
    B, T, Cc, H, W = pixels.shape
    nbytes = float(pixels.numel() * pixels.element_size() + patches_vis.numel() * 2 + target.numel() * 4)
    with _Timed("patchify_target", 0.0, nbytes):
        if pixels.dtype == torch.uint8:
            if pixel_norm is None:
                raise BvcError("uint8 pixels need pixel_norm=(mean, std)")
            mean = (C.c_float * 3)(*[float(v) for v in pixel_norm[0]])
            std = (C.c_float * 3)(*[float(v) for v in pixel_norm[1]])
            _check(load().bvc_patchify_target_u8(_p(pixels), mean, std, _p(slot), B, T, Cc, H, W, ts, ps, nv,
                                                 _p(patches_vis), _p(target), 1 if norm_pix else 0, _stream()),
                   "bvc_patchify_target_u8")
        else:
            _check(load().bvc_patchify_target(_p(pixels), _p(slot), B, T, Cc, H, W, ts, ps, nv, _p(patches_vis),
                                              _p(target), 1 if norm_pix else 0, _stream()), "bvc_patchify_target")
    _count()


def gemm(a, b, M, N, K, *, lda=None, ldb=None, a_mn=False, b_mn=False, out_f32=None, out_bf16=None, ldo=None,
         out_seg=0, out_seg_stride=0, out_seg_off=0, alpha=1.0, alpha_dev=None, bias=None, act=0, aux_out=None,
         aux_in=None, ld_aux=0, res=None, ldr=0, res_idx=None, target=None, ldt=0, loss_partial=None,
         logits_out=None, k_splits=1, block_n=0, colsum=None, cta_pair=0):
    """out = epilogue(alpha * A . B^T); see include/bvc.h bvc_gemm_bf16 for the epilogue order."""
    _cuda(a, b, out_f32, out_bf16)
    g = GemmArgs()
    g.a, g.b = a.data_ptr(), b.data_ptr()
    g.lda = lda if lda is not None else (M if a_mn else K)
    g.ldb = ldb if ldb is not None else (N if b_mn else K)
    g.a_mn_major, g.b_mn_major = int(a_mn), int(b_mn)
    g.M, g.N, g.K, g.k_splits = M, N, K, k_splits
    g.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
    g.out_bf16 = out_bf16.data_ptr() if out_bf16 is not None else None
    g.ldo = ldo if ldo is not None else N
    g.out_seg, g.out_seg_stride, g.out_seg_off = out_seg, out_seg_stride, out_seg_off
    g.alpha_host = alpha
    g.alpha_dev = alpha_dev.data_ptr() if alpha_dev is not None else None
    g.bias = bias.data_ptr() if bias is not None else None
    g.act = act
    g.aux_out = aux_out.data_ptr() if aux_out is not None else None
    g.aux_in = aux_in.data_ptr() if aux_in is not None else None
    g.ld_aux = ld_aux
    g.res = res.data_ptr() if res is not None else None
    g.ldr = ldr
    g.res_idx = res_idx.data_ptr() if res_idx is not None else None
    g.target = target.data_ptr() if target is not None else None
    g.ldt = ldt
    g.loss_partial = loss_partial.data_ptr() if loss_partial is not None else None
    g.logits_out = logits_out.data_ptr() if logits_out is not None else None
    g.block_n = block_n
    g.cta_pair = cta_pair
    g.colsum = colsum.data_ptr() if colsum is not None else None
    detail = ""
    if _prof is not None:
        detail = (f"M{M} N{N} K{K} {'T' if a_mn else 'N'}{'T' if b_mn else 'N'}"
                  f"{' gelu' if act == 1 else ' dgelu' if act == 2 else ''}{' res' if res is not None else ''}"
                  f"{' loss' if target is not None else ''}{' f32' if out_f32 is not None else ''}"
                  f"{' bf16' if out_bf16 is not None else ''}{' splitk' if k_splits != 1 else ''}")
    with _Timed("gemm", 2.0 * M * N * K, 0.0, detail):
        _check(load().bvc_gemm_bf16(C.byref(g), _stream()), "bvc_gemm_bf16")
    _count()


def gemm_loss_slots(M, N, block_n=0):
    return int(load().bvc_gemm_loss_slots(M, N, block_n))


def loss_finalize(partials, numel, status, loss):
    _cuda(partials, loss)
    with _Timed("loss_finalize", 0.0, float(partials.numel() * 4)):
        _check(load().bvc_loss_finalize(_p(partials), partials.numel(), float(numel), _p(status), _p(loss), _stream()),
               "bvc_loss_finalize")
    _count()


def layernorm_fwd(x, gamma, beta, eps, M, d, y, mean, rstd, ldx=None, seg=(0, 0, 0)):
    _cuda(x, gamma, beta, y, mean, rstd)
    with _Timed("layernorm_fwd", 0.0, float(M) * d * 6):
        _check(load().bvc_layernorm_fwd(_p(x), ldx if ldx is not None else d, seg[0], seg[1], seg[2], _p(gamma), _p(beta),
                                        eps, M, d, _p(y), _p(mean), _p(rstd), _stream()), "bvc_layernorm_fwd")
    _count()


def layernorm_bwd(dy, x, mean, rstd, gamma, dres, M, d, dx_f32, dx_bf16, dgamma, dbeta, ldx=None, seg=(0, 0, 0),
                  dxsum=None):
    _cuda(dy, x, mean, rstd, gamma, dgamma, dbeta)
    with _Timed("layernorm_bwd", 0.0, float(M) * d * (2 + 4 + (4 if dres is not None else 0) + (4 if dx_f32 is not None else 0) + (2 if dx_bf16 is not None else 0)), f"M{M} d{d}"):
        _check(load().bvc_layernorm_bwd(_p(dy), _p(x), ldx if ldx is not None else d, seg[0], seg[1], seg[2], _p(mean),
                                        _p(rstd), _p(gamma), _p(dres), M, d, _p(dx_f32), _p(dx_bf16), _p(dgamma),
                                        _p(dbeta), _p(dxsum), _stream()), "bvc_layernorm_bwd")
    _count()


def colsum(inp, M, N, out, ld=None, seg=(0, 0, 0), scale=1.0, scale_dev=None):
    _cuda(inp, out)
    is_f32 = 1 if inp.dtype == torch.float32 else 0
    with _Timed("colsum", 0.0, float(M) * N * inp.element_size(), f"M{M} N{N}"):
        _check(load().bvc_colsum(_p(inp), is_f32, ld if ld is not None else N, seg[0], seg[1], seg[2], M, N, scale,
                                 _p(scale_dev), _p(out), _stream()), "bvc_colsum")
    _count()


def cast_bf16(src, dst):
    _cuda(src, dst)
    with _Timed("cast", 0.0, float(src.numel()) * 6):
        _check(load().bvc_cast_f32_to_bf16(_p(src), _p(dst), src.numel(), _stream()), "bvc_cast_f32_to_bf16")
    _count()


def cast_multi(table, n_entries, total_elems):
    _cuda(table)
    with _Timed("cast", 0.0, float(total_elems) * 6):
        _check(load().bvc_cast_multi(_p(table), n_entries, _stream()), "bvc_cast_multi")
    _count()


def rows_to_bf16(src, M, d, dst, ld=None, seg=(0, 0, 0)):
    _cuda(src, dst)
    with _Timed("rows_to_bf16", 0.0, float(M) * d * 6):
        _check(load().bvc_rows_to_bf16(_p(src), ld if ld is not None else d, seg[0], seg[1], seg[2], M, d, _p(dst),
                                       _stream()), "bvc_rows_to_bf16")
    _count()


def decoder_mask_rows(x, mask_token, pos, msk_idx, B, N, nv, d):
    _cuda(x, mask_token, pos, msk_idx)
    with _Timed("decoder_mask_rows", 0.0, float(B) * (N - nv) * d * 8):
        _check(load().bvc_decoder_mask_rows(_p(x), _p(mask_token), _p(pos), _p(msk_idx), B, N, nv, d, _stream()),
               "bvc_decoder_mask_rows")
    _count()


def attn_fwd(qkv, B, S, H, scale, out, lse):
    _cuda(qkv, out, lse)
    with _Timed("attn_fwd", 4.0 * B * H * S * S * 64, 0.0, f"B{B} S{S} H{H}"):
        _check(load().bvc_attn_fwd(_p(qkv), B, S, H, scale, _p(out), _p(lse), _stream()), "bvc_attn_fwd")
    _count()


def attn_bwd(qkv, out, dout, lse, B, S, H, scale, delta, dqkv, dq_accum=None, dq_accum_zeroed=False):
    """dq_accum: optional fp32 [B, S, H, 64] scratch; with it sequences > 160 tokens run the one-pass backward.
    dq_accum_zeroed: the caller already cleared it (stream-ordered before this call)."""
    _cuda(qkv, out, dout, lse, delta, dqkv, dq_accum)
    with _Timed("attn_bwd", 8.0 * B * H * S * S * 64, 0.0, f"B{B} S{S} H{H}"):
        _check(load().bvc_attn_bwd(_p(qkv), _p(out), _p(dout), _p(lse), B, S, H, scale, _p(delta), _p(dqkv),
                                   _p(dq_accum), int(bool(dq_accum_zeroed)), _stream()), "bvc_attn_bwd")
    _count(3)


def sgd_step(table, n_entries, total_elems, lr, momentum, dampening, weight_decay, nesterov, grad_scale, found_inf,
             bytes_per_elem):
    _cuda(table, grad_scale, found_inf)
    with _Timed("sgd_step", 0.0, float(total_elems) * bytes_per_elem):
        _check(load().bvc_sgd_step(_p(table), n_entries, lr, momentum, dampening, weight_decay, 1 if nesterov else 0,
                                   _p(grad_scale), _p(found_inf), _stream()), "bvc_sgd_step")
    _count()


def adam_step(table, v_table, n_entries, total_elems, lr, beta1, beta2, eps, weight_decay, decoupled, step, grad_scale,
              found_inf, bytes_per_elem):
    _cuda(table, v_table, step, grad_scale, found_inf)
    with _Timed("adam_step", 0.0, float(total_elems) * bytes_per_elem):
        _check(load().bvc_adam_step(_p(table), _p(v_table), n_entries, lr, beta1, beta2, eps, weight_decay,
                                    1 if decoupled else 0, _p(step), _p(grad_scale), _p(found_inf), _stream()),
               "bvc_adam_step")
    _count(2)


def grad_nonfinite(table, n_entries, total_elems, found_inf):
    _cuda(table, found_inf)
    with _Timed("grad_nonfinite", 0.0, float(total_elems) * 4):
        _check(load().bvc_grad_nonfinite(_p(table), n_entries, _p(found_inf), _stream()), "bvc_grad_nonfinite")
    _count()


def nce_normalize_split(feats, n, D, eps, a_split, b_split, bk_split, inv_norm):
    _cuda(feats, a_split, b_split, bk_split, inv_norm)
    is_bf16 = 1 if feats.dtype == torch.bfloat16 else 0
    with _Timed("nce_rows", 0.0, float(n) * D * (feats.element_size() + 18)):
        _check(load().bvc_nce_normalize_split(_p(feats), is_bf16, feats.stride(0), n, D, eps, _p(a_split), _p(b_split),
                                              _p(bk_split), _p(inv_norm), _stream()), "bvc_nce_normalize_split")
    _count()


def nce_partial_slots(n):
    return int(load().bvc_nce_partial_slots(n))


def nce_loss(S, pos_u8, neg_u8, n, partials, out4):
    _cuda(S, pos_u8, neg_u8, partials, out4)
    with _Timed("nce_loss", 0.0, float(n) * n * 6):
        _check(load().bvc_nce_loss(_p(S), S.stride(0), _p(pos_u8), _p(neg_u8), n, _p(partials), _p(out4), _stream()),
               "bvc_nce_loss")
    _count(2)


def nce_grad(S, pos_u8, neg_u8, n, out4, grad_out, g_split):
    _cuda(S, pos_u8, neg_u8, out4, grad_out, g_split)
    with _Timed("nce_grad", 0.0, float(n) * n * 14):
        _check(load().bvc_nce_grad(_p(S), S.stride(0), _p(pos_u8), _p(neg_u8), n, _p(out4), _p(grad_out), _p(g_split),
                                   _stream()), "bvc_nce_grad")
    _count()


def nce_normalize_bwd(dfhat, feats, inv_norm, n, D, eps, dfeats):
    _cuda(dfhat, feats, inv_norm, dfeats)
    is_bf16 = 1 if feats.dtype == torch.bfloat16 else 0
    with _Timed("nce_rows", 0.0, float(n) * D * (feats.element_size() + 8)):
        _check(load().bvc_nce_normalize_bwd(_p(dfhat), _p(feats), is_bf16, feats.stride(0), _p(inv_norm), n, D, eps,
                                            _p(dfeats), _stream()), "bvc_nce_normalize_bwd")
    _count()


# ------------------------------------------------------------------------------------------------ JEPA pieces
def jepa_apply_masks(x, idx, B, N, D, n_masks, K, repeat, out, status=None):
    _cuda(x, idx, out)
    nbytes = float(n_masks) * B * K * D * x.element_size() * (1 + repeat)
    with _Timed("jepa_apply_masks", 0.0, nbytes):
        _check(load().bvc_jepa_apply_masks(_p(x), x.element_size(), B, N, D, _p(idx), n_masks, K, repeat, _p(out),
                                           _p(status), _stream()), "bvc_jepa_apply_masks")
    _count()


def jepa_apply_masks_bwd(dy, idx, B, N, D, n_masks, K, dx):
    _cuda(dy, idx, dx)
    with _Timed("jepa_apply_masks_bwd", 0.0, float(dx.numel() + 3 * dy.numel()) * dy.element_size()):
        _check(load().bvc_jepa_apply_masks_bwd(_p(dy), dy.element_size(), B, N, D, _p(idx), n_masks, K, _p(dx),
                                               _stream()), "bvc_jepa_apply_masks_bwd")
    _count(n_masks)


def repeat_interleave_batch(x, slab_bytes, B, n_groups, repeat, out):
    _cuda(x, out)
    with _Timed("repeat_interleave_batch", 0.0, float(n_groups) * B * slab_bytes * (1 + repeat)):
        _check(load().bvc_repeat_interleave_batch(_p(x), slab_bytes, B, n_groups, repeat, _p(out), _stream()),
               "bvc_repeat_interleave_batch")
    _count()


def jepa_targets(h, idx, B, N, D, n_masks, K, repeat, eps, out, status=None):
    _cuda(h, idx, out)
    nbytes = float(n_masks) * B * K * D * (h.element_size() + 4 * repeat)
    with _Timed("jepa_targets", 0.0, nbytes):
        _check(load().bvc_jepa_targets(_p(h), 1 if h.dtype == torch.float32 else 0, B, N, D, _p(idx), n_masks, K,
                                       repeat, eps, _p(out), _p(status), _stream()), "bvc_jepa_targets")
    _count()


def smooth_l1_slots(n):
    return int(load().bvc_smooth_l1_slots(n))


def smooth_l1_fwd(z, h, n, beta, partials):
    _cuda(z, h, partials)
    with _Timed("smooth_l1_fwd", 0.0, float(n) * (z.element_size() + 4)):
        _check(load().bvc_smooth_l1_fwd(_p(z), 1 if z.dtype == torch.float32 else 0, _p(h), n, beta, _p(partials),
                                        _stream()), "bvc_smooth_l1_fwd")
    _count()


def smooth_l1_bwd(z, h, n, beta, grad_out, dz):
    _cuda(z, h, grad_out, dz)
    with _Timed("smooth_l1_bwd", 0.0, float(n) * (2 * z.element_size() + 4)):
        _check(load().bvc_smooth_l1_bwd(_p(z), 1 if z.dtype == torch.float32 else 0, _p(h), n, beta, _p(grad_out),
                                        _p(dz), _stream()), "bvc_smooth_l1_bwd")
    _count()


def ema_update(table, n_entries, momentum, total_elems):
    _cuda(table)
    with _Timed("ema_update", 0.0, float(total_elems) * 12):
        _check(load().bvc_ema_update(_p(table), n_entries, float(momentum), _stream()), "bvc_ema_update")
    _count()


