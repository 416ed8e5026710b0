"""The predictive (JEPA) path's non-ViT pieces on libbvc.so -- drop-ins, same names and
argument meaning, for

  apply_masks(x, masks)                      pretraining/predictive/mask.py:58-67
  repeat_interleave_batch(x, B, repeat)      pretraining/predictive/tensors.py:65-71
  F.smooth_l1_loss(z, h)                     pretraining/predictive/pretrain_jepa.py:399-402
  the momentum update of the target encoder  pretraining/predictive/pretrain_jepa.py:426-432  -> ema_update

plus `jepa_targets(h, masks_pred, n_enc_masks)`, the target branch of pretrain_jepa.py:384-392 (F.layer_norm ->
apply_masks -> repeat_interleave_batch) as ONE pass that only normalises the rows it gathers.  CUDA only: there is no
CPU path.  Index / copy work is bit-exact against torch; the EMA reproduces torch's three roundings.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib as L

_STRICT = os.environ.get("BVC_STRICT_MASK", "0") == "1"


def _stack_masks(masks, B):
    if isinstance(masks, torch.Tensor):
        masks = [masks]
    if len(masks) == 0:
        raise ValueError("masks must hold at least one index tensor")
    K = masks[0].shape[1]
    for m in masks:
        if m.dim() != 2 or m.shape[0] != B or m.shape[1] != K:
            raise ValueError("every mask must be [B, K] with the same K (the collator truncates them to a common length)")
    idx = torch.stack([m.to(torch.int64) for m in masks], 0).contiguous()   # [n_masks, B, K]
    return idx, len(masks), K


def _check_status(status, what):
    if status is not None and _STRICT and int(status) != 0:
        raise RuntimeError(f"{what}: index out of range (torch.gather would raise)")


class _ApplyMasksFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, n_masks, K):
        B, N, D = x.shape
        out = torch.empty((n_masks * B, K, D), dtype=x.dtype, device=x.device)
        status = torch.zeros(1, dtype=torch.int32, device=x.device) if _STRICT else None
        L.jepa_apply_masks(x, idx, B, N, D, n_masks, K, 1, out, status)
        _check_status(status, "apply_masks")
        ctx.save_for_backward(idx)
        ctx.dims = (B, N, D, n_masks, K)
        return out

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        B, N, D, n_masks, K = ctx.dims
        dy = dy.contiguous()
        dx = torch.empty((B, N, D), dtype=dy.dtype, device=dy.device)
        L.jepa_apply_masks_bwd(dy, idx, B, N, D, n_masks, K, dx)
        return dx, None, None, None


def apply_masks(x, masks):
    """mask.py:58-67: x [B, N, D], masks = list of [B, K] index tensors -> [len(masks) * B, K, D] (autograd-aware)."""
    if x.dim() != 3:
        raise ValueError("x must be [B, N, D]")
    if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise ValueError("x must be fp32 / bf16 / fp16")
    idx, n_masks, K = _stack_masks(masks, x.shape[0])
    return _ApplyMasksFn.apply(x.contiguous(), idx.to(x.device), n_masks, K)


class _RepeatInterleaveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, B, repeat):
        n_groups = x.shape[0] // B
        slab = x[0].numel() * x.element_size()
        out = torch.empty((n_groups * repeat * B,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        L.repeat_interleave_batch(x, slab, B, n_groups, repeat, out)
        ctx.dims = (B, repeat, n_groups)
        return out

    @staticmethod
    def backward(ctx, dy):
        B, repeat, n_groups = ctx.dims
        # the sum over the `repeat` copies of each group (torch's cat backward adds them in copy order)
        g = dy.reshape((n_groups, repeat, B) + tuple(dy.shape[1:]))
        dx = g[:, 0].clone()
        for r in range(1, repeat):
            dx += g[:, r]
        return dx.reshape((n_groups * B,) + tuple(dy.shape[1:])), None, None


def repeat_interleave_batch(x, B, repeat):
    """tensors.py:65-71: [n*B, ...] -> [n*repeat*B, ...], each block of B rows repeated `repeat` times in place."""
    if len(x) % B != 0:
        raise ValueError("len(x) must be a multiple of B")
    if (x[0].numel() * x.element_size()) % 16 != 0:
        raise ValueError("rows must be a multiple of 16 bytes")
    return _RepeatInterleaveFn.apply(x.contiguous(), B, repeat)


def jepa_targets(h, masks_pred, n_enc_masks, eps=1e-5):
    """pretrain_jepa.py:384-392 (inside torch.no_grad()): layer_norm over the feature dim, keep the patches of every
    prediction mask, repeat each block for every context mask.  h [B, N, D] fp32 / bf16 -> fp32
    [len(masks_pred) * n_enc_masks * B, K, D]."""
    if h.dim() != 3 or h.shape[2] % 4 != 0 or h.shape[2] > 1024:
        raise ValueError("h must be [B, N, D] with D % 4 == 0 and D <= 1024")
    if h.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("h must be fp32 or bf16")
    B, N, D = h.shape
    idx, n_masks, K = _stack_masks(masks_pred, B)
    out = torch.empty((n_masks * n_enc_masks * B, K, D), dtype=torch.float32, device=h.device)
    status = torch.zeros(1, dtype=torch.int32, device=h.device) if _STRICT else None
    L.jepa_targets(h.detach().contiguous(), idx.to(h.device), B, N, D, n_masks, K, n_enc_masks, eps, out, status)
    _check_status(status, "jepa_targets")
    return out


class _SmoothL1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, h, beta):
        n = z.numel()
        dev = z.device
        partials = torch.empty(L.smooth_l1_slots(n), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        L.smooth_l1_fwd(z, h, n, beta, partials)
        L.loss_finalize(partials, n, None, loss)
        ctx.save_for_backward(z, h)
        ctx.beta = beta
        return loss

    @staticmethod
    def backward(ctx, g):
        z, h = ctx.saved_tensors
        dz = torch.empty_like(z)
        L.smooth_l1_bwd(z, h, z.numel(), ctx.beta, g.to(torch.float32).reshape(1).contiguous(), dz)
        return dz, None, None


def smooth_l1_loss(z, h, beta=1.0):
    """F.smooth_l1_loss(z, h) with mean reduction (pretrain_jepa.py:399-402): z fp32 / bf16 (the predictor output,
    receives the gradient), h fp32 (the no-grad target); the loss is fp32 as under autocast."""
    if z.shape != h.shape:
        raise ValueError("z and h must have the same shape")
    if z.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("z must be fp32 or bf16")
    if beta <= 0:
        raise ValueError("beta must be positive")
    return _SmoothL1Fn.apply(z.contiguous(), h.detach().to(torch.float32).contiguous(), float(beta))


class _EmaEntry(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("src", C.c_void_p), ("n", C.c_int64)]


_ema_tables = {}


@torch.no_grad()
def ema_update(encoder_params, target_params, m):
    """pretrain_jepa.py:426-432: for (q, k) in zip(encoder.parameters(), target_encoder.parameters()):
    k.mul_(m).add_((1. - m) * q) -- every pair in ONE launch.  The pointer table is cached per parameter set."""
    qs, ks = list(encoder_params), list(target_params)
    if len(qs) != len(ks) or not qs:
        raise ValueError("parameter lists must be non-empty and of equal length")
    key = tuple((k.data_ptr(), q.data_ptr(), k.numel()) for q, k in zip(qs, ks))
    hit = _ema_tables.get(key)
    if hit is None:
        for q, k in zip(qs, ks):
            if q.dtype != torch.float32 or k.dtype != torch.float32 or q.shape != k.shape:
                raise ValueError("EMA pairs must be fp32 tensors of equal shape")
            if not (q.is_contiguous() and k.is_contiguous() and q.is_cuda and k.is_cuda):
                raise ValueError("EMA pairs must be contiguous CUDA tensors")
        arr = (_EmaEntry * len(qs))()
        for i, (q, k) in enumerate(zip(qs, ks)):
            arr[i].dst, arr[i].src, arr[i].n = k.data_ptr(), q.data_ptr(), k.numel()
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
        hit = (raw.to(ks[0].device), len(qs), sum(k.numel() for k in ks))
        if len(_ema_tables) > 16:
            _ema_tables.clear()
        _ema_tables[key] = hit
    L.ema_update(hit[0], hit[1], float(m), hit[2])
