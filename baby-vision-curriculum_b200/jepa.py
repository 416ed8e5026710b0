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
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("x must be fp32 or bf16 (the backward accumulates in x's dtype)")
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
    key = tuple((k.data_ptr(), q.data_ptr(), k.numel(), k.dtype, q.dtype) for q, k in zip(qs, ks))
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


# ------------------------------------------------------------------------------------------------ host side: masks
# The predictive path's masks are made on the host, in the DataLoader's collate function, and shifted to a temporal slot
# before they reach the device.  Mirrors with the reference's names, arguments and random-number
# consumption, so a seeded run produces the same index tensors (tests/test_jepa_oracle.py pins them on fixtures made by
# the reference's own classes).
def update_masks(masks, image_size, patch_size, num_frames, tubelet_size, isencoder=False):
    """pretraining/predictive/mask.py:21-38: per-frame patch indices -> spatio-temporal token indices.  Context masks
    stay in the first temporal slot, prediction masks move to the last one (T - 1).  In place, like the reference."""
    per_frame = (image_size // patch_size) ** 2
    slot = 0 if isencoder else num_frames // tubelet_size - 1
    for i, m in enumerate(masks):
        m += slot * per_frame
        masks[i] = m
    return masks


class MaskCollator(object):
    """pretraining/predictive/mask.py:70-219 (the multi-block collator of pretrain_jepa.py:226-235): per batch one
    prediction-block size and one context-block size from a generator seeded with a counter shared by the DataLoader
    workers; per sample `npred` prediction blocks and `nenc` context blocks at uniformly random corners (global torch
    RNG), the context blocks restricted to the complement of the sample's prediction blocks unless `allow_overlap`;
    every mask truncated to the shortest one of the batch.  Returns (collated batch, masks_enc, masks_pred) with
    masks_* = list of int64 [B, K] tensors of kept patch indices in ascending order."""

    def __init__(self, input_size=(224, 224), patch_size=16, enc_mask_scale=(0.2, 0.8), pred_mask_scale=(0.2, 0.8),
                 aspect_ratio=(0.3, 3.0), nenc=1, npred=2, min_keep=4, allow_overlap=False):
        from multiprocessing import Value
        if not isinstance(input_size, tuple):
            input_size = (input_size,) * 2
        self.patch_size = patch_size
        self.height, self.width = input_size[0] // patch_size, input_size[1] // patch_size
        self.enc_mask_scale, self.pred_mask_scale, self.aspect_ratio = enc_mask_scale, pred_mask_scale, aspect_ratio
        self.nenc, self.npred, self.min_keep, self.allow_overlap = nenc, npred, min_keep, allow_overlap
        self._itr_counter = Value("i", -1)   # shared across worker processes

    def step(self):
        with self._itr_counter.get_lock():
            self._itr_counter.value += 1
            return self._itr_counter.value

    def _block_size(self, generator, scale, aspect_ratio_scale):
        import math
        u = torch.rand(1, generator=generator).item()          # ONE draw serves scale and aspect ratio (mask.py:104-112)
        keep = int(self.height * self.width * (scale[0] + u * (scale[1] - scale[0])))
        ar = aspect_ratio_scale[0] + u * (aspect_ratio_scale[1] - aspect_ratio_scale[0])
        h = min(int(round(math.sqrt(keep * ar))), self.height - 1)
        w = min(int(round(math.sqrt(keep / ar))), self.width - 1)
        return h, w

    def _block(self, size, forbidden=None):
        """One block of `size` at a random corner -> (kept indices, its rectangle).  `forbidden`: rectangles the block
        must avoid; after every 20 failed draws the LAST rectangle of the list is dropped (mask.py:126-150)."""
        h, w = size
        tries, budget = 0, 20
        grid = torch.arange(self.height * self.width).view(self.height, self.width)
        while True:
            top = int(torch.randint(0, self.height - h, (1,)))
            left = int(torch.randint(0, self.width - w, (1,)))
            keep = torch.zeros((self.height, self.width), dtype=torch.bool)
            keep[top:top + h, left:left + w] = True
            if forbidden is not None:
                for (t, l, hh, ww) in forbidden[:max(len(forbidden) - tries, 0)]:
                    keep[t:t + hh, l:l + ww] = False
            idx = grid[keep]                                      # row-major = ascending, what nonzero() returns
            if len(idx) > self.min_keep:
                return idx, (top, left, h, w)
            budget -= 1
            if budget == 0:
                tries += 1
                budget = 20

    def __call__(self, batch):
        B = len(batch)
        collated = torch.utils.data.default_collate(batch)
        g = torch.Generator()
        g.manual_seed(self.step())
        p_size = self._block_size(g, self.pred_mask_scale, self.aspect_ratio)
        e_size = self._block_size(g, self.enc_mask_scale, (1., 1.))
        pred, enc = [], []
        for _ in range(B):
            blocks = [self._block(p_size) for _ in range(self.npred)]
            pred.append([b[0] for b in blocks])
            forbidden = None if self.allow_overlap else [b[1] for b in blocks]
            enc.append([self._block(e_size, forbidden)[0] for _ in range(self.nenc)])
        kp = min(min(len(m) for m in ms) for ms in pred)
        ke = min(min(len(m) for m in ms) for ms in enc)
        kp, ke = min(kp, self.height * self.width), min(ke, self.height * self.width)
        masks_pred = torch.utils.data.default_collate([[m[:kp] for m in ms] for ms in pred])
        masks_enc = torch.utils.data.default_collate([[m[:ke] for m in ms] for ms in enc])
        return collated, masks_enc, masks_pred
