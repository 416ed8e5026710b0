"""SimCLR loss of the contrastive path on libbvc.so -- drop-in for `info_nce_loss(temperature, masks, feats)` of
pretraining/contrastive/pretrain_simclr.py:114-128 (the training driver builds `criterion = partial(info_nce_loss,
temperature, masks)`, :292) and for its mask helpers (:86-91, :284-291).

Same semantics as the reference, quirks included: cosine similarity with each norm clamped at
1e-8, divided by the temperature; boolean indexing flattens, so the loss is ONE global logsumexp over all negatives
minus the mean of the positives; the positive mask is the tri-diagonal |i - j| = 1.  The reference materialises an
n x n x D fp32 tensor for the similarity; here it is one tcgen05 GEMM over the row-normalised features (operands split
into bf16 hi + lo, so the logits keep ~16 mantissa bits) and one masked pass over the n x n scores, with a hand-written
backward (a second GEMM).  CUDA only: there is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L

EPS = 1e-8  # torch.nn.functional.cosine_similarity default, which the reference relies on


def get_special_matrix(n):
    """pretrain_simclr.py:86-91: 1 where |i - j| == 1 (vectorised; the reference builds it with a Python double loop)."""
    i = np.arange(n)
    return (np.abs(i[:, None] - i[None, :]) == 1).astype(np.int64)


def make_masks(mask_size, device):
    """pretrain_simclr.py:285-291 -> (pos_mask, neg_mask), bool [mask_size, mask_size] on `device`."""
    self_mask = torch.eye(mask_size, dtype=torch.bool, device=device)
    pos_mask = torch.tensor(get_special_matrix(mask_size), dtype=torch.bool, device=device)
    neg_mask = torch.ones_like(pos_mask)
    neg_mask[pos_mask | self_mask] = False
    return pos_mask, neg_mask


class _InfoNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, pos_u8, neg_u8, temperature):
        n, D = feats.shape
        dev = feats.device
        bf = torch.bfloat16
        a_split = torch.empty((n, 3 * D), dtype=bf, device=dev)
        b_split = torch.empty((n, 3 * D), dtype=bf, device=dev)
        bk_split = torch.empty((3 * n, D), dtype=bf, device=dev)
        inv_norm = torch.empty(n, dtype=torch.float32, device=dev)
        L.nce_normalize_split(feats, n, D, EPS, a_split, b_split, bk_split, inv_norm)
        S = torch.empty((n, n), dtype=torch.float32, device=dev)
        L.gemm(a_split, b_split, n, n, 3 * D, out_f32=S, alpha=1.0 / temperature)
        partials = torch.empty(L.nce_partial_slots(n), dtype=torch.float32, device=dev)
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        L.nce_loss(S, pos_u8, neg_u8, n, partials, out4)
        ctx.saved = (feats, pos_u8, neg_u8, S, out4, bk_split, inv_norm)
        ctx.temperature = temperature
        return out4[0].clone()

    @staticmethod
    def backward(ctx, g):
        feats, pos_u8, neg_u8, S, out4, bk_split, inv_norm = ctx.saved
        ctx.saved = None
        n, D = feats.shape
        dev = feats.device
        go = g.detach().to(torch.float32).reshape(1).contiguous()
        g_split = torch.empty((n, 3 * n), dtype=torch.bfloat16, device=dev)
        L.nce_grad(S, pos_u8, neg_u8, n, out4, go, g_split)
        dfhat = torch.empty((n, D), dtype=torch.float32, device=dev)
        L.gemm(g_split, bk_split, n, D, 3 * n, b_mn=True, ldb=D, out_f32=dfhat, alpha=1.0 / ctx.temperature)
        dfeats = torch.empty((n, D), dtype=torch.float32, device=dev)
        L.nce_normalize_bwd(dfhat, feats, inv_norm, n, D, EPS, dfeats)
        return dfeats.to(feats.dtype), None, None, None


def info_nce_loss(temperature, masks, feats, mode="train"):
    """Same call as the reference's info_nce_loss (pretrain_simclr.py:114): masks = (pos_mask, neg_mask) bool [n, n],
    feats [n, D] fp32 or bf16 on CUDA; returns the 0-dim fp32 loss with autograd history."""
    if feats.dim() != 2:
        raise ValueError("feats must be [n, D]")
    if not feats.is_cuda:
        raise L.BvcError("info_nce_loss (bvc-b200) runs on CUDA only; there is no CPU path")
    n, D = feats.shape
    pos_mask, neg_mask = masks
    if tuple(pos_mask.shape) != (n, n) or tuple(neg_mask.shape) != (n, n):
        raise ValueError(f"masks must be [{n}, {n}]")
    if n % 8 or D % 8:
        raise ValueError("info_nce_loss (bvc-b200): n and D must be multiples of 8")
    if feats.dtype not in (torch.float32, torch.bfloat16):
        feats = feats.float()
    f = feats if feats.stride(1) == 1 else feats.contiguous()
    pos_u8 = pos_mask.to(device=f.device).contiguous().view(torch.uint8) if pos_mask.dtype == torch.bool else \
        (pos_mask != 0).to(device=f.device).contiguous().view(torch.uint8)
    neg_u8 = neg_mask.to(device=f.device).contiguous().view(torch.uint8) if neg_mask.dtype == torch.bool else \
        (neg_mask != 0).to(device=f.device).contiguous().view(torch.uint8)
    with torch.cuda.device(f.device):
        return _InfoNCEFn.apply(f, pos_u8, neg_u8, float(temperature))
