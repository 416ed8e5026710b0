"""CPU restatement of the reference's SimCLR loss -- TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product
path): pretraining/contrastive/pretrain_simclr.py `info_nce_loss` (:114-128), `get_special_matrix` (:86-91) and the mask
construction of the training driver (:284-292).  Pinned against fixtures generated from the reference's own functions
(tools/make_golden_simclr.py -> tests/golden/simclr_*.npz).

Semantics as implemented by the reference: feats [n, D]; S = cos_sim(f_i, f_j) / T with every
norm clamped at eps = 1e-8 (torch.nn.functional.cosine_similarity); boolean-mask indexing flattens, so
    loss = logsumexp(S[neg_mask])  -  mean(S[pos_mask])
is ONE global log-sum-exp over all negatives, not a per-row NT-Xent.  pos_mask is the tri-diagonal |i - j| = 1,
neg_mask everything else off the diagonal.
"""
import numpy as np
import torch


def get_special_matrix(n):
    """pretrain_simclr.py:86-91 -- 1 where |i - j| == 1."""
    i = np.arange(n)
    return (np.abs(i[:, None] - i[None, :]) == 1).astype(np.int64)


def make_masks(n):
    """pretrain_simclr.py:286-291 -> (pos_mask, neg_mask) bool [n, n]."""
    self_mask = torch.eye(n, dtype=torch.bool)
    pos_mask = torch.tensor(get_special_matrix(n), dtype=torch.bool)
    neg_mask = torch.ones_like(pos_mask)
    neg_mask[pos_mask | self_mask] = False
    return pos_mask, neg_mask


def info_nce_loss(temperature, masks, feats, eps=1e-8):
    """pretrain_simclr.py:114-128 in fp64 (feats any float dtype; returns a 0-dim fp64 tensor with autograd history)."""
    f = feats.double()
    fn = f / f.norm(dim=1, keepdim=True).clamp_min(eps)   # ATen cosine_similarity: each operand / max(|x|, eps)
    s = (fn @ fn.t()) / temperature
    pos_mask, neg_mask = masks
    neg_part = torch.logsumexp(s[neg_mask], dim=-1)
    return neg_part - s[pos_mask].mean()


def loss_and_grad(temperature, masks, feats):
    f = feats.detach().clone().double().requires_grad_(True)
    loss = info_nce_loss(temperature, masks, f)
    loss.backward()
    return loss.detach(), f.grad.detach()
