"""CPU restatement of the predictive (JEPA) path's non-ViT pieces -- TEST INFRASTRUCTURE ONLY (imported by tests/, never
by the product path).  Follows pretraining/predictive/mask.py:58-67 (apply_masks), tensors.py:65-71
(repeat_interleave_batch), pretrain_jepa.py:384-392 (target branch), :399-402 (smooth-L1) and :426-432 (EMA update).
Pinned against fixtures produced by the reference's own functions and mask collator
(tools/make_golden_jepa.py -> tests/golden/jepa_*.npz)."""
import torch


def apply_masks(x, masks):
    """mask.py:58-67: keep rows masks[i][b, :] of x[b] for every mask i; blocks concatenated along the batch."""
    B = x.shape[0]
    rows = torch.arange(B)[:, None]
    return torch.cat([x[rows, m.long()] for m in masks], dim=0)


def repeat_interleave_batch(x, B, repeat):
    """tensors.py:65-71: every block of B rows repeated `repeat` times, blocks kept in order."""
    n = len(x) // B
    return torch.cat([x[i * B:(i + 1) * B] for i in range(n) for _ in range(repeat)], dim=0)


def layer_norm_rows(h, eps=1e-5):
    """F.layer_norm(h, (D,)) without affine: biased variance over the last dim."""
    mu = h.mean(-1, keepdim=True)
    var = ((h - mu) ** 2).mean(-1, keepdim=True)
    return (h - mu) / torch.sqrt(var + eps)


def jepa_targets(h, masks_pred, n_enc_masks, eps=1e-5):
    """pretrain_jepa.py:384-392."""
    B = len(h)
    t = apply_masks(layer_norm_rows(h, eps), masks_pred)
    return repeat_interleave_batch(t, B, n_enc_masks)


def smooth_l1_loss(z, h, beta=1.0):
    """F.smooth_l1_loss(z, h), mean reduction (pretrain_jepa.py:400)."""
    d = z - h
    a = d.abs()
    return torch.where(a < beta, 0.5 * d * d / beta, a - 0.5 * beta).mean()


def ema_update(params_q, params_k, m):
    """pretrain_jepa.py:430-431 on fp32 tensors, with torch's roundings: the Python scalars m and (1. - m) are rounded
    to fp32, each product and the sum are rounded to fp32."""
    mf = torch.tensor(m, dtype=torch.float64).to(torch.float32)
    omf = torch.tensor(1.0 - m, dtype=torch.float64).to(torch.float32)
    return [(k * mf) + (q * omf) for q, k in zip(params_q, params_k)]
