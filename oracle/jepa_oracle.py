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


def vit_block(x, p, num_heads, eps=1e-6):
    """vision_transformer.py:186-231 -- `Block.forward`: y = x + proj(softmax(q k^T * scale) v) with ONE fused qkv Linear
    reshaped [B, N, 3, heads, head_dim] (:200), then y + fc2(gelu(fc1(norm2(y)))) (:167-183, exact-erf nn.GELU).
    `p` maps the block's state-dict names (norm1.weight, attn.qkv.weight, ..., mlp.fc2.bias) to tensors; attn.qkv.bias
    may be absent (qkv_bias=False).  Computes in the dtype of x / p."""
    import torch.nn.functional as F
    B, N, C = x.shape
    hd = C // num_heads
    u = F.layer_norm(x, (C,), p["norm1.weight"], p["norm1.bias"], eps)
    qkv = F.linear(u, p["attn.qkv.weight"], p.get("attn.qkv.bias")).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(y, p["attn.proj.weight"], p["attn.proj.bias"])
    u = F.layer_norm(x, (C,), p["norm2.weight"], p["norm2.bias"], eps)
    return x + F.linear(F.gelu(F.linear(u, p["mlp.fc1.weight"], p["mlp.fc1.bias"])), p["mlp.fc2.weight"], p["mlp.fc2.bias"])


def vit_block_params(dim, mlp_ratio=4.0, seed=0, qkv_bias=True):
    """Seeded 'trained-like' parameters of one block (state-dict names of vision_transformer.Block)."""
    g = torch.Generator().manual_seed(seed)
    ff = int(dim * mlp_ratio)
    shapes = {"norm1.weight": (dim,), "norm1.bias": (dim,), "attn.qkv.weight": (3 * dim, dim), "attn.qkv.bias": (3 * dim,),
              "attn.proj.weight": (dim, dim), "attn.proj.bias": (dim,), "norm2.weight": (dim,), "norm2.bias": (dim,),
              "mlp.fc1.weight": (ff, dim), "mlp.fc1.bias": (ff,), "mlp.fc2.weight": (dim, ff), "mlp.fc2.bias": (dim,)}
    out = {}
    for k, s in shapes.items():
        if k == "attn.qkv.bias" and not qkv_bias:
            continue
        if k.startswith("norm") and k.endswith("weight"):
            out[k] = 1.0 + 0.2 * torch.randn(s, generator=g)
        elif k.endswith("weight"):
            out[k] = 0.08 * torch.randn(s, generator=g)
        else:
            out[k] = 0.1 * torch.randn(s, generator=g)
    return out
