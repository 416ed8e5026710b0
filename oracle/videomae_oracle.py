"""CPU oracle for the VideoMAE pretraining step -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product path (the package next to this directory) never does; it fails loudly when libbvc.so
is missing.

What this restates (HF-free; plain torch CPU tensor arithmetic in fp32 or fp64, numpy for the integer
parts).  "HF:" = transformers 5.5.0 transformers/models/videomae/modeling_videomae.py, the third-party file
that holds the arithmetic of the reference's hot path
(/root/reference/pretraining/generative/pretrain_videomae.py:292-304, call at :301):

    tube_mask / random_mask       pretraining/generative/mask.py:3-24, :26-46
    sinusoid_table                HF:78-91
    mask_to_index                 HF:121-122 and HF:587-588 (boolean indexing == ascending index lists)
    patchify_embed_order          HF:157-177 (Conv3d k=s=(2,16,16) == GEMM over K-order (c,t,ph,pw))
    norm_pix_target               HF:598-643, HF:669-670
    block_forward                 HF:348-366, HF:236-266, HF:181-206, HF:281-284, HF:314-317, HF:327-331
    forward_loss                  HF:540-680 (VideoMAEForPreTraining.forward), HF:501-512 (decoder)
    loss_allreduce                pretraining/generative/ddputils.py:53-68

Pinned: tests/golden/*.npz were produced by tools/make_golden.py, which runs the *real* HF model and the
reference's own mask.py in the build container; tests/test_oracle.py checks this restatement against them
(the reference itself ships no tests or golden vectors -- SURVEY.md section 8c).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # transformers/utils/constants.py:1
IMAGENET_STD = (0.229, 0.224, 0.225)  # transformers/utils/constants.py:2


@dataclass
class OracleConfig:
    """The fields of transformers.VideoMAEConfig the path reads (pretrain_videomae.py:51-57)."""

    image_size: int = 224
    patch_size: int = 16
    num_channels: int = 3
    num_frames: int = 16
    tubelet_size: int = 2
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    decoder_num_attention_heads: int = 6
    decoder_hidden_size: int = 384
    decoder_num_hidden_layers: int = 4
    decoder_intermediate_size: int = 1536
    layer_norm_eps: float = 1e-12
    norm_pix_loss: bool = True

    @property
    def grid(self):
        g = self.image_size // self.patch_size
        return (self.num_frames // self.tubelet_size, g, g)

    @property
    def seq_len(self):
        t, h, w = self.grid
        return t * h * w

    @property
    def patch_dim(self):
        return self.num_channels * self.tubelet_size * self.patch_size * self.patch_size


CONFIGS = {
    # tiny: for golden fixtures that fit in git
    "tiny": dict(image_size=32, num_frames=4, hidden_size=64, num_hidden_layers=2, num_attention_heads=1,
                 intermediate_size=128, decoder_num_attention_heads=1, decoder_hidden_size=64,
                 decoder_num_hidden_layers=1, decoder_intermediate_size=128),
    # BASELINE.json configs[0]
    "small": dict(hidden_size=384, num_hidden_layers=12, num_attention_heads=6, intermediate_size=1536,
                  decoder_num_attention_heads=3, decoder_hidden_size=192, decoder_num_hidden_layers=4,
                  decoder_intermediate_size=768),
    # BASELINE.json configs[1]; pretrain_videomae.py:50-57
    "base": dict(),
    # BASELINE.json configs[4]
    "large": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                  decoder_num_attention_heads=8, decoder_hidden_size=512, decoder_num_hidden_layers=4,
                  decoder_intermediate_size=2048),
}


def make_config(name: str, **over) -> OracleConfig:
    kw = dict(CONFIGS[name])
    kw.update(over)
    return OracleConfig(**kw)


# ---------------------------------------------------------------------------------------------------
# masks (integer semantics)
# ---------------------------------------------------------------------------------------------------
def tube_mask(input_size, mask_ratio, rng=np.random):
    """mask.py:3-24. One per-frame {0,1} vector, shuffled, tiled over the temporal slots. 1 = masked.
    `rng` defaults to numpy's global RNG, which is what the reference uses (and never seeds)."""
    frames, height, width = input_size
    per_frame = height * width
    n_mask = int(mask_ratio * per_frame)  # mask.py:8
    m = np.hstack([np.zeros(per_frame - n_mask), np.ones(n_mask)])
    rng.shuffle(m)
    return np.tile(m, (frames, 1)).flatten()


def random_mask(input_size, mask_ratio, rng=np.random):
    """mask.py:26-46."""
    frames, height, width = input_size
    total = frames * height * width
    n_mask = int(mask_ratio * total)
    m = np.hstack([np.zeros(total - n_mask), np.ones(n_mask)])
    rng.shuffle(m)
    return m


def batch_tube_masks(batch, input_size, mask_ratio, rng=np.random):
    """pretrain_videomae.py:294-297: float64 zeros, one generator call per sample, .bool()."""
    n = input_size[0] * input_size[1] * input_size[2]
    out = np.zeros((batch, n))
    for i in range(batch):
        out[i, :] = tube_mask(input_size, mask_ratio, rng)
    return torch.from_numpy(out).bool()


def mask_to_index(mask: torch.Tensor):
    """x[~mask] / x[mask] with a per-row reshape (HF:121-122, 587-588) == per-row ascending index lists.
    Raises ValueError when rows disagree on the count (HF's reshape fails the same way)."""
    m = mask.cpu().numpy().astype(bool)
    nv = (~m).sum(1)
    if not (nv == nv[0]).all():
        raise ValueError("every row of bool_masked_pos must mask the same number of tokens")
    vis = np.stack([np.nonzero(~r)[0] for r in m]).astype(np.int32)
    msk = np.stack([np.nonzero(r)[0] for r in m]).astype(np.int32)
    return torch.from_numpy(vis), torch.from_numpy(msk)


# ---------------------------------------------------------------------------------------------------
# position table, patch orders, target
# ---------------------------------------------------------------------------------------------------
def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """HF:78-91 (float64 arithmetic through numpy, stored fp32). Returns [n_position, d_hid]."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    # np.power(10000, 2*(j//2)/d_hid): python-int / int division -> float64, as in HF:83
    ang = pos / np.power(10000, 2 * (j // 2) / d_hid)[None, :]
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(ang).to(torch.float32)


def patchify_embed_order(x: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    """[B,T,C,H,W] -> [B,N,K] with token n = (tg*Hg+hg)*Wg+wg and k = ((c*ts+t)*ps+ph)*ps+pw: the im2col
    of Conv3d(k=s=(ts,ps,ps)) after the permute at HF:175 (weight [D,C,ts,ps,ps] flattened)."""
    B, T, C, H, W = x.shape
    ts, ps = cfg.tubelet_size, cfg.patch_size
    x = x.view(B, T // ts, ts, C, H // ps, ps, W // ps, ps)
    x = x.permute(0, 1, 4, 6, 3, 2, 5, 7)  # B,tg,hg,wg,c,t,ph,pw
    return x.reshape(B, (T // ts) * (H // ps) * (W // ps), C * ts * ps * ps)


def patchify_target_order(x: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    """[B,T,C,H,W] -> [B,N,ts*ps*ps,C] (HF:615-633: view + permute(0,1,4,6,2,5,7,3))."""
    B, T, C, H, W = x.shape
    ts, ps = cfg.tubelet_size, cfg.patch_size
    x = x.view(B, T // ts, ts, C, H // ps, ps, W // ps, ps)
    x = x.permute(0, 1, 4, 6, 2, 5, 7, 3)
    return x.reshape(B, (T // ts) * (H // ps) * (W // ps), ts * ps * ps, C)


def norm_pix_target(x: torch.Tensor, msk_idx: torch.Tensor, cfg: OracleConfig) -> torch.Tensor:
    """HF:598-670: un-normalise with the ImageNet constants, per-(token,channel) mean / unbiased var over the
    ts*ps*ps pixels, normalise with eps 1e-6 added to the std, gather masked rows (ascending).
    Returns [B, Nm, ts*ps*ps*C] in x.dtype, feature order (t,ph,pw,c)."""
    if cfg.num_channels == 3:
        mean = torch.tensor(IMAGENET_MEAN, dtype=x.dtype)[None, None, :, None, None]
        std = torch.tensor(IMAGENET_STD, dtype=x.dtype)[None, None, :, None, None]
        frames = x * std + mean
    else:
        frames = x
    p = patchify_target_order(frames, cfg)
    if cfg.norm_pix_loss:
        p = (p - p.mean(dim=-2, keepdim=True)) / (p.var(dim=-2, unbiased=True, keepdim=True).sqrt() + 1e-6)
    p = p.reshape(p.shape[0], p.shape[1], -1)
    idx = msk_idx.long()[:, :, None].expand(-1, -1, p.shape[-1])
    return torch.gather(p, 1, idx)


# ---------------------------------------------------------------------------------------------------
# transformer block
# ---------------------------------------------------------------------------------------------------
def block_forward(h, p, prefix, n_heads, eps):
    """One pre-LN block (HF:348-366). `p` maps HF state-dict names to tensors."""
    d = h.shape[-1]
    dh = d // n_heads
    u = F.layer_norm(h, (d,), p[prefix + "layernorm_before.weight"], p[prefix + "layernorm_before.bias"], eps)
    a = prefix + "attention.attention."
    q = F.linear(u, p[a + "query.weight"], p[a + "q_bias"])
    k = F.linear(u, p[a + "key.weight"])  # k bias is a fresh zeros tensor, HF:239
    v = F.linear(u, p[a + "value.weight"], p[a + "v_bias"])
    B, S, _ = h.shape

    def heads(t):
        return t.view(B, S, n_heads, dh).transpose(1, 2)

    q, k, v = heads(q), heads(k), heads(v)
    s = (q @ k.transpose(-1, -2)) * (dh ** -0.5)  # HF:181-206 (eager) == SDPA math
    ctx = torch.softmax(s, dim=-1) @ v
    ctx = ctx.transpose(1, 2).reshape(B, S, d)
    o = prefix + "attention.output.dense."
    h = h + F.linear(ctx, p[o + "weight"], p[o + "bias"])  # HF:281-284 + residual HF:357
    u = F.layer_norm(h, (d,), p[prefix + "layernorm_after.weight"], p[prefix + "layernorm_after.bias"], eps)
    i = prefix + "intermediate.dense."
    f = F.gelu(F.linear(u, p[i + "weight"], p[i + "bias"]))  # exact erf GELU, HF:314-317
    o = prefix + "output.dense."
    return h + F.linear(f, p[o + "weight"], p[o + "bias"])  # HF:327-331


def forward_loss(params, pixel_values, bool_masked_pos, cfg: OracleConfig, return_parts=False):
    """VideoMAEForPreTraining.forward (HF:540-680) in the dtype of `params` (fp32 or fp64), gather-first:
    only visible patches are embedded (== embed-all-then-gather, HF:111-122).  Returns (loss, logits)
    [+ dict of intermediates].  Autograd-capable: pass params with requires_grad to get gradients."""
    p = params
    dt = p["mask_token"].dtype
    x = pixel_values.to(dt)
    B, T, C, H, W = x.shape
    if C != cfg.num_channels:
        raise ValueError("channel dimension of pixel_values does not match the configuration")  # HF:166-169
    if H != cfg.image_size or W != cfg.image_size:
        raise ValueError("input image size does not match the model")  # HF:170-173
    if bool_masked_pos is None:
        raise ValueError("One must provided a boolean mask ")  # HF:582-583
    vis_idx, msk_idx = mask_to_index(bool_masked_pos)
    Nv, Nm = vis_idx.shape[1], msk_idx.shape[1]
    D, Dd = cfg.hidden_size, cfg.decoder_hidden_size

    patches = patchify_embed_order(x, cfg)
    pv = torch.gather(patches, 1, vis_idx.long()[:, :, None].expand(-1, -1, patches.shape[-1]))
    w = p["videomae.embeddings.patch_embeddings.projection.weight"].reshape(D, -1)
    b = p["videomae.embeddings.patch_embeddings.projection.bias"]
    pos = sinusoid_table(cfg.seq_len, D).to(dt)
    h = F.linear(pv, w, b) + pos[vis_idx.long()]  # HF:111-122
    parts = {"patches_vis": pv, "embed": h}
    for i in range(cfg.num_hidden_layers):
        h = block_forward(h, p, f"videomae.encoder.layer.{i}.", cfg.num_attention_heads, cfg.layer_norm_eps)
    parts["encoder_out"] = h
    # use_mean_pooling=True -> no final encoder LayerNorm (HF:415-418, 472-473)
    z = F.linear(h, p["encoder_to_decoder.weight"])  # HF:576
    pos_d = sinusoid_table(cfg.seq_len, Dd).to(dt)
    xf = torch.cat([z + pos_d[vis_idx.long()], p["mask_token"] + pos_d[msk_idx.long()]], dim=1)  # HF:585-591
    parts["decoder_in"] = xf
    for j in range(cfg.decoder_num_hidden_layers):
        xf = block_forward(xf, p, f"decoder.decoder_layers.{j}.", cfg.decoder_num_attention_heads,
                           cfg.layer_norm_eps)
    xf = xf[:, -Nm:]  # HF:506
    xf = F.layer_norm(xf, (Dd,), p["decoder.norm.weight"], p["decoder.norm.bias"], 1e-5)  # nn.LayerNorm default
    logits = F.linear(xf, p["decoder.head.weight"], p["decoder.head.bias"])  # HF:510
    with torch.no_grad():
        labels = norm_pix_target(x, msk_idx, cfg)
    loss = F.mse_loss(logits, labels)  # HF:672-673
    if return_parts:
        parts.update(labels=labels, vis_idx=vis_idx, msk_idx=msk_idx)
        return loss, logits, parts
    return loss, logits


def loss_allreduce(local_losses):
    """ddputils.py:56-64: x / world_size then all_reduce(SUM) == mean over ranks; backward is identity."""
    w = len(local_losses)
    return sum(l / w for l in local_losses)


# ---------------------------------------------------------------------------------------------------
# parameter helpers
# ---------------------------------------------------------------------------------------------------
def param_shapes(cfg: OracleConfig):
    """HF state-dict names and shapes (probed against HF)."""
    D, Dd = cfg.hidden_size, cfg.decoder_hidden_size
    out = {"mask_token": (1, 1, Dd)}
    pe = "videomae.embeddings.patch_embeddings.projection."
    out[pe + "weight"] = (D, cfg.num_channels, cfg.tubelet_size, cfg.patch_size, cfg.patch_size)
    out[pe + "bias"] = (D,)

    def block(prefix, d, ff):
        a = prefix + "attention.attention."
        out[a + "q_bias"] = (d,)
        out[a + "v_bias"] = (d,)
        for n in ("query", "key", "value"):
            out[a + n + ".weight"] = (d, d)
        out[prefix + "attention.output.dense.weight"] = (d, d)
        out[prefix + "attention.output.dense.bias"] = (d,)
        out[prefix + "intermediate.dense.weight"] = (ff, d)
        out[prefix + "intermediate.dense.bias"] = (ff,)
        out[prefix + "output.dense.weight"] = (d, ff)
        out[prefix + "output.dense.bias"] = (d,)
        for n in ("layernorm_before", "layernorm_after"):
            out[prefix + n + ".weight"] = (d,)
            out[prefix + n + ".bias"] = (d,)

    for i in range(cfg.num_hidden_layers):
        block(f"videomae.encoder.layer.{i}.", D, cfg.intermediate_size)
    out["encoder_to_decoder.weight"] = (Dd, D)
    for j in range(cfg.decoder_num_hidden_layers):
        block(f"decoder.decoder_layers.{j}.", Dd, cfg.decoder_intermediate_size)
    out["decoder.norm.weight"] = (Dd,)
    out["decoder.norm.bias"] = (Dd,)
    out["decoder.head.weight"] = (cfg.patch_dim, Dd)
    out["decoder.head.bias"] = (cfg.patch_dim,)
    return out


def init_params(cfg: OracleConfig, seed=0, perturb=False, dtype=torch.float32):
    """HF-style init (N(0,0.02) weights, zero biases, LN 1/0, zero mask_token/q_bias/v_bias;
    modeling_utils.py:2285-2325) from a CPU generator -- NOT bit-identical to HF's own init order; parity tests
    always copy one state-dict into both sides.  perturb=True gives a 'trained-like' state (weights x4,
    non-zero biases / LN affine / mask_token) so that the network output actually matters for the loss."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        is_ln = "layernorm" in name or name.startswith("decoder.norm")
        if name.endswith("weight") and not is_ln:
            t = torch.randn(shape, generator=g) * (0.08 if perturb else 0.02)
        elif name.endswith("weight"):
            t = torch.ones(shape) + (0.2 * torch.randn(shape, generator=g) if perturb else 0)
        else:
            t = 0.1 * torch.randn(shape, generator=g) if perturb else torch.zeros(shape)
        out[name] = t.to(dtype)
    return out


def synthetic_clip(batch, cfg: OracleConfig, seed=0, image_like=False):
    """SURVEY.md section 8d synthetic inputs: randn, or a uint8-image-like variant through the reference's
    Normalize(0.5, 0.25) (homeview.py:218-231)."""
    g = torch.Generator().manual_seed(seed)
    shape = (batch, cfg.num_frames, cfg.num_channels, cfg.image_size, cfg.image_size)
    if image_like:
        u8 = torch.randint(0, 256, shape, generator=g)
        return ((u8.float() / 255.0) - 0.5) / 0.25
    return torch.randn(shape, generator=g)


def grads_of(params, pixel_values, mask, cfg, grad_scale=1.0):
    """loss, logits and d(loss*grad_scale)/d(param) for every parameter."""
    ps = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    loss, logits = forward_loss(ps, pixel_values, mask, cfg)
    (loss * grad_scale).backward()
    return loss.detach(), logits.detach(), {k: v.grad for k, v in ps.items()}


def encode_unmasked(params, pixel_values, cfg: OracleConfig):
    """HF VideoMAEModel.forward with bool_masked_pos=None (HF:420-470: embeddings of ALL tokens + sinusoid table, the
    encoder blocks, no final LayerNorm under use_mean_pooling=True) -- the un-masked encoder pass of
    benchmarks/compute_embeddings_videomae.py:261.  -> last_hidden_state [B, N, hidden]."""
    p = params
    patches = patchify_embed_order(pixel_values, cfg).to(p["videomae.embeddings.patch_embeddings.projection.weight"].dtype)
    w = p["videomae.embeddings.patch_embeddings.projection.weight"]
    h = F.linear(patches, w.reshape(w.shape[0], -1), p["videomae.embeddings.patch_embeddings.projection.bias"])
    h = h + sinusoid_table(cfg.seq_len, cfg.hidden_size).to(h.dtype)[None]
    for i in range(cfg.num_hidden_layers):
        h = block_forward(h, p, f"videomae.encoder.layer.{i}.", cfg.num_attention_heads, cfg.layer_norm_eps)
    return h
