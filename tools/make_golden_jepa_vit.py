#!/usr/bin/env python
"""Golden vectors for the predictive path's ViT block and for the un-masked VideoMAE encoder pass, from the REAL code
(run in the build container, where /root/reference exists):

  * /root/reference/pretraining/predictive/vision_transformer.py `Block` (:213-231, fused-qkv `Attention` :186-210,
    `MLP` :167-183) with norm_layer = partial(nn.LayerNorm, eps=1e-6) as vit_base() configures it -- output and every
    parameter / input gradient of one block in fp64, for seeded parameters and inputs (regenerated from the seed by
    oracle.jepa_oracle.vit_block_params, so only outputs are stored);
  * transformers.VideoMAEModel (HF:420-470) with bool_masked_pos=None on the tiny configuration -- the encoder pass of
    benchmarks/compute_embeddings_videomae.py:261 (all 1568 tokens at full size).

    python tools/make_golden_jepa_vit.py      ->  tests/golden/jepa_vit_block.npz, tests/golden/tiny_encode.npz
"""
import os
import sys
from functools import partial

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/pretraining/predictive")
import vision_transformer as RV  # noqa: E402  (the reference's own file)

from oracle import jepa_oracle as JO  # noqa: E402
from oracle import videomae_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

out = {}
for tag, dim, heads, B, N, qkv_bias in (("d128", 128, 2, 2, 40, True), ("d192_nobias", 192, 3, 1, 72, False)):
    params = JO.vit_block_params(dim, seed=7, qkv_bias=qkv_bias)
    blk = RV.Block(dim=dim, num_heads=heads, mlp_ratio=4.0, qkv_bias=qkv_bias, norm_layer=partial(nn.LayerNorm, eps=1e-6))
    blk.load_state_dict(params, strict=True)
    blk = blk.double()
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(B, N, dim, generator=g) * 1.5 + 0.2).double().requires_grad_(True)
    w = torch.randn(B, N, dim, generator=g).double()
    y = blk(x)
    (y * w).sum().backward()
    out[f"{tag}.y"] = y.detach().numpy().astype(np.float32)
    out[f"{tag}.dx"] = x.grad.numpy().astype(np.float32)
    for k, p in blk.named_parameters():
        if tag == "d128":   # full gradients (fp32 copies of the fp64 run) for one case, norms for the other
            out[f"{tag}.grad.{k}"] = p.grad.numpy().astype(np.float32)
        out[f"{tag}.gradnorm.{k}"] = np.array(float(p.grad.norm()))
    out[f"{tag}.meta"] = np.array([dim, heads, B, N, int(qkv_bias)])
np.savez_compressed(os.path.join(OUT, "jepa_vit_block.npz"), **out)
print("jepa_vit_block.npz", {k: v.shape for k, v in out.items() if k.endswith((".y", ".meta"))})

import transformers  # noqa: E402

cfg = O.make_config("tiny")
params = O.init_params(cfg, seed=1, perturb=True)
c = transformers.VideoMAEConfig(
    image_size=cfg.image_size, patch_size=cfg.patch_size, num_channels=cfg.num_channels, num_frames=cfg.num_frames,
    tubelet_size=cfg.tubelet_size, hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_hidden_layers,
    num_attention_heads=cfg.num_attention_heads, intermediate_size=cfg.intermediate_size, use_mean_pooling=True,
    decoder_num_attention_heads=cfg.decoder_num_attention_heads, decoder_hidden_size=cfg.decoder_hidden_size,
    decoder_num_hidden_layers=cfg.decoder_num_hidden_layers, decoder_intermediate_size=cfg.decoder_intermediate_size,
    norm_pix_loss=True)
m = transformers.VideoMAEForPreTraining(c)
m.load_state_dict(params, strict=True)
x = O.synthetic_clip(2, cfg, seed=9, image_like=True)
with torch.no_grad():
    h = m.videomae(x).last_hidden_state
np.savez_compressed(os.path.join(OUT, "tiny_encode.npz"), last_hidden_state=h.numpy())
print("tiny_encode.npz", tuple(h.shape))
