#!/usr/bin/env python
"""Golden vectors for the SimCLR loss from the reference's OWN functions (run in the build container, where
/root/reference exists): pretrain_simclr.info_nce_loss / get_special_matrix, fp32 and fp64, at a tiny size (full tensors),
the reference's real size (batch 32 -> n = 64, D = 512) and BASELINE.json config 3's size (batch 512 -> n = 1024)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/pretraining/contrastive")
import pretrain_simclr as R  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def masks(n):
    self_mask = torch.eye(n, dtype=torch.bool)
    pos = torch.tensor(R.get_special_matrix(n), dtype=torch.bool)
    neg = torch.ones_like(pos)
    neg[pos | self_mask] = False
    return pos, neg


for tag, n, D, scale in (("tiny", 16, 32, 1.0), ("ref", 64, 512, 1.0), ("cfg3", 1024, 512, 3.0)):
    feats = torch.from_numpy(np.random.default_rng(100 + n).standard_normal((n, D)).astype(np.float32)) * scale
    # correlated neighbours, like two views of the same clip
    feats[1::2] = 0.7 * feats[0::2] + 0.3 * feats[1::2]
    out = {"temperature": np.float64(0.1), "special": R.get_special_matrix(min(n, 16)), "n": n, "D": D, "scale": scale,
           "feats_checksum": np.float64(feats.double().sum())}
    if n <= 64:
        out["feats"] = feats.numpy()  # the large case is regenerated from the numpy seed (tests/helpers.simclr_feats)
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        f = feats.to(dt).detach().clone().requires_grad_(True)
        loss = R.info_nce_loss(0.1, masks(n), f)
        loss.backward()
        out[f"loss_{name}"] = loss.detach().double().numpy()
        gr = f.grad.detach().double().numpy()
        out[f"grad_norm_{name}"] = np.linalg.norm(gr)
        out[f"grad_{name}"] = gr if n <= 64 else gr[:8]
    np.savez_compressed(os.path.join(OUT, f"simclr_{tag}.npz"), **out)
    print(tag, float(out["loss_f64"]), float(out["grad_norm_f64"]))
