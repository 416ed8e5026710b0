#!/usr/bin/env python
"""Golden vectors for the JEPA pieces from the reference's OWN code (run in the build container, where /root/reference
exists): masks from predictive/mask.py MaskCollator + update_masks with the training script's settings
(pretrain_jepa.py:186-194: 1 context mask, 4 prediction masks, 224 / 16 / 16 frames / tubelet 2 -> N = 1568), then
apply_masks / repeat_interleave_batch (mask.py:58-67, tensors.py:65-71) composed as pretrain_jepa.py:384-392, the
smooth-L1 loss of :400 with its gradient, and the momentum update of :430-431."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, "/root/reference/pretraining/predictive")
import mask as RM  # noqa: E402
import tensors as RT  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_masks(B, seed):
    torch.manual_seed(seed)
    coll = RM.MaskCollator(input_size=(224, 224), patch_size=16, pred_mask_scale=(0.15, 0.2), enc_mask_scale=(0.85, 1.0),
                           aspect_ratio=(0.75, 1.5), nenc=1, npred=4, allow_overlap=False, min_keep=10)
    _, m_enc, m_pred = coll([torch.zeros(1) for _ in range(B)])
    m_enc = RM.update_masks(m_enc, 224, 16, 16, 2, isencoder=True)
    m_pred = RM.update_masks(m_pred, 224, 16, 16, 2, isencoder=False)
    return m_enc, m_pred


for tag, B, D, seed in (("tiny", 2, 16, 5), ("vitb", 4, 768, 6)):
    m_enc, m_pred = make_masks(B, seed)
    N = 1568
    rng = np.random.default_rng(200 + D)
    h = torch.from_numpy(rng.standard_normal((B, N, D)).astype(np.float32) * 1.7 + 0.3)
    hn = F.layer_norm(h, (D,))
    t = RT.repeat_interleave_batch(RM.apply_masks(hn, m_pred), B, repeat=len(m_enc))
    ctx = RM.apply_masks(h, m_enc)                                  # the context encoder's gather (vision_transformer.py)
    z = (t + torch.from_numpy(rng.standard_normal(tuple(t.shape)).astype(np.float32)) * 0.8).requires_grad_(True)
    loss = F.smooth_l1_loss(z, t)
    (loss * 3.0).backward()
    # gather backward through the reference function
    hx = h.clone().requires_grad_(True)
    both = RM.apply_masks(hx, list(m_pred))   # the 4 prediction blocks overlap: rows gathered more than once
    w = torch.from_numpy(rng.standard_normal(tuple(both.shape)).astype(np.float32))
    (both * w).sum().backward()
    # momentum update
    q = [torch.from_numpy(rng.standard_normal(s).astype(np.float32)) for s in ((D, 7), (13,), (5, D))]
    k = [torch.from_numpy(rng.standard_normal(s).astype(np.float32)) for s in ((D, 7), (13,), (5, D))]
    m = 0.996 + 3 * (1.0 - 0.996) / 1000
    k_new = [kk.clone() for kk in k]
    with torch.no_grad():
        for pq, pk in zip(q, k_new):
            pk.data.mul_(m).add_((1. - m) * pq.detach().data)
    out = {"B": B, "D": D, "N": N, "seed": 200 + D, "momentum": np.float64(m),
           "masks_enc": torch.stack(list(m_enc)).numpy(), "masks_pred": torch.stack(list(m_pred)).numpy(),
           "loss": loss.detach().double().numpy(), "targets_checksum": np.float64(t.double().sum()),
           "targets_abs_checksum": np.float64(t.double().abs().sum()), "ctx_checksum": np.float64(ctx.double().sum()),
           "dz_checksum": np.float64(z.grad.double().abs().sum()), "dh_checksum": np.float64(hx.grad.double().abs().sum()),
           "targets_head": t[:, :2, :8].numpy(), "ctx_head": ctx[:, :2, :8].numpy(), "dz_head": z.grad[:, :2, :8].numpy(),
           "dh_rows": hx.grad[:, 1372:1380, :8].numpy()}
    for i in range(3):
        out[f"ema_k{i}"] = k_new[i].numpy()
    if tag == "tiny":
        out["targets"] = t.numpy()
        out["ctx"] = ctx.numpy()
        out["dz"] = z.grad.numpy()
    np.savez_compressed(os.path.join(OUT, f"jepa_{tag}.npz"), **out)
    print(tag, "K_enc", m_enc[0].shape, "K_pred", m_pred[0].shape, "loss", float(loss), "targets", tuple(t.shape))
