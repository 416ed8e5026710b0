#!/usr/bin/env python
"""Generate tests/golden/* by running the REAL reference code in the build container:

  * transformers.VideoMAEForPreTraining (the third-party file that holds the hot path's arithmetic,
    transformers 5.5.0) -- loss, logits, labels-by-hook and every parameter gradient;
  * /root/reference/pretraining/generative/mask.py -- TubeMaskingGenerator / RandomMaskingGenerator.

Inputs are regenerated from seeds by oracle.videomae_oracle (torch CPU generators are deterministic), so only
outputs are stored.  Run from the repo root:  python tools/make_golden.py
This script needs /root/reference and is never run on the GPU box; the fixtures travel instead.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/pretraining/generative")

import transformers  # noqa: E402
from transformers.models.videomae import modeling_videomae as hf  # noqa: E402

import mask as refmask  # noqa: E402  (the reference's own file)
from oracle import videomae_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def hf_model(cfg: O.OracleConfig, params):
    c = transformers.VideoMAEConfig(
        image_size=cfg.image_size, patch_size=cfg.patch_size, num_channels=cfg.num_channels,
        num_frames=cfg.num_frames, tubelet_size=cfg.tubelet_size, hidden_size=cfg.hidden_size,
        num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
        intermediate_size=cfg.intermediate_size, initializer_range=0.02, use_mean_pooling=True,
        decoder_num_attention_heads=cfg.decoder_num_attention_heads, decoder_hidden_size=cfg.decoder_hidden_size,
        decoder_num_hidden_layers=cfg.decoder_num_hidden_layers,
        decoder_intermediate_size=cfg.decoder_intermediate_size, norm_pix_loss=True)
    m = transformers.VideoMAEForPreTraining(c)
    missing = m.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.train()


def run_hf(cfg, params, x, mask, bf16=False):
    m = hf_model(cfg, params)
    if bf16:  # the reference's own mixed-precision path (pretrain_videomae.py:306-308), on CPU autocast
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out = m(x, bool_masked_pos=mask)
    else:
        out = m(x, bool_masked_pos=mask)
    out.loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    return out.loss.detach(), out.logits.detach(), grads


def masks():
    d = {}
    for name, size, ratio in (("tube_8x14x14_r90", (8, 14, 14), 0.9), ("tube_1x14x14_r90", (1, 14, 14), 0.9),
                              ("tube_2x2x2_r50", (2, 2, 2), 0.5)):
        np.random.seed(0)
        g = refmask.TubeMaskingGenerator(size, ratio)
        d[name] = np.stack([g() for _ in range(4)]).astype(np.uint8)
    np.random.seed(0)
    g = refmask.RandomMaskingGenerator((8, 14, 14), 0.9)
    d["random_8x14x14_r90"] = np.stack([g() for _ in range(2)]).astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "masks.npz"), **d)
    print("masks:", {k: v.shape for k, v in d.items()})


def sinusoid():
    d = {}
    for n, dim in ((1568, 768), (1568, 384), (196, 192), (8, 64)):
        t = hf.get_sinusoid_encoding_table(n, dim)[0].numpy()
        rows = sorted(set([0, 1, 2, n // 2, n - 1]))
        d[f"n{n}_d{dim}_rows"] = np.array(rows)
        d[f"n{n}_d{dim}"] = t[rows]
    np.savez_compressed(os.path.join(OUT, "sinusoid.npz"), **d)


def tiny_step():
    """Full outputs for the tiny config (fits in git)."""
    res = {}
    for tag, perturb in (("init", False), ("perturbed", True)):
        cfg = O.make_config("tiny")
        params = O.init_params(cfg, seed=1, perturb=perturb)
        x = O.synthetic_clip(3, cfg, seed=2, image_like=perturb)
        np.random.seed(3)
        mask = O.batch_tube_masks(3, cfg.grid, 0.5)
        loss, logits, grads = run_hf(cfg, params, x, mask)
        res[f"{tag}.mask"] = mask.numpy()
        res[f"{tag}.loss"] = loss.numpy()
        res[f"{tag}.logits"] = logits.numpy()
        for k, g in grads.items():
            res[f"{tag}.grad.{k}"] = g.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "tiny_step.npz"), **res)
    print("tiny_step: loss init %.7f perturbed %.7f" % (res["init.loss"], res["perturbed.loss"]))


def summarised_step(name, batch, fname):
    """Full-size configs: loss, per-parameter gradient norms, strided logits samples."""
    out = {}
    for tag, perturb in (("init", False), ("perturbed", True)):
        cfg = O.make_config(name)
        params = O.init_params(cfg, seed=0, perturb=perturb)
        x = O.synthetic_clip(batch, cfg, seed=0, image_like=perturb)
        np.random.seed(0)
        mask = O.batch_tube_masks(batch, cfg.grid, 0.9)
        loss, logits, grads = run_hf(cfg, params, x, mask)
        bloss, blogits, bgrads = run_hf(cfg, params, x, mask, bf16=True)
        flat = logits.flatten()
        bsamp, fsamp = blogits.flatten()[::100003][:64].double(), flat[::100003][:64].double()
        bf16_dev = {k: abs(float(bgrads[k].double().norm()) - float(g.double().norm())) / max(float(g.double().norm()), 1e-30)
                    for k, g in grads.items()}
        bg = float(torch.sqrt(sum(g.double().pow(2).sum() for g in bgrads.values())))
        fg = float(torch.sqrt(sum(g.double().pow(2).sum() for g in grads.values())))
        out[tag] = {
            # how far the reference's OWN bf16-autocast path is from its fp32 path on these inputs
            "hf_bf16_loss_rel": abs(float(bloss) - float(loss)) / float(loss),
            "hf_bf16_logits_sample_rel_l2": float((bsamp - fsamp).norm() / fsamp.norm()),
            "hf_bf16_grad_norm_rel": bf16_dev,
            "hf_bf16_grad_global_norm_rel": abs(bg - fg) / fg,
            "hf_bf16_grad_global_rel_l2": float(torch.sqrt(sum((bgrads[k].double() - g.double()).pow(2).sum()
                                                               for k, g in grads.items()))) / fg,
            "loss": float(loss),
            "mask_row0_first_visible": [int(i) for i in np.nonzero(~mask[0].numpy())[0][:8]],
            "logits_sample_stride": 100003,
            "logits_sample": [float(v) for v in flat[::100003][:64]],
            "logits_l2": float(flat.double().norm()),
            "grad_norms": {k: float(g.double().norm()) for k, g in grads.items()},
            "grad_global_norm": float(torch.sqrt(sum(g.double().pow(2).sum() for g in grads.values()))),
        }
        print(name, tag, "loss", out[tag]["loss"], "gnorm", out[tag]["grad_global_norm"])
    out["meta"] = {"config": name, "batch": batch, "param_seed": 0, "clip_seed": 0, "np_mask_seed": 0,
                   "mask_ratio": 0.9, "transformers": transformers.__version__, "torch": torch.__version__,
                   "dtype": "float32 (CPU)"}
    with open(os.path.join(OUT, fname), "w") as f:
        json.dump(out, f, indent=1)


def target_samples():
    """HF's label tensor is internal; capture it by running HF with logits forced to zero:
    head weight/bias = 0 -> loss = mean(labels^2); and directly re-run HF:598-670 through a hooked MSELoss."""
    cfg = O.make_config("tiny")
    params = O.init_params(cfg, seed=1)
    x = O.synthetic_clip(2, cfg, seed=5, image_like=True)
    np.random.seed(6)
    mask = O.batch_tube_masks(2, cfg.grid, 0.5)
    m = hf_model(cfg, params)
    captured = {}
    orig = hf.MSELoss.forward

    def spy(self, inp, tgt):
        captured["labels"] = tgt.detach().clone()
        return orig(self, inp, tgt)

    hf.MSELoss.forward = spy
    try:
        m(x, bool_masked_pos=mask)
    finally:
        hf.MSELoss.forward = orig
    np.savez_compressed(os.path.join(OUT, "tiny_labels.npz"), mask=mask.numpy(), labels=captured["labels"].numpy())
    print("labels", captured["labels"].shape)


if __name__ == "__main__":
    torch.set_num_threads(8)
    masks()
    sinusoid()
    tiny_step()
    target_samples()
    summarised_step("small", 2, "small_step.json")
    if "--base" in sys.argv:
        summarised_step("base", 2, "base_step.json")
    if "--large" in sys.argv:   # BASELINE.json configs[4]'s model (ViT-L/16), one clip
        summarised_step("large", 1, "large_step.json")
