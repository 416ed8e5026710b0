"""CPU: the JEPA oracle (oracle/jepa_oracle.py) against fixtures made by the reference's own functions
(tools/make_golden_jepa.py: predictive/mask.py MaskCollator, update_masks, apply_masks; tensors.py
repeat_interleave_batch; pretrain_jepa.py:384-402, :426-432)."""
import numpy as np
import pytest
import torch

from oracle import jepa_oracle as J
from tests.helpers import jepa_case


@pytest.mark.parametrize("tag", ["tiny", "vitb"])
def test_jepa_oracle_matches_reference_fixtures(tag):
    g, h, m_enc, m_pred, noise, w, q, k = jepa_case(tag)
    B = h.shape[0]
    t = J.jepa_targets(h, m_pred, len(m_enc))
    ctx = J.apply_masks(h, m_enc)
    assert torch.equal(ctx[:, :2, :8], torch.from_numpy(g["ctx_head"]))          # index work: bit-exact
    assert float(ctx.double().sum()) == float(g["ctx_checksum"])
    np.testing.assert_allclose(t[:, :2, :8].numpy(), g["targets_head"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(float(t.double().abs().sum()), float(g["targets_abs_checksum"]), rtol=1e-6)
    if "targets" in g:
        np.testing.assert_allclose(t.numpy(), g["targets"], rtol=0, atol=2e-6)
        assert torch.equal(ctx, torch.from_numpy(g["ctx"]))
    # smooth-L1 and its gradient (upstream gradient 3.0 as in the generator)
    t_ref = torch.from_numpy(g["targets"]) if "targets" in g else t
    z = (t_ref + noise).requires_grad_(True)
    loss = J.smooth_l1_loss(z, t_ref)
    (loss * 3.0).backward()
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=2e-6)
    np.testing.assert_allclose(z.grad[:, :2, :8].numpy(), g["dz_head"], rtol=1e-5, atol=1e-9)
    # gather backward: rows hit by several prediction blocks accumulate
    hx = h.clone().requires_grad_(True)
    (J.apply_masks(hx, m_pred) * w).sum().backward()
    np.testing.assert_allclose(hx.grad[:, 1372:1380, :8].numpy(), g["dh_rows"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(float(hx.grad.double().abs().sum()), float(g["dh_checksum"]), rtol=1e-6)
    # repeat_interleave_batch: block order
    x = torch.arange(3 * B * 2, dtype=torch.float32).reshape(3 * B, 2)
    r = J.repeat_interleave_batch(x, B, 2)
    assert r.shape[0] == 6 * B and torch.equal(r[:B], x[:B]) and torch.equal(r[B:2 * B], x[:B])
    assert torch.equal(r[2 * B:3 * B], x[B:2 * B])
    # momentum update: bit-exact (three fp32 roundings)
    new = J.ema_update(q, k, float(g["momentum"]))
    for i, kn in enumerate(new):
        assert torch.equal(kn, torch.from_numpy(g[f"ema_k{i}"]))


@pytest.mark.parametrize("tag,B,seed", [("tiny", 2, 5), ("vitb", 4, 6)])
def test_mask_collator_mirror_reproduces_the_reference_masks(tag, B, seed):
    """bvc_b200.MaskCollator / update_masks (host side of the JEPA path, predictive/mask.py:21-38, :70-219) consume the
    random numbers exactly like the reference's classes: the same seed gives the index tensors the fixture generator
    got from the reference (tools/make_golden_jepa.py: torch.manual_seed(seed), first call of a fresh collator)."""
    import bvc_b200 as bvc
    g = jepa_case(tag)[0]
    torch.manual_seed(seed)
    coll = bvc.MaskCollator(input_size=(224, 224), patch_size=16, pred_mask_scale=(0.15, 0.2), enc_mask_scale=(0.85, 1.0),
                            aspect_ratio=(0.75, 1.5), nenc=1, npred=4, allow_overlap=False, min_keep=10)
    batch, m_enc, m_pred = coll([torch.zeros(1) for _ in range(B)])
    assert batch.shape == (B, 1) and len(m_enc) == 1 and len(m_pred) == 4
    m_enc = bvc.update_masks(m_enc, 224, 16, 16, 2, isencoder=True)
    m_pred = bvc.update_masks(m_pred, 224, 16, 16, 2, isencoder=False)
    assert torch.equal(torch.stack(list(m_enc)), torch.from_numpy(g["masks_enc"]))
    assert torch.equal(torch.stack(list(m_pred)), torch.from_numpy(g["masks_pred"]))
    assert int(torch.stack(list(m_pred)).min()) >= 7 * 196 and int(torch.stack(list(m_enc)).max()) < 196
    assert coll.step() == 1   # the shared counter advanced once for the batch above


@pytest.mark.parametrize("tag", ["d128", "d192_nobias"])
def test_vit_block_oracle_matches_reference_block(golden_dir, tag):
    """oracle.jepa_oracle.vit_block against the fp64 output / gradients of the reference's own `Block`
    (vision_transformer.py:213-231, fixtures by tools/make_golden_jepa_vit.py)."""
    import os
    g = np.load(os.path.join(golden_dir, "jepa_vit_block.npz"))
    dim, heads, B, N, qkv_bias = (int(v) for v in g[f"{tag}.meta"])
    params = {k: v.double().requires_grad_(True) for k, v in J.vit_block_params(dim, seed=7, qkv_bias=bool(qkv_bias)).items()}
    gen = torch.Generator().manual_seed(11)
    x = (torch.randn(B, N, dim, generator=gen) * 1.5 + 0.2).double().requires_grad_(True)
    w = torch.randn(B, N, dim, generator=gen).double()
    y = J.vit_block(x, params, heads, eps=1e-6)
    (y * w).sum().backward()
    np.testing.assert_allclose(y.detach().numpy(), g[f"{tag}.y"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(x.grad.numpy(), g[f"{tag}.dx"], rtol=0, atol=2e-5)
    for k, p in params.items():
        np.testing.assert_allclose(float(p.grad.norm()), float(g[f"{tag}.gradnorm.{k}"]), rtol=1e-9)
        if f"{tag}.grad.{k}" in g.files:
            np.testing.assert_allclose(p.grad.numpy(), g[f"{tag}.grad.{k}"], rtol=0, atol=1e-5 * float(g[f"{tag}.gradnorm.{k}"]) + 1e-7)


def test_unmasked_encoder_oracle_matches_hf_fixture(golden_dir):
    """oracle.videomae_oracle.encode_unmasked against transformers.VideoMAEModel(bool_masked_pos=None) (HF:420-470)."""
    import os
    from oracle import videomae_oracle as O
    g = np.load(os.path.join(golden_dir, "tiny_encode.npz"))
    cfg = O.make_config("tiny")
    params = O.init_params(cfg, seed=1, perturb=True)
    x = O.synthetic_clip(2, cfg, seed=9, image_like=True)
    h = O.encode_unmasked(params, x, cfg)
    np.testing.assert_allclose(h.numpy(), g["last_hidden_state"], rtol=0, atol=5e-5)
