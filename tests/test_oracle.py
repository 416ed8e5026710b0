"""Pin the CPU oracle (oracle/videomae_oracle.py) against fixtures produced by the real reference code
(tools/make_golden.py: HF transformers VideoMAEForPreTraining + the reference's mask.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import videomae_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_tube_and_random_masks_match_reference_generators(golden_dir):
    g = _load(golden_dir, "masks.npz")
    for name, size, ratio in (("tube_8x14x14_r90", (8, 14, 14), 0.9), ("tube_1x14x14_r90", (1, 14, 14), 0.9),
                              ("tube_2x2x2_r50", (2, 2, 2), 0.5)):
        np.random.seed(0)
        got = np.stack([O.tube_mask(size, ratio) for _ in range(4)]).astype(np.uint8)
        assert np.array_equal(got, g[name]), name
    np.random.seed(0)
    got = np.stack([O.random_mask((8, 14, 14), 0.9) for _ in range(2)]).astype(np.uint8)
    assert np.array_equal(got, g["random_8x14x14_r90"])


def test_tube_mask_structure():
    np.random.seed(7)
    m = O.tube_mask((8, 14, 14), 0.9).reshape(8, 196)
    assert m.dtype == np.float64 and m.sum() == 8 * 176
    assert (m == m[0]).all()  # one pattern repeated in every temporal slot


def test_sinusoid_table_bit_equal(golden_dir):
    g = _load(golden_dir, "sinusoid.npz")
    for n, d in ((1568, 768), (1568, 384), (196, 192), (8, 64)):
        t = O.sinusoid_table(n, d).numpy()
        rows = g[f"n{n}_d{d}_rows"]
        assert np.array_equal(t[rows], g[f"n{n}_d{d}"]), (n, d)


def test_mask_to_index_is_boolean_indexing():
    torch.manual_seed(0)
    np.random.seed(1)
    mask = O.batch_tube_masks(3, (2, 3, 3), 0.5)
    x = torch.randn(3, 18, 5)
    vis, msk = O.mask_to_index(mask)
    assert torch.equal(x[~mask].reshape(3, -1, 5), torch.gather(x, 1, vis.long()[:, :, None].expand(-1, -1, 5)))
    assert torch.equal(x[mask].reshape(3, -1, 5), torch.gather(x, 1, msk.long()[:, :, None].expand(-1, -1, 5)))
    bad = mask.clone()
    bad[0, :] = False
    with pytest.raises(ValueError):
        O.mask_to_index(bad)


def test_patch_embed_order_is_conv3d():
    cfg = O.make_config("tiny")
    x = O.synthetic_clip(2, cfg, seed=3)
    w = torch.randn(8, 3, 2, 16, 16)
    ref = torch.nn.functional.conv3d(x.permute(0, 2, 1, 3, 4), w, stride=(2, 16, 16)).flatten(2).transpose(1, 2)
    got = O.patchify_embed_order(x, cfg) @ w.reshape(8, -1).T
    assert torch.allclose(ref, got, atol=2e-4, rtol=1e-4)


def test_norm_pix_target_matches_hf_labels(golden_dir):
    g = _load(golden_dir, "tiny_labels.npz")
    cfg = O.make_config("tiny")
    x = O.synthetic_clip(2, cfg, seed=5, image_like=True)
    mask = torch.from_numpy(g["mask"])
    _, msk = O.mask_to_index(mask)
    got = O.norm_pix_target(x, msk, cfg).numpy()
    assert np.array_equal(got, g["labels"])  # same fp32 ops in the same order -> bit-equal


@pytest.mark.parametrize("tag,perturb", [("init", False), ("perturbed", True)])
def test_tiny_step_loss_logits_grads(golden_dir, tag, perturb):
    g = _load(golden_dir, "tiny_step.npz")
    cfg = O.make_config("tiny")
    params = O.init_params(cfg, seed=1, perturb=perturb)
    x = O.synthetic_clip(3, cfg, seed=2, image_like=perturb)
    mask = torch.from_numpy(g[f"{tag}.mask"])
    np.random.seed(3)
    assert torch.equal(mask, O.batch_tube_masks(3, cfg.grid, 0.5))
    loss, logits, grads = O.grads_of(params, x, mask, cfg)
    assert abs(float(loss) - float(g[f"{tag}.loss"])) <= 2e-6 * abs(float(g[f"{tag}.loss"]))
    np.testing.assert_allclose(logits.numpy(), g[f"{tag}.logits"], rtol=1e-4, atol=2e-5)
    for k, v in grads.items():
        ref = g[f"{tag}.grad.{k}"]
        denom = max(np.linalg.norm(ref), 1e-12)
        assert np.linalg.norm(v.numpy() - ref) / denom < 2e-5, k


def test_small_step_summary(golden_dir):
    """BASELINE.json configs[0]: ViT-S/16, 16x224x224, tube mask 0.9, batch 2 (fp32 CPU)."""
    with open(os.path.join(golden_dir, "small_step.json")) as f:
        g = json.load(f)
    cfg = O.make_config("small")
    params = O.init_params(cfg, seed=0, perturb=True)
    x = O.synthetic_clip(2, cfg, seed=0, image_like=True)
    np.random.seed(0)
    mask = O.batch_tube_masks(2, cfg.grid, 0.9)
    loss, logits, grads = O.grads_of(params, x, mask, cfg)
    ref = g["perturbed"]
    assert abs(float(loss) - ref["loss"]) <= 1e-5 * ref["loss"]
    samp = logits.flatten()[::ref["logits_sample_stride"]][:64].numpy()
    np.testing.assert_allclose(samp, np.array(ref["logits_sample"]), rtol=2e-3, atol=2e-4)
    for k, v in grads.items():
        assert abs(float(v.double().norm()) - ref["grad_norms"][k]) <= 1e-4 * ref["grad_norms"][k] + 1e-9, k


def test_loss_allreduce_is_mean():
    assert abs(O.loss_allreduce([1.0, 3.0]) - 2.0) < 1e-12


# ------------------------------------------------------------------------------------------------ SimCLR loss oracle
@pytest.mark.parametrize("tag", ["tiny", "ref", "cfg3"])
def test_simclr_oracle_vs_reference_golden(golden_dir, tag):
    """oracle/simclr_oracle.py against fixtures produced by the reference's own info_nce_loss / get_special_matrix
    (tools/make_golden_simclr.py): masks bit-equal, loss and gradients to fp64 round-off."""
    import numpy as np
    import torch
    from oracle import simclr_oracle as SO
    g = np.load(os.path.join(golden_dir, f"simclr_{tag}.npz"))
    from tests.helpers import simclr_feats
    feats = simclr_feats(g)
    n = feats.shape[0]
    assert np.array_equal(SO.get_special_matrix(min(n, 16)), g["special"])
    loss, grad = SO.loss_and_grad(float(g["temperature"]), SO.make_masks(n), feats)
    assert abs(float(loss) - float(g["loss_f64"])) <= 1e-10 * abs(float(g["loss_f64"]))
    assert abs(float(grad.norm()) - float(g["grad_norm_f64"])) <= 1e-9 * float(g["grad_norm_f64"])
    ref = torch.from_numpy(g["grad_f64"])
    got = grad if n <= 64 else grad[:8]
    assert float((got - ref).norm() / ref.norm()) <= 1e-9
    # the reference's own fp32 run differs from fp64 by this much (context for the GPU tolerance)
    assert abs(float(g["loss_f32"]) - float(g["loss_f64"])) <= 1e-5 * abs(float(g["loss_f64"]))


# ------------------------------------------------------------------------- the reference's single-frame control condition
@pytest.mark.parametrize("perturb", [False, True])
def test_single_frame_config_vs_live_hf(perturb):
    """The reference's complexity-control runs train the same model on single frames (num_frames = 1, tubelet_size = 1:
    N = 196, patch vector 768; slurmscripts/complexity_control/slurm_dev_mst.bash).  The committed
    fixtures only cover tubelet 2, so this case pins the restatement on the real HF model run live (transformers is
    part of the image; fp32 CPU): loss to 2e-6, logits 1e-4, gradients 2e-5 rel-L2."""
    transformers = pytest.importorskip("transformers")
    cfg = O.make_config("tiny", num_frames=1, tubelet_size=1, image_size=64)
    assert cfg.grid == (1, 4, 4) and cfg.patch_dim == 768
    params = O.init_params(cfg, seed=1, perturb=perturb)
    x = O.synthetic_clip(3, cfg, seed=2, image_like=perturb)
    np.random.seed(3)
    mask = O.batch_tube_masks(3, cfg.grid, 0.75)
    hf = transformers.VideoMAEForPreTraining(transformers.VideoMAEConfig(
        image_size=cfg.image_size, num_frames=cfg.num_frames, tubelet_size=cfg.tubelet_size, hidden_size=cfg.hidden_size,
        num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
        intermediate_size=cfg.intermediate_size, use_mean_pooling=True,
        decoder_num_attention_heads=cfg.decoder_num_attention_heads, decoder_hidden_size=cfg.decoder_hidden_size,
        decoder_num_hidden_layers=cfg.decoder_num_hidden_layers,
        decoder_intermediate_size=cfg.decoder_intermediate_size, norm_pix_loss=True))
    hf.load_state_dict(params)
    hf.train()
    out = hf(x, bool_masked_pos=mask)
    out.loss.backward()
    loss, logits, grads = O.grads_of(params, x, mask, cfg)
    assert abs(float(loss) - float(out.loss.detach())) <= 2e-6 * abs(float(out.loss.detach()))
    np.testing.assert_allclose(logits.numpy(), out.logits.detach().numpy(), rtol=1e-4, atol=2e-5)
    for k, p in hf.named_parameters():
        ref = p.grad
        assert float((grads[k] - ref).norm()) <= 2e-5 * max(float(ref.norm()), 1e-12), k
