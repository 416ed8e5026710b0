"""GPU parity of the JEPA pieces through the C-ABI: bvc_b200.apply_masks /
repeat_interleave_batch / jepa_targets / smooth_l1_loss / ema_update against the CPU oracle and the reference-made
fixtures.  Index, copy and EMA work must be BIT-EXACT; layer-norm targets within 2e-6 absolute (fp32 statistics),
the loss within 2e-6 relative, its gradient within 1e-6 relative."""
import numpy as np
import pytest
import torch

from oracle import jepa_oracle as J
from tests.helpers import jepa_case

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("tag", ["tiny", "vitb"])
def test_jepa_pieces_match_oracle_and_fixtures(tag):
    import bvc_b200 as bvc
    dev = _dev()
    g, h, m_enc, m_pred, noise, w, q, k = jepa_case(tag)
    B, N, D = h.shape
    hd = h.to(dev)
    me, mp = [m.to(dev) for m in m_enc], [m.to(dev) for m in m_pred]
    # apply_masks: bit-exact, fp32 and bf16
    ctx = bvc.apply_masks(hd, me)
    assert torch.equal(ctx.cpu(), J.apply_masks(h, m_enc))
    assert float(ctx.double().sum()) == float(g["ctx_checksum"])
    hb = h.to(torch.bfloat16)
    assert torch.equal(bvc.apply_masks(hb.to(dev), mp).cpu(), J.apply_masks(hb, m_pred))
    # its backward: same accumulation order as autograd -> bit-exact in fp32
    hx = hd.clone().requires_grad_(True)
    (bvc.apply_masks(hx, mp) * w.to(dev)).sum().backward()
    hr = h.clone().requires_grad_(True)
    (J.apply_masks(hr, m_pred) * w).sum().backward()
    assert torch.equal(hx.grad.cpu(), hr.grad)
    np.testing.assert_allclose(hx.grad[:, 1372:1380, :8].cpu().numpy(), g["dh_rows"], rtol=0, atol=1e-6)
    # repeat_interleave_batch: bit-exact, with autograd
    x = J.apply_masks(h, m_pred)
    xr = x.to(dev).requires_grad_(True)
    r = bvc.repeat_interleave_batch(xr, B, 3)
    assert torch.equal(r.cpu(), J.repeat_interleave_batch(x, B, 3))
    r.sum().backward()
    assert torch.equal(xr.grad.cpu(), torch.full_like(x, 3.0))
    # fused target branch (layer_norm + apply_masks + repeat)
    t = bvc.jepa_targets(hd, mp, len(m_enc))
    t_or = J.jepa_targets(h.double(), m_pred, len(m_enc))
    np.testing.assert_allclose(t.cpu().numpy(), t_or.numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(t[:, :2, :8].cpu().numpy(), g["targets_head"], rtol=0, atol=2e-6)
    t1 = bvc.jepa_targets(hd, mp, 1)
    assert torch.equal(bvc.jepa_targets(hd, mp, 3).cpu(), J.repeat_interleave_batch(t1.cpu(), B, 3))
    tb = bvc.jepa_targets(hb.to(dev), mp, 1)
    np.testing.assert_allclose(tb.cpu().numpy(), J.jepa_targets(hb.double(), m_pred, 1).numpy(), rtol=0, atol=2e-6)
    # smooth-L1 forward / backward with an upstream gradient
    t_ref = J.jepa_targets(h, m_pred, len(m_enc))
    z = (t_ref + noise).to(dev).requires_grad_(True)
    loss = bvc.smooth_l1_loss(z, t_ref.to(dev))
    (loss * 3.0).backward()
    zo = (t_ref + noise).double().requires_grad_(True)
    lo = J.smooth_l1_loss(zo, t_ref.double())
    (lo * 3.0).backward()
    assert abs(float(loss) - float(lo)) / float(lo) < 2e-6
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 5e-6
    gerr = (z.grad.cpu().double() - zo.grad).norm() / zo.grad.norm()
    assert float(gerr) < 1e-6
    zb = (t_ref + noise).to(torch.bfloat16).to(dev).requires_grad_(True)   # bf16 predictor output (autocast)
    lb = bvc.smooth_l1_loss(zb, t_ref.to(dev))
    lb.backward()
    zbo = zb.detach().cpu().double().requires_grad_(True)
    lbo = J.smooth_l1_loss(zbo, t_ref.double())
    lbo.backward()
    assert abs(float(lb) - float(lbo)) / float(lbo) < 2e-6
    assert zb.grad.dtype == torch.bfloat16
    assert float((zb.grad.cpu().double() - zbo.grad).norm() / zbo.grad.norm()) < 4e-3   # bf16 rounding of dz
    # momentum update: bit-exact against torch's mul_ / add_ and the fixture
    qd, kd = [p.to(dev) for p in q], [p.to(dev) for p in k]
    bvc.ema_update(qd, kd, float(g["momentum"]))
    for i, kn in enumerate(kd):
        assert torch.equal(kn.cpu(), torch.from_numpy(g[f"ema_k{i}"]))
    assert all(torch.equal(a.cpu(), b) for a, b in zip(qd, q))


def test_jepa_edge_cases():
    import bvc_b200 as bvc
    dev = _dev()
    x = torch.randn(2, 10, 8, device=dev)
    with pytest.raises(ValueError):
        bvc.apply_masks(x, [])
    with pytest.raises(ValueError):
        bvc.apply_masks(x, [torch.zeros(2, 3, dtype=torch.int64, device=dev), torch.zeros(2, 4, dtype=torch.int64, device=dev)])
    with pytest.raises(ValueError):
        bvc.repeat_interleave_batch(torch.randn(5, 8, device=dev), 2, 2)
    with pytest.raises(bvc.BvcError):
        bvc.apply_masks(torch.randn(2, 10, 8), [torch.zeros(2, 3, dtype=torch.int64)])   # CPU tensors: no fallback
    # a single mask, K = 1, duplicated index across masks, large row (D = 1024)
    xl = torch.randn(3, 7, 1024, device=dev)
    m = [torch.tensor([[6], [0], [3]], device=dev), torch.tensor([[6], [1], [3]], device=dev)]
    out = bvc.apply_masks(xl, m)
    assert torch.equal(out.cpu(), J.apply_masks(xl.cpu(), [t.cpu() for t in m]))
    t = bvc.jepa_targets(xl, m, 2)
    np.testing.assert_allclose(t.cpu().numpy(), J.jepa_targets(xl.cpu().double(), [t_.cpu() for t_ in m], 2).numpy(),
                               rtol=0, atol=3e-6)
    # odd element count through the scalar tail of the loss kernels
    z = torch.randn(7, 3, device=dev, requires_grad=True)
    h = torch.randn(7, 3, device=dev)
    l = bvc.smooth_l1_loss(z, h, beta=0.5)
    l.backward()
    zr = z.detach().cpu().double().requires_grad_(True)
    lr = torch.nn.functional.smooth_l1_loss(zr, h.cpu().double(), beta=0.5)
    lr.backward()
    assert abs(float(l) - float(lr)) < 1e-6 and float((z.grad.cpu().double() - zr.grad).abs().max()) < 1e-7


# ------------------------------------------------------------------------------------------------------------------
# the predictive path's ViT block on the VideoMAE step's kernels (bvc_b200.jepa_vit)
class _RefLikeAttention(torch.nn.Module):
    """Structure of vision_transformer.py:186-197 (the reference tree is absent on the GPU box)."""

    def __init__(self, dim, heads, qkv_bias):
        super().__init__()
        self.num_heads, self.scale = heads, (dim // heads) ** -0.5
        self.qkv = torch.nn.Linear(dim, 3 * dim, bias=qkv_bias)
        self.attn_drop = torch.nn.Dropout(0.)
        self.proj = torch.nn.Linear(dim, dim)
        self.proj_drop = torch.nn.Dropout(0.)


class _RefLikeMLP(torch.nn.Module):
    def __init__(self, dim, ff):
        super().__init__()
        self.fc1, self.act, self.fc2, self.drop = torch.nn.Linear(dim, ff), torch.nn.GELU(), torch.nn.Linear(ff, dim), torch.nn.Dropout(0.)


class _RefLikeBlock(torch.nn.Module):
    def __init__(self, dim, heads, qkv_bias):
        super().__init__()
        self.norm1 = torch.nn.LayerNorm(dim, eps=1e-6)
        self.attn = _RefLikeAttention(dim, heads, qkv_bias)
        self.drop_path = torch.nn.Identity()
        self.norm2 = torch.nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _RefLikeMLP(dim, 4 * dim)


@pytest.mark.parametrize("tag", ["d128", "d192_nobias"])
def test_jepa_vit_block_matches_reference_fixture_and_oracle(golden_dir, tag):
    """bvc_b200.jepa_vit.Block (fused qkv Linear, LayerNorm eps 1e-6) against the fp64 output / gradients of the
    reference's own Block (fixture) and the live oracle.  Tolerances: bf16 operands, fp32 accumulation -- output
    rel-L2 5e-3, input gradient 1.5e-2, parameter gradients 2e-2 per tensor (norms 5e-3)."""
    import os
    from functools import partial
    import bvc_b200 as bvc
    from tests.helpers import rel_l2
    dev = _dev()
    g = np.load(os.path.join(golden_dir, "jepa_vit_block.npz"))
    dim, heads, B, N, qkv_bias = (int(v) for v in g[f"{tag}.meta"])
    params = J.vit_block_params(dim, seed=7, qkv_bias=bool(qkv_bias))
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(B, N, dim, generator=gen) * 1.5 + 0.2
    w = torch.randn(B, N, dim, generator=gen)
    blk = bvc.jepa_vit.Block(dim, heads, mlp_ratio=4.0, qkv_bias=bool(qkv_bias),
                             norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    blk.load_state_dict(params, strict=True)
    blk = blk.to(dev)
    xd = x.to(dev).requires_grad_(True)
    y = blk(xd)
    (y * w.to(dev)).sum().backward()
    assert rel_l2(y.detach().cpu(), torch.from_numpy(g[f"{tag}.y"])) <= 5e-3
    assert rel_l2(xd.grad.cpu(), torch.from_numpy(g[f"{tag}.dx"])) <= 1.5e-2
    for k, p in blk.named_parameters():
        n_ref = float(g[f"{tag}.gradnorm.{k}"])
        assert abs(float(p.grad.double().norm()) - n_ref) <= 5e-3 * n_ref, k
        if f"{tag}.grad.{k}" in g.files:
            assert rel_l2(p.grad.cpu(), torch.from_numpy(g[f"{tag}.grad.{k}"])) <= 2e-2, k
    # live oracle (fp64) on the same inputs
    pd = {k: v.double() for k, v in params.items()}
    assert rel_l2(y.detach().cpu(), J.vit_block(x.double(), pd, heads, eps=1e-6)) <= 5e-3
    # from_reference: adopts the sub-modules of a reference-structured block (parameters shared, state-dict keys equal)
    ref_like = _RefLikeBlock(dim, heads, bool(qkv_bias))
    ref_like.load_state_dict(params, strict=True)
    ref_like = ref_like.to(dev)
    holder = torch.nn.Module()
    holder.blocks = torch.nn.ModuleList([ref_like])
    bvc.jepa_vit.convert_blocks(holder)
    conv = holder.blocks[0]
    assert isinstance(conv, bvc.jepa_vit.Block) and conv.attn.qkv.weight is ref_like.attn.qkv.weight
    assert set(conv.state_dict()) == set(params)
    y2 = conv(x.to(dev))
    assert torch.equal(y2, y.detach())
    # two chained blocks: the activation-gradient side channel (bf16 twin + column sums) is shared between them
    blk2 = bvc.jepa_vit.Block(dim, heads, qkv_bias=bool(qkv_bias), norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    blk2.load_state_dict(params, strict=True)
    blk2 = blk2.to(dev)
    blk.zero_grad(set_to_none=True)
    xc = x.to(dev).requires_grad_(True)
    (blk2(blk(xc)) * w.to(dev)).sum().backward()
    xo = x.double().requires_grad_(True)
    (J.vit_block(J.vit_block(xo, pd, heads, eps=1e-6), pd, heads, eps=1e-6) * w.double()).sum().backward()
    assert rel_l2(xc.grad.cpu(), xo.grad) <= 2e-2


def test_unmasked_encoder_pass(golden_dir):
    """model.encode(x) == HF VideoMAEModel(bool_masked_pos=None).last_hidden_state (compute_embeddings_videomae.py:261):
    tiny configuration against the HF-made fixture and the live oracle, and ViT-S at ALL 1568 tokens against the live
    oracle (fp32 CPU).  Tolerance: rel-L2 1e-2 (bf16 operands through 12 blocks of x4-perturbed weights; measured on
    B200: 6.1e-3 at ViT-S / 1568 tokens)."""
    import os
    import bvc_b200 as bvc
    from oracle import videomae_oracle as O
    from tests.helpers import bvc_config, rel_l2
    dev = _dev()
    g = np.load(os.path.join(golden_dir, "tiny_encode.npz"))
    cfg = O.make_config("tiny")
    params = O.init_params(cfg, seed=1, perturb=True)
    x = O.synthetic_clip(2, cfg, seed=9, image_like=True)
    model = bvc.VideoMAEForPreTraining(bvc_config(cfg))
    model.load_state_dict(params, strict=True)
    model = model.to(dev).eval()
    with torch.no_grad():
        h = model.encode(x.to(dev))
    assert h.shape == (2, cfg.seq_len, cfg.hidden_size) and h.dtype == torch.float32
    assert rel_l2(h.cpu(), torch.from_numpy(g["last_hidden_state"])) <= 1e-2
    cfg = O.make_config("small")
    params = O.init_params(cfg, seed=2, perturb=True)
    x = O.synthetic_clip(2, cfg, seed=3, image_like=True)
    model = bvc.VideoMAEForPreTraining(bvc_config(cfg))
    model.load_state_dict(params, strict=True)
    model = model.to(dev).eval()
    with torch.no_grad():
        h = model.encode(x.to(dev))
        ref = O.encode_unmasked(params, x, cfg)
    assert h.shape == (2, 1568, 384)
    assert rel_l2(h.cpu(), ref) <= 1e-2
