"""GPU parity of bvc_b200.info_nce_loss (libbvc.so: bvc_nce_* + bvc_gemm_bf16) against the CPU oracle
(oracle/simclr_oracle.py, fp64) and the golden fixtures produced by the reference's own info_nce_loss
(tools/make_golden_simclr.py), on identical features and masks, at a tiny size, the reference's size (batch 32 -> n = 64,
D = 512) and BASELINE.json config 3's size (batch 512 -> n = 1024).

Tolerance (north_star: loss / gradients within 1e-3 relative): loss 2e-5 relative, gradient rel-L2 1e-3.  The similarity
GEMM runs on bf16 tensor cores with hi + lo operand splits; measured on B200: loss <= 3e-6, gradients <= 3e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import simclr_oracle as SO
from tests.helpers import rel_l2, simclr_feats

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["tiny", "ref", "cfg3"])
def test_info_nce_vs_oracle_and_golden(golden_dir, tag):
    import bvc_b200 as bvc
    g = np.load(os.path.join(golden_dir, f"simclr_{tag}.npz"))
    feats = simclr_feats(g)
    n = feats.shape[0]
    T = float(g["temperature"])
    ref_loss, ref_grad = SO.loss_and_grad(T, SO.make_masks(n), feats)
    dev = torch.device("cuda:0")
    masks = bvc.make_simclr_masks(n, dev)
    assert torch.equal(masks[0].cpu(), SO.make_masks(n)[0]) and torch.equal(masks[1].cpu(), SO.make_masks(n)[1])
    f = feats.to(dev).requires_grad_(True)
    loss = bvc.info_nce_loss(T, masks, f)
    (loss * 3.0).backward()  # an arbitrary upstream gradient (GradScaler)
    rl = abs(float(loss) - float(ref_loss)) / abs(float(ref_loss))
    rg = rel_l2(f.grad.cpu() / 3.0, ref_grad)
    print(f"[{tag}] loss {float(loss):.6f} ref {float(ref_loss):.6f} rel {rl:.2e}; grad rel-L2 {rg:.2e}")
    assert rl <= 2e-5 and rg <= 1e-3
    # and against what the reference itself printed (fp64 run of pretrain_simclr.info_nce_loss)
    assert abs(float(loss) - float(g["loss_f64"])) <= 2e-5 * abs(float(g["loss_f64"]))
    gn = float(f.grad.double().norm()) / 3.0
    assert abs(gn - float(g["grad_norm_f64"])) <= 1e-3 * float(g["grad_norm_f64"])


def test_info_nce_bf16_features_and_general_masks():
    """bf16 features (what the model emits under autocast) and arbitrary (non-symmetric) masks."""
    import bvc_b200 as bvc
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(5)
    n, D = 48, 64
    feats = torch.randn(n, D, generator=gen).to(torch.bfloat16)
    pos = torch.rand(n, n, generator=gen) < 0.05
    neg = (torch.rand(n, n, generator=gen) < 0.6) & ~pos
    pos[0, 1] = True
    ref_loss, ref_grad = SO.loss_and_grad(0.2, (pos, neg), feats.float())
    f = feats.to(dev).requires_grad_(True)
    loss = bvc.info_nce_loss(0.2, (pos.to(dev), neg.to(dev)), f)
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-4 * abs(float(ref_loss))
    assert f.grad.dtype == torch.bfloat16
    assert rel_l2(f.grad.float().cpu(), ref_grad) <= 6e-3  # bf16 rounding of the returned gradient
