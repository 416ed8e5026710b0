"""Shared helpers for the GPU parity tests: run the bvc model and the CPU oracle on identical seeded inputs."""
import numpy as np
import torch

from oracle import videomae_oracle as O


def bvc_config(cfg: O.OracleConfig):
    import bvc_b200 as bvc
    return bvc.VideoMAEConfig(
        image_size=cfg.image_size, patch_size=cfg.patch_size, num_channels=cfg.num_channels,
        num_frames=cfg.num_frames, tubelet_size=cfg.tubelet_size, hidden_size=cfg.hidden_size,
        num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
        intermediate_size=cfg.intermediate_size, decoder_num_attention_heads=cfg.decoder_num_attention_heads,
        decoder_hidden_size=cfg.decoder_hidden_size, decoder_num_hidden_layers=cfg.decoder_num_hidden_layers,
        decoder_intermediate_size=cfg.decoder_intermediate_size, layer_norm_eps=cfg.layer_norm_eps,
        norm_pix_loss=cfg.norm_pix_loss)


def run_bvc(cfg, params, x, mask, grad_scale=1.0, device="cuda:0"):
    """loss, logits, grads of the CUDA path for a state-dict / inputs given as CPU tensors."""
    import bvc_b200 as bvc
    model = bvc.VideoMAEForPreTraining(bvc_config(cfg))
    model.load_state_dict(params, strict=True)
    model = model.to(device).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):  # the reference wraps the call in autocast (:306-308)
        out = model(x.to(device), bool_masked_pos=mask.to(device))
    (out.loss * grad_scale).backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters()}
    missing = [k for k, p in model.named_parameters() if p.grad is None]
    assert not missing, missing
    return out.loss.detach().float().cpu(), out.logits.detach().float().cpu(), grads, model


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def grad_report(grads, ref):
    """per-tensor (rel-L2 error, rel norm error) and the global figures."""
    rows = {}
    num = den = 0.0
    for k, r in ref.items():
        g = grads[k].double()
        r = r.double() if torch.is_tensor(r) else torch.from_numpy(np.asarray(r)).double()
        e = float((g - r).norm())
        n = float(r.norm())
        rows[k] = (e / max(n, 1e-30), abs(float(g.norm()) - n) / max(n, 1e-30), n)
        num += e * e
        den += n * n
    return rows, (num / max(den, 1e-60)) ** 0.5


def simclr_feats(g):
    """Features of a tests/golden/simclr_*.npz fixture: stored for the small cases, regenerated from the numpy seed
    (checked against the stored checksum) for the large one -- tools/make_golden_simclr.py."""
    if "feats" in g.files:
        return torch.from_numpy(g["feats"])
    n, D, scale = int(g["n"]), int(g["D"]), float(g["scale"])
    feats = torch.from_numpy(np.random.default_rng(100 + n).standard_normal((n, D)).astype(np.float32)) * scale
    feats[1::2] = 0.7 * feats[0::2] + 0.3 * feats[1::2]
    assert abs(float(feats.double().sum()) - float(g["feats_checksum"])) < 1e-6
    return feats


def jepa_case(tag):
    """Inputs of tests/golden/jepa_<tag>.npz regenerated in the generator's draw order (tools/make_golden_jepa.py):
    returns (golden npz, h, masks_enc list, masks_pred list, z noise, gather-backward weights w, q list, k list)."""
    import os
    import numpy as np
    import torch
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"jepa_{tag}.npz"))
    B, D, N = int(g["B"]), int(g["D"]), int(g["N"])
    rng = np.random.default_rng(int(g["seed"]))
    h = torch.from_numpy(rng.standard_normal((B, N, D)).astype(np.float32) * 1.7 + 0.3)
    m_enc = [torch.from_numpy(m) for m in g["masks_enc"]]
    m_pred = [torch.from_numpy(m) for m in g["masks_pred"]]
    K = m_pred[0].shape[1]
    tshape = (len(m_pred) * len(m_enc) * B, K, D)
    noise = torch.from_numpy(rng.standard_normal(tshape).astype(np.float32)) * 0.8
    w = torch.from_numpy(rng.standard_normal((len(m_pred) * B, K, D)).astype(np.float32))
    shapes = ((D, 7), (13,), (5, D))
    q = [torch.from_numpy(rng.standard_normal(s).astype(np.float32)) for s in shapes]
    k = [torch.from_numpy(rng.standard_normal(s).astype(np.float32)) for s in shapes]
    return g, h, m_enc, m_pred, noise, w, q, k
