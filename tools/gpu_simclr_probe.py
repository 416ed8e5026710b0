#!/usr/bin/env python
"""Time bvc_b200.info_nce_loss (forward + backward) against the reference's formulation run with torch on the same GPU
(F.cosine_similarity(feats[:, None], feats[None]) materialises n x n x D) at n = 2 x batch, D = 512."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bvc_b200 as bvc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
masks = bvc.make_simclr_masks(n, dev)
feats = torch.randn(n, D, device=dev)


def ref_loss(f):  # pretrain_simclr.py:114-128
    cos = F.cosine_similarity(f[:, None, :], f[None, :, :], dim=-1) / 0.1
    return torch.logsumexp(cos[masks[1]], dim=-1) - cos[masks[0]].mean()


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def step(loss_fn):
    f = feats.clone().requires_grad_(True)
    loss_fn(f).backward()
    return f.grad


a = timeit(lambda: step(lambda f: bvc.info_nce_loss(0.1, masks, f)))
try:
    b = timeit(lambda: step(ref_loss), iters=3)
except RuntimeError as ex:  # out of memory at large n
    b = float("nan")
    print("reference formulation failed:", str(ex)[:80])
ga, gb = step(lambda f: bvc.info_nce_loss(0.1, masks, f)), step(ref_loss)
print(f"PROBE simclr loss n{n} D{D}: bvc fwd+bwd {a*1e3:.1f} us | torch reference formulation {b*1e3:.1f} us "
      f"({b/a:.0f}x) | grad rel-L2 vs torch fp32 {float((ga-gb).norm()/gb.norm()):.2e}", flush=True)
