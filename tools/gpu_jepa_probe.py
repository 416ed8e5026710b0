#!/usr/bin/env python
"""Time the JEPA pieces at ViT-B sizes (B = 64, N = 1568, D = 768; 4 prediction masks of 36, 1 context mask of 71;
86 M encoder parameters for the EMA) and print achieved HBM GB/s (algorithmic bytes / CUDA-event time)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bvc_b200 as bvc  # noqa: E402

dev = torch.device("cuda:0")
B, N, D = 64, 1568, 768
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e3  # us


g = torch.Generator(device=dev).manual_seed(0)
h = torch.randn(B, N, D, device=dev, generator=g)
mp = [torch.stack([torch.randperm(196, device=dev, generator=g)[:36] + 7 * 196 for _ in range(B)]) for _ in range(4)]
me = [torch.stack([torch.randperm(196, device=dev, generator=g)[:71] for _ in range(B)])]
us = timeit(lambda: bvc.jepa_targets(h, mp, 1))
by = 4 * B * 36 * D * 8
print(f"PROBE jepa_targets (LN + gather + repeat) rows {4*B*36}: {us:.1f} us  {by/us/1e3:.0f} GB/s")
us = timeit(lambda: bvc.apply_masks(h, me))
by = B * 71 * D * 8
print(f"PROBE apply_masks rows {B*71}: {us:.1f} us  {by/us/1e3:.0f} GB/s")
t = bvc.jepa_targets(h, mp, 1)
z = (t + 0.5 * torch.randn_like(t)).to(torch.bfloat16)
us = timeit(lambda: bvc.smooth_l1_loss(z, t))
print(f"PROBE smooth_l1 fwd n {t.numel()}: {us:.1f} us  {t.numel()*6/us/1e3:.0f} GB/s")
shapes = [(768, 768)] * 48 + [(3072, 768)] * 12 + [(768, 3072)] * 12 + [(768,)] * 100 + [(768, 1536)]
q = [torch.randn(s, device=dev) for s in shapes]
k = [torch.randn(s, device=dev) for s in shapes]
n = sum(p.numel() for p in q)
us = timeit(lambda: bvc.ema_update(q, k, 0.997))
print(f"PROBE ema_update {n/1e6:.1f} M params, {len(q)} tensors, one launch: {us:.1f} us  {n*12/us/1e3:.0f} GB/s")
k2 = [p.clone() for p in k]


def torch_ema():
    for pq, pk in zip(q, k2):
        pk.data.mul_(0.997).add_((1. - 0.997) * pq.detach().data)


us = timeit(torch_ema)
print(f"PROBE torch EMA loop (the reference's formulation): {us:.1f} us")
