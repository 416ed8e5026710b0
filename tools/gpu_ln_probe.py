#!/usr/bin/env python
"""Time LayerNorm forward/backward and colsum at the model's shapes (ncu target)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bvc_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, iters=10):
    for _ in range(2):
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
    return ts[len(ts) // 2]


for M, d in ((10240, 768), (100352, 384)):
    x = torch.randn(M, d, device=dev)
    gam, bet = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    y = torch.zeros(M, d, device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.zeros(M, device=dev), torch.zeros(M, device=dev)
    ms = t(lambda: L.layernorm_fwd(x, gam, bet, 1e-12, M, d, y, mean, rstd))
    print(f"PROBE ln_fwd {M}x{d}: {ms*1e3:.1f} us {M*d*6/ms/1e6:.0f} GB/s", flush=True)
    dxf, dxb = torch.zeros_like(x), torch.zeros_like(y)
    dg, db, ds = (torch.zeros(d, device=dev) for _ in range(3))
    ms = t(lambda: L.layernorm_bwd(y, x, mean, rstd, gam, x, M, d, dxf, dxb, dg, db, dxsum=ds))
    print(f"PROBE ln_bwd {M}x{d}: {ms*1e3:.1f} us {M*d*16/ms/1e6:.0f} GB/s", flush=True)
    out = torch.zeros(d, device=dev)
    ms = t(lambda: L.colsum(y, M, d, out))
    print(f"PROBE colsum bf16 {M}x{d}: {ms*1e3:.1f} us {M*d*2/ms/1e6:.0f} GB/s", flush=True)
