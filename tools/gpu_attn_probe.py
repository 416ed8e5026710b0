#!/usr/bin/env python
"""Time attention forward / backward at one shape; target for ncu source-level captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bvc_b200 import _lib as L  # noqa: E402

B, S, H = (int(v) for v in sys.argv[1:4])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = torch.device("cuda:0")
d = H * 64
qkv = torch.randn(B, S, 3, H, 64, device=dev).to(torch.bfloat16)
out = torch.zeros(B, S, d, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(B, H, S, device=dev)
do = torch.randn(B, S, d, device=dev).to(torch.bfloat16)
dqkv = torch.zeros_like(qkv)
delta = torch.zeros(B, H, S, device=dev)
scale = 0.125


def t(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


f = t(lambda: L.attn_fwd(qkv, B, S, H, scale, out, lse))
b = t(lambda: L.attn_bwd(qkv, out, do, lse, B, S, H, scale, delta, dqkv))
fl = 4.0 * B * H * S * S * 64
print(f"PROBE attn B{B} S{S} H{H}: fwd {f*1e3:.1f} us {fl/f/1e9:.0f} TFLOP/s | bwd {b*1e3:.1f} us {2*fl/b/1e9:.0f} TFLOP/s (credited 2x fwd)",
      flush=True)
if S > 160:  # with the fp32 dQ workspace: the one-pass kernel (memset + kernel + convert)
    acc = torch.empty(B, S, H, 64, device=dev)
    b1 = t(lambda: L.attn_bwd(qkv, out, do, lse, B, S, H, scale, delta, dqkv, acc))
    print(f"PROBE attn B{B} S{S} H{H}: bwd one-pass {b1*1e3:.1f} us {2*fl/b1/1e9:.0f} TFLOP/s (credited 2x fwd)", flush=True)
