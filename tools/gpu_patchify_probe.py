#!/usr/bin/env python
"""Time bvc_patchify_target at the headline shape (B=64, 16x224x224, mask 0.9); ncu target."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bvc_b200 as bvc  # noqa: E402
from bvc_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
x = torch.randn(B, 16, 3, 224, 224, device=dev)
np.random.seed(0)
mask = bvc.batch_masks(bvc.TubeMaskingGenerator((8, 14, 14), 0.9), B).to(dev).view(torch.uint8)
nv, N = 160, 1568
vis = torch.zeros(B, nv, device=dev, dtype=torch.int32)
msk = torch.zeros(B, N - nv, device=dev, dtype=torch.int32)
slot = torch.zeros(B, N, device=dev, dtype=torch.int32)
status = torch.zeros(1, device=dev, dtype=torch.int32)
L.mask_to_index(mask, nv, vis, msk, slot, status)
pv = torch.zeros(B * nv, 1536, device=dev, dtype=torch.bfloat16)
tgt = torch.zeros(B * (N - nv), 1536, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(2):
    L.patchify_target(x, slot, 2, 16, nv, pv, tgt, True)
ts = []
for _ in range(iters):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    L.patchify_target(x, slot, 2, 16, nv, pv, tgt, True)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ts.sort()
by = x.numel() * 4 + pv.numel() * 2 + tgt.numel() * 4
print(f"PROBE patchify_target B{B}: median {ts[len(ts)//2]*1e3:.1f} us min {ts[0]*1e3:.1f} us  "
      f"{by/ts[len(ts)//2]/1e6:.0f} GB/s ({by/1e6:.0f} MB algorithmic)", flush=True)
