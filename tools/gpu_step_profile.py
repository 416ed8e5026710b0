#!/usr/bin/env python
"""Kineto (CUPTI) timeline of the bench training step: every GPU kernel of 3 steps, libbvc and torch-native alike
(optimizer / GradScaler / memsets), aggregated by name, plus the idle share of the stream.  Not a bench number --
profiler overhead is inside the wall time; it answers "what runs that the libbvc event profiler does not see".
    python tools/gpu_step_profile.py [--batch 64] [--fused-opt]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bvc_b200 as bvc  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--fused-opt", action="store_true")
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = bvc.VideoMAEForPreTraining(bvc.VideoMAEConfig(**bench.CONFIGS["base"])).to(dev).train()
if a.fused_opt:
    opt = bvc.FusedSGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True)
else:
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True)
scaler = torch.amp.GradScaler("cuda")
x = torch.randn(a.batch, 16, 3, 224, 224, device=dev)
np.random.seed(0)
m = bvc.batch_masks(bvc.TubeMaskingGenerator((8, 14, 14), 0.9), a.batch).to(dev)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        opt.zero_grad()
        loss = model(x, bool_masked_pos=m).loss
        loss = bvc.AllReduce.apply(loss)
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
agg = {}
for e in evs:
    k = e.name[:70]
    v = agg.setdefault(k, [0, 0.0])
    v[0] += 1
    v[1] += e.time_range.end - e.time_range.start
print(f"STEPPROF span {(t1 - t0) / a.steps / 1e3:.3f} ms/step  busy {busy / a.steps / 1e3:.3f} ms/step  "
      f"idle {(t1 - t0 - busy) / a.steps / 1e3:.3f} ms/step  kernels/step {len(evs) / a.steps:.0f}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"STEPPROF {us / a.steps / 1e3:8.3f} ms/step  x{n / a.steps:6.1f}  {k}")
# gaps: idle time between consecutive kernels, by the kernel that follows the gap
gaps = {}
for p, q in zip(evs[:-1], evs[1:]):
    g = q.time_range.start - p.time_range.end
    if g > 0:
        v = gaps.setdefault(q.name[:50], [0, 0.0])
        v[0] += 1
        v[1] += g
for k, (n, us) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"STEPGAP  {us / a.steps / 1e3:8.3f} ms/step  x{n / a.steps:6.1f}  before {k}")
