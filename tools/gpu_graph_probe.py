#!/usr/bin/env python
"""Feasibility probe: the whole training step (zero_grad, forward, loss, scaled backward, FusedSGD under GradScaler)
captured into ONE CUDA graph and replayed; compared with the eager loop on a twin model.
    python tools/gpu_graph_probe.py [batch] [steps]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bvc_b200 as bvc  # noqa: E402
from bench import CONFIGS, make_masks  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")


def build():
    torch.manual_seed(0)
    m = bvc.VideoMAEForPreTraining(bvc.VideoMAEConfig(**CONFIGS["base"])).to(dev).train()
    m.static_mask_count = True
    o = bvc.FusedSGD(m.parameters(), lr=0.1, momentum=0.9, nesterov=True, shadow_from=m)
    s = torch.amp.GradScaler("cuda")
    return m, o, s


g = torch.Generator().manual_seed(1)
clips = [torch.randn(B, 16, 3, 224, 224, generator=g).to(dev) for _ in range(2)]
masks = [make_masks(B, i).to(dev) for i in range(4)]


def step(m, o, s, x, mk):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o.zero_grad()
        loss = m(x, bool_masked_pos=mk).loss
    s.scale(loss).backward()
    s.step(o)
    s.update()
    return loss


def timeit(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, 1e3 * host / n


# eager reference trajectory
m0, o0, s0 = build()
ref = []
for i in range(4 + K):
    ref.append(float(step(m0, o0, s0, clips[i % 2], masks[i % 4])))
gpu_ms, host_ms = timeit(lambda i: step(m0, o0, s0, clips[i % 2], masks[i % 4]), K)
print(f"GRAPH-PROBE eager  B{B}: {gpu_ms:.3f} ms/step (host enqueue {host_ms:.3f} ms/step)", flush=True)
del m0, o0, s0

m1, o1, s1 = build()
sx, sm = torch.empty_like(clips[0]), torch.empty_like(masks[0])
got = []
for i in range(4):  # warm-up in eager mode on a side stream (as the torch.cuda.graphs recipe asks)
    sx.copy_(clips[i % 2])
    sm.copy_(masks[i % 4])
    got.append(float(step(m1, o1, s1, sx, sm)))
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
sx.copy_(clips[0])
sm.copy_(masks[0])
with torch.cuda.graph(graph, capture_error_mode=os.environ.get("BVC_CAPTURE_MODE", "thread_local")):
    static_loss = step(m1, o1, s1, sx, sm)
torch.cuda.synchronize()
print("GRAPH-PROBE captured", flush=True)
# NOTE: capture does not execute: replay step 4 now
for i in range(4, 4 + K):
    sx.copy_(clips[i % 2])
    sm.copy_(masks[i % 4])
    graph.replay()
    got.append(float(static_loss))
err = max(abs(a - b) / abs(b) for a, b in zip(got, ref))
print("GRAPH-PROBE losses eager", [f"{v:.6f}" for v in ref[:8]])
print("GRAPH-PROBE losses graph", [f"{v:.6f}" for v in got[:8]])
print(f"GRAPH-PROBE max rel loss diff over {len(got)} steps: {err:.2e}", flush=True)


def replay(i):
    sx.copy_(clips[i % 2])
    sm.copy_(masks[i % 4])
    graph.replay()


gpu_ms, host_ms = timeit(replay, K)
print(f"GRAPH-PROBE graph  B{B}: {gpu_ms:.3f} ms/step (host {host_ms:.3f} ms/step, incl. the 2 device-to-device input copies)", flush=True)
