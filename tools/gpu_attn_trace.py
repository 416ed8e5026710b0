#!/usr/bin/env python
"""Per-phase SM-clock trace of the attention pipelines (CTA 0; mode 1 = backward KV pass, 0 = backward Q pass,
2 = forward: columns are wait S | LDTM | max + rescale | exp + pack | STTM for softmax group 0): builds csrc/attn.cu with -DBVC_TRACE into
libbvc_trace.so (done on the build box: `nvcc ... -DBVC_TRACE -shared csrc/attn.cu csrc/attn_small.cu -o libbvc_trace.so`; set BVC_ATTN_SMALL=0 to
trace the general kernels at S <= 192), runs one
pass and prints, per streamed tile, how long each role spent in each phase.
    python tools/gpu_attn_trace.py B S H mode_kv
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "baby-vision-curriculum_b200", os.environ.get("BVC_TRACE_LIB", "libbvc_trace.so")))
B, S, H, mode = (int(v) for v in sys.argv[1:5])
dev = torch.device("cuda:0")
d = H * 64
qkv = torch.randn(B, S, 3, H, 64, device=dev).to(torch.bfloat16)
do = torch.randn(B, S, d, device=dev).to(torch.bfloat16)
lse = torch.randn(B, H, S, device=dev) + 5
delta = torch.randn(B, H, S, device=dev)
dqkv = torch.zeros_like(qkv)
P = lambda t: C.c_void_p(t.data_ptr())
out = torch.zeros(B, S, d, device=dev, dtype=torch.bfloat16)


def run_pass():
    if mode == 2:  # forward
        return lib.bvc_debug_attn_fwd(P(qkv), B, S, H, C.c_float(0.125), P(out), P(lse), None)
    return lib.bvc_debug_attn_bwd_pass(P(qkv), P(do), P(lse), P(delta), B, S, H, C.c_float(0.125), P(dqkv), mode, None)


for _ in range(3):
    rc = run_pass()
    assert rc == 0, rc
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
run_pass()
e.record()
torch.cuda.synchronize()
if False:
  for _ in range(3):
    rc = lib.bvc_debug_attn_bwd_pass(P(qkv), P(do), P(lse), P(delta), B, S, H, C.c_float(0.125), P(dqkv), mode, None)
    assert rc == 0, rc
print(f"TRACE pass mode_kv={mode} B{B} S{S} H{H}: {s.elapsed_time(e)*1e3:.1f} us")
R, T, K = 3, 256, 8
buf = np.zeros(R * T * K, dtype=np.int64)
lib.bvc_debug_trace_copy(buf.ctypes.data_as(C.POINTER(C.c_longlong)), buf.size)
tr = buf.reshape(R, T, K)
n_it = (S + 127) // 128
n_items = (((S + 127) // 128) * H * B + 147) // 148
if mode == 2:
    n_items = ((((S + 127) // 128 + 1) // 2) * H * B + 147) // 148
G = min(T, n_it * n_items)
t0 = tr[0, 0, 0]
print("TRACE compute warp (half 0): per tile: wait_sdp | ldtm | math | wait_pds_free | sttm ; [epilogue wait | store]")
for g in range(min(G, 40)):
    c = tr[0, g]
    ep = f" | EPI wait {c[6]-c[5]} store {c[7]-c[6]}" if (g % n_it) == n_it - 1 else ""
    m = tr[2, g]
    print(f"TRACE g{g:3d} t={c[0]-t0:7d} C: {c[1]-c[0]:5d} {c[2]-c[1]:5d} {c[3]-c[2]:5d} {c[4]-c[3]:5d} {c[5]-c[4]:5d}{ep}"
          f"   M: t={m[0]-t0:7d} wait_tma {m[1]-m[0]:5d} wait_sdp_free {m[2]-m[1]:5d} issue_sdp {m[3]-m[2]:5d} "
          f"wait_pds_full {m[4]-m[3]:5d} issue_acc {m[5]-m[4]:5d}")
# averages over steady-state tiles
cs = tr[0, 2:G]
print("TRACE avg compute:", [int(np.mean(cs[:, i + 1] - cs[:, i])) for i in range(5)], "period",
      int(np.mean(np.diff(tr[0, 2:G, 0]))))
ms = tr[2, 2:G - 1]
print("TRACE avg mma    :", [int(np.mean(ms[:, i + 1] - ms[:, i])) for i in range(5)], "period",
      int(np.mean(np.diff(tr[2, 2:G - 1, 0]))))
