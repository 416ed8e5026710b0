#!/usr/bin/env python
"""Time one GEMM shape (optionally under BVC_GEMM_DEBUG bisect flags); used for ncu source-level captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bvc_b200 import _lib as L  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "bf16"
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 20
dev = torch.device("cuda:0")
A = torch.randn(M, K, device=dev).to(torch.bfloat16)
B = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
of = torch.zeros(M, N, device=dev)
res = torch.randn(M, N, device=dev)
aux = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
bias = torch.randn(N, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run():
    if mode == "bf16":
        L.gemm(A, B, M, N, K, out_bf16=ob, bias=bias, block_n=bn)
    elif mode == "res":
        L.gemm(A, B, M, N, K, out_f32=of, bias=bias, res=res, ldr=N, block_n=bn)
    elif mode == "gelu":
        L.gemm(A, B, M, N, K, out_bf16=ob, bias=bias, act=1, aux_out=aux, ld_aux=N, block_n=bn)
    elif mode == "dgelu":
        L.gemm(A, B, M, N, K, out_bf16=ob, act=2, aux_in=aux, ld_aux=N, block_n=bn)


for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    run()
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ts.sort()
med = ts[len(ts) // 2]
print(f"PROBE M{M} N{N} K{K} {mode} bn{bn} debug={os.environ.get('BVC_GEMM_DEBUG', '0')}: median {med*1e3:.1f} us "
      f"min {ts[0]*1e3:.1f} us  {2*M*N*K/med/1e9:.0f} TFLOP/s (L2 flushed between runs)", flush=True)
