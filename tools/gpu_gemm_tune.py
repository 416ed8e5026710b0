#!/usr/bin/env python
"""Time every GEMM shape of the ViT-B step (B=64) at each tile width; prints the best BN per shape."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bvc_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(200 << 20, dtype=torch.uint8, device=dev)
# (M, N, K, a_mn, b_mn, mode)
SHAPES = [
    (10240, 768, 1536, 0, 0, "res"), (10240, 2304, 768, 0, 0, "plain"), (10240, 768, 768, 0, 0, "res"),
    (10240, 3072, 768, 0, 0, "gelu"), (10240, 768, 3072, 0, 0, "res"),
    (10240, 3072, 768, 0, 1, "dgelu"), (10240, 768, 3072, 0, 1, "plain"), (10240, 768, 768, 0, 1, "plain"),
    (10240, 768, 2304, 0, 1, "plain"),
    (768, 3072, 10240, 1, 1, "wgrad"), (3072, 768, 10240, 1, 1, "wgrad"), (768, 768, 10240, 1, 1, "wgrad"),
    (2304, 768, 10240, 1, 1, "wgrad"), (768, 1536, 10240, 1, 1, "wgrad"),
    (100352, 1152, 384, 0, 0, "plain"), (100352, 384, 384, 0, 0, "res"), (100352, 1536, 384, 0, 0, "gelu"),
    (100352, 384, 1536, 0, 0, "res"), (90112, 1536, 384, 0, 0, "loss"),
    (100352, 1536, 384, 0, 1, "dgelu"), (100352, 384, 1536, 0, 1, "plain"), (100352, 384, 384, 0, 1, "plain"),
    (100352, 384, 1152, 0, 1, "plain"), (90112, 384, 1536, 0, 1, "plain"),
    (384, 1536, 100352, 1, 1, "wgrad"), (1536, 384, 100352, 1, 1, "wgrad"), (384, 384, 100352, 1, 1, "wgrad"),
    (1152, 384, 100352, 1, 1, "wgrad"), (1536, 384, 90112, 1, 1, "wgrad"), (10240, 384, 768, 0, 0, "plain"),
]


def run_case(M, N, K, a_mn, b_mn, mode, bn, ks=0, pair=1):
    A = torch.randn(K if a_mn else M, M if a_mn else K, device=dev).to(torch.bfloat16)
    B = (torch.randn(K if b_mn else N, N if b_mn else K, device=dev) * 0.05).to(torch.bfloat16)
    kw = dict(a_mn=bool(a_mn), b_mn=bool(b_mn), block_n=bn, cta_pair=pair)
    if mode == "wgrad":
        of = torch.zeros(M, N, device=dev)
        fn = lambda: L.gemm(A, B, M, N, K, out_f32=of, k_splits=ks, **kw)  # noqa: E731
    elif mode == "plain":
        ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        fn = lambda: L.gemm(A, B, M, N, K, out_bf16=ob, **kw)  # noqa: E731
    elif mode == "res":
        of = torch.zeros(M, N, device=dev)
        res = torch.randn(M, N, device=dev)
        bias = torch.zeros(N, device=dev)
        fn = lambda: L.gemm(A, B, M, N, K, out_f32=of, res=res, ldr=N, bias=bias, **kw)  # noqa: E731
    elif mode == "gelu":
        ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        aux = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        bias = torch.zeros(N, device=dev)
        fn = lambda: L.gemm(A, B, M, N, K, out_bf16=ob, bias=bias, act=1, aux_out=aux, ld_aux=N, **kw)  # noqa: E731
    elif mode == "dgelu":
        ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        aux = torch.randn(M, N, device=dev).to(torch.bfloat16)
        fn = lambda: L.gemm(A, B, M, N, K, out_bf16=ob, act=2, aux_in=aux, ld_aux=N, **kw)  # noqa: E731
    elif mode == "loss":
        ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        tgt = torch.randn(M, N, device=dev)
        bias = torch.zeros(N, device=dev)
        part = torch.zeros(L.gemm_loss_slots(M, N, bn), device=dev)
        fn = lambda: L.gemm(A, B, M, N, K, out_bf16=ob, bias=bias, target=tgt, ldt=N, loss_partial=part, **kw)  # noqa: E731
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[2]


for sh in SHAPES:
    M, N, K, a_mn, b_mn, mode = sh
    res = {}
    for bn in (64, 128, 192, 256):
        try:
            res[f"bn{bn}"] = run_case(M, N, K, a_mn, b_mn, mode, bn)
        except Exception as ex:  # noqa: BLE001
            res[f"bn{bn}"] = float("inf")
    for bn in (128, 192, 256):  # CTA-pair (cta_group::2) tiles, 256 x bn
        try:
            res[f"p{bn}"] = run_case(M, N, K, a_mn, b_mn, mode, bn, pair=2)
        except Exception as ex:  # noqa: BLE001
            res[f"p{bn}"] = float("inf")
    best = min(res, key=res.get)
    auto = run_case(M, N, K, a_mn, b_mn, mode, 0, pair=0)
    print(f"TUNE M{M} N{N} K{K} {'T' if a_mn else 'N'}{'T' if b_mn else 'N'} {mode:6s} " +
          " ".join(f"{b}={t*1e3:6.1f}" for b, t in res.items()) +
          f" | auto={auto*1e3:6.1f} best={best} {2*M*N*K/res[best]/1e9:6.0f} TF/s", flush=True)
