#!/usr/bin/env python
"""Pure-write / pure-read / copy HBM bandwidth of this B200 (torch kernels, CUDA events, best of 10): the context for
the store-heavy GEMM epilogues (act + pre-activation writes) and the streaming kernels."""
import torch

dev = torch.device("cuda:0")
n = 1 << 29  # 512 Mi elements
a = torch.empty(n, dtype=torch.bfloat16, device=dev)
b = torch.empty(n, dtype=torch.bfloat16, device=dev)


def t(fn, reps=10):
    for _ in range(2):
        fn()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


w = t(lambda: a.zero_())
c = t(lambda: b.copy_(a))
r = t(lambda: torch.sum(a.view(torch.int16)[: n // 2].view(torch.int32)))  # read-only pass over 0.5 GB
print(f"MEMBW write-only (zero_ 1 GiB): {2*n/w/1e6:.0f} GB/s | copy (1 GiB read + 1 GiB write): {4*n/c/1e6:.0f} GB/s | "
      f"read-only (sum over 0.5 GiB): {n/r/1e6:.0f} GB/s", flush=True)
