#!/usr/bin/env python
"""Fit used by csrc/gemm.cu: Q(a) = -log2(Phi(-a)) on [0, 6] as a degree-7 polynomial (reweighted least squares on
Chebyshev nodes ~ minimax in Q, i.e. uniform RELATIVE error of Phi(-a) including the tails).  GELU(x) = x*Phi(x) then
costs 7 FMAs and one ex2.  Prints the coefficients and the error figures quoted in the kernel comment."""
import numpy as np
from scipy import special

A, DEG = 6.0, 7
Qf = lambda a: -special.log_ndtr(-a) / np.log(2)  # noqa: E731
n = 400
a = (np.cos(np.pi * (np.arange(n) + 0.5) / n) + 1) * A / 2
V, y, w = np.vander(a, DEG + 1, increasing=True), Qf(a), np.ones(n)
for _ in range(30):
    coef = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)[0]
    err = np.abs(V @ coef - y)
    w = w * (1 + err / err.max())
aa = np.linspace(0, A, 200001)
ww = 2.0 ** (-np.polyval(coef[::-1], aa))
print("coefficients (increasing power):", repr(coef))
print("max rel err Phi(-a):", np.abs(ww / special.ndtr(-aa) - 1).max())
print("max abs err GELU(x<0):", np.abs(-aa * ww + aa * special.ndtr(-aa)).max())
