"""Kernel-level parity of every libbvc.so entry point against fp64 torch references on the same seeded inputs
(tools/gpu_selftest.py holds the cases: GEMM in all four operand-major combinations, on one-CTA tiles and on CTA pairs
(tcgen05 cta_group::2, group "gemmpair"), every fused epilogue incl. split-K,
GELU / GELU' / residual / row-gather / segment scatter / fused MSE / fused column sums, LayerNorm forward / backward with
segment remaps, column sums, casts, decoder mask rows, tube-mask indexing + patchify + normalised-pixel target
(bit-exact indices and visible-patch rows), attention forward / backward at S = 8 ... 1568 incl. ragged tiles and the
short-sequence kernel of attn_small.cu).
Group "benchshapes" runs the kernel shapes of the benchmark step itself (ViT-B/16 at batch 64: M = 10240 / 100352 /
90112 rows, attention at B64 S160 H12 and B64 S1568 H6, patchify at B = 64) with the tile shapes the dispatcher picks.
Each group runs in its own process so that a device trap in one group cannot poison the others; a group passes when
every case printed PASS.  The end-to-end step parity lives in test_model_gpu.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("group", ["rows", "patchify", "gemm00", "gemm01", "gemm10", "gemm11", "gemmx", "gemmpair", "attn",
                                   "benchshapes"])
def test_kernel_group(group):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_selftest.py"), group], capture_output=True,
                       text=True, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith(("PASS", "FAIL", "SELFTEST"))]
    fails = [ln for ln in lines if ln.startswith("FAIL")]
    assert r.returncode == 0 and not fails and any(ln.startswith("SELFTEST") for ln in lines), \
        "\n".join(fails or lines[-5:]) + "\n" + r.stderr[-2000:]
    assert sum(ln.startswith("PASS") for ln in lines) >= 4
