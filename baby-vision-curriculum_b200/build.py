"""Build libbvc.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python baby-vision-curriculum_b200/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  No torch / pybind dependency: plain
nvcc -shared with the static CUDA runtime; the library is loaded with ctypes (see _lib.py).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libbvc.so")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["gemm.cu", "gemm_bn64.cu", "gemm_bn128.cu", "gemm_bn192.cu", "gemm_bn256.cu", "gemm_pair128.cu", "gemm_pair192.cu", "gemm_pair256.cu", "rows.cu", "patchify.cu", "attn.cu", "attn_small.cu", "optim.cu", "nce.cu", "jepa.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# --use_fast_math only where the hot loops are exponentials on the MUFU pipe (softmax, GELU) feeding bf16 outputs: the
# tensor-core kernels.  The fp32 statistics / optimizer / index kernels (rows, patchify, optim, nce, jepa) keep IEEE
# division, square root and denormals.
FAST_MATH = ("gemm", "attn")


def _flags(src):
    return FLAGS + (["--use_fast_math"] if os.path.basename(src).startswith(FAST_MATH) else [])


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update((" ".join(FLAGS) + "|" + ",".join(FAST_MATH)).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "bvc.h"))
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "stamp")
    dig = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(OBJ_DIR, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [NVCC] + _flags(s) + ["-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(s)}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {s}")
    with open(os.path.join(OBJ_DIR, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
