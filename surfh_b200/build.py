"""Build libsurfh_b200.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the tree)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "libsurfh_b200.so")
SOURCES = ["capi.cu"]
HEADERS = ["common.cuh", "host_util.cuh", "fft_plan.cuh", "kernels_fft.cuh", "kernels_lmm.cuh", "kernels_slit.cuh", "kernels_gemm.cuh", "kernels_gemm_tma.cuh", "kernels_ozaki.cuh", "kernels_precond.cuh", "kernels_shepard.cuh", "kernels_cg.cuh",
           os.path.join("..", "..", "include", "surfh_b200.h")]


def _cufft_dir() -> str:
    """cuFFT ships with the torch wheels (nvidia-cufft); the toolkit here has only the header."""
    try:
        import nvidia.cufft as m
        d = os.path.join(os.path.dirname(m.__file__), "lib")
        if os.path.exists(os.path.join(d, "libcufft.so.11")):
            return d
    except ImportError:
        pass
    for d in ("/usr/local/cuda/lib64", "/usr/local/cuda/targets/x86_64-linux/lib"):
        if os.path.exists(os.path.join(d, "libcufft.so")) or os.path.exists(os.path.join(d, "libcufft.so.11")):
            return d
    raise RuntimeError("libcufft not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """`out` / `defines` build an experimental variant beside the product library (kernel A/B runs)."""
    if out == OUT and not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cufft = _cufft_dir()
    lib = "-l:libcufft.so.11" if os.path.exists(os.path.join(cufft, "libcufft.so.11")) else "-lcufft"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-D" + d for d in defines]
    cmd += ["-o", out, "-L" + cufft, lib, "-Xlinker", "-rpath=" + cufft]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libsurfh_b200.so")
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
