"""Build the in-tree native artefacts.

  build_cuda()   -> vierkant_b200/lib/libvierkant_bcn_cuda.so   (the product: nvcc, sm_100a only)
  build_host()   -> vierkant_b200/lib/libvierkant_bcn_host.so   (C++20 drop-in of vierkant::bcn::compress over the C ABI)

nvcc cross-compiles without a GPU.  Numerics flags: -fmad=false -prec-div=true -prec-sqrt=true (the device code
additionally uses explicit round-to-nearest intrinsics everywhere a result must match the reference bit for bit) and
-ffp-contract=off for the host-side table/weight setup.
"""
from __future__ import annotations

import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vierkant_b200", "csrc")
LIBDIR = os.path.join(ROOT, "vierkant_b200", "lib")
CUDA_SO = os.path.join(LIBDIR, "libvierkant_bcn_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def cuda_sources() -> list[str]:
    out = []
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".cpp", ".h")):
            out.append(os.path.join(CSRC, name))
    out.append(os.path.join(ROOT, "include", "vierkant_bcn_cuda.h"))
    return out


def find_nvcc() -> str | None:
    nvcc = shutil.which("nvcc")
    if nvcc is None and os.path.exists("/usr/local/cuda/bin/nvcc"):
        nvcc = "/usr/local/cuda/bin/nvcc"
    return nvcc


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    """Compile libvierkant_bcn_cuda.so for sm_100a.  Returns the path."""
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = cuda_sources()
    if not force and _newer(CUDA_SO, srcs):
        return CUDA_SO
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: libvierkant_bcn_cuda.so cannot be built (and there is no CPU fallback)")
    extra = os.environ.get("VKT_NVCC_EXTRA", "").split()  # tuning experiments only (e.g. -DVKT_BC7_THREADS=128)
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", os.environ.get("VKT_CUDA_SO_OUT", CUDA_SO),
           os.path.join(CSRC, "bcn_cuda.cu"), os.path.join(CSRC, "bc7_tables.cpp")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.run(cmd, check=True, cwd=ROOT)
    return CUDA_SO
