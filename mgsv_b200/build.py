"""In-tree build of the C-ABI CUDA library (libmade_b200.so) for sm_100a.

nvcc cross-compiles without a GPU; the .so stays next to this file so that it travels to the GPU
box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmade_b200.so")
OBJ_DIR = os.path.join(HERE, "_build")
SOURCES = ["runtime.cu", "span_kernels.cu", "rank_kernels.cu", "gemm_tc.cu", "attn.cu", "xpool.cu",
           "prep.cu", "ragged.cu", "losses.cu", "ffn_fused.cu", "exact_f32.cu", "ca_fusion.cu", "api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


# MADE_DIAG=1 (with --force): compile the kernel diagnostics in — clock-stamp traces and ablation switches of the GEMM and
# X-Pool kernels (scripts/diag_gemm_trace.py, diag_xpool_trace.py, MADE_GEMM_DEBUG / MADE_XPOOL_DEBUG).  The product
# build carries none of their branches; rebuild with --force (and without MADE_DIAG) afterwards.
if os.environ.get("MADE_DIAG") == "1":
    NVCC_FLAGS += ["-DMADE_GEMM_DIAG", "-DMADE_XPOOL_DIAG", "-DMADE_XPOOL_TRACE", "-DMADE_FFN_DIAG"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: made_b200 has no non-CUDA path")


def _deps_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(HERE, "..", "include", "made_b200.h"))
    return max(os.path.getmtime(p) for p in paths)


def needs_build() -> bool:
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _deps_mtime()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libmade_b200.so. Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    newest_hdr = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    newest_hdr = max(newest_hdr, os.path.getmtime(os.path.join(HERE, "..", "include", "made_b200.h")))

    def compile_one(src: str):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        srcp = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(srcp), newest_hdr):
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, "-c", srcp, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    objs = [o for o, _ in results]
    cmd = [nvcc, "-shared", "-o", LIB + ".tmp", *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
