"""Build libCudaRasterizer.so (sm_100a) in-tree with nvcc.

The library has no Python / torch dependency: it is the C-ABI boundary of include/gm_rasterizer.h
plus the CudaRasterizer:: C++ symbols the reference's glue links against.  It is written to
gaussianmesh_b200/diff_gaussian_rasterizater/libCudaRasterizer.so -- the place and name the
reference's `-L<package dir> -lCudaRasterizer` expects (dgr/rasterize_points.py:8-10).

Usage:  python -m gaussianmesh_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "diff_gaussian_rasterizater"
LIB = OUT_DIR / "libCudaRasterizer.so"
BUILD = PKG / "build"

SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "blend_fwd.cu", "blend_bwd.cu", "geom_bwd.cu", "mesh.cu", "acap.cu", "loss.cu",
           "optimizer.cu"]

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", str(CSRC), "-I", str(PKG.parent / "include"),
] + os.environ.get("GM_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libCudaRasterizer.so")
    return nvcc


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.rglob("*.cu")) + list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.h"))
                    + [PKG.parent / "include" / "gm_rasterizer.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    BUILD.mkdir(exist_ok=True)
    stamp = BUILD / "fingerprint"
    fp = _fingerprint()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str) -> tuple[str, str]:
        obj = BUILD / (src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return src, r.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        logs = list(ex.map(compile_one, SOURCES))
    (BUILD / "ptxas.log").write_text("\n".join(f"==== {s}\n{log}" for s, log in logs))
    if verbose:
        print((BUILD / "ptxas.log").read_text())

    objs = [str(BUILD / (s[:-3] + ".o")) for s in SOURCES]
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(fp)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
