"""Build libdm_b200.so (the C-ABI library of include/dm_b200.h) in-tree with nvcc for sm_100a.

    python image_compression_analysis_b200/csrc/build.py [--force]

nvcc cross-compiles without a GPU.  The .so lands next to the Python package so that it travels
to the GPU box with the repo snapshot; objects go to build/ (git-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
PKG = HERE.parent
ROOT = PKG.parent
SOURCES = ["lib.cu", "stats.cu", "fused_bip.cu", "fused_bsq.cu", "validity.cu", "spectral.cu", "sobel.cu", "ssim.cu", "layout.cu", "adjacent.cu", "p2p.cu"]
HEADERS = [HERE / "dm_common.cuh", ROOT / "include" / "dm_b200.h"]
LIB = PKG / "libdm_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I", str(ROOT / "include"), "-I", str(HERE)]
if os.environ.get("DM_DEBUG_HOOKS"):      # experiment build for tools/probe_fused.py (build with --force, and again without)
    FLAGS.append("-DDM_DEBUG_HOOKS")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = True) -> Path:
    objdir = ROOT / "build" / "dm_b200"
    objdir.mkdir(parents=True, exist_ok=True)
    jobs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        if force or _stale(obj, [HERE / src, *HEADERS, Path(__file__)]):
            jobs.append([NVCC, *FLAGS, "-c", str(HERE / src), "-o", str(obj)])

    def run(cmd):
        if verbose:
            print("[build]", " ".join(cmd[-3:]), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [str(objdir / (s + ".o")) for s in SOURCES]
    if force or jobs or not LIB.exists():
        run([NVCC, "-shared", "-o", str(LIB), *objs, "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
