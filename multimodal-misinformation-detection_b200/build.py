"""Build libmmd.so (the C-ABI CUDA library) in-tree for sm_100a.

    python multimodal-misinformation-detection_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so lands next to the Python host layer
(mmd_retrieval/libmmd.so) so that it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT_DIR = HERE / "mmd_retrieval"
LIB = OUT_DIR / "libmmd.so"
OBJ_DIR = HERE / "build"

SOURCES = ["api.cu", "normalize.cu", "topk_fused.cu", "topk_merge.cu", "rescore.cu", "dedupe.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "--expt-relaxed-constexpr", "-DMMD_BUILDING"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _newest_source_mtime() -> float:
    files = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "mmd_retrieval.h",
                                                                 Path(__file__)]
    return max(f.stat().st_mtime for f in files)


def up_to_date() -> bool:
    return LIB.exists() and LIB.stat().st_mtime >= _newest_source_mtime()


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    if not force and up_to_date():
        return LIB
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    if os.environ.get("MMD_STATS"):
        extra.append("-DMMD_STATS")      # developer build: wait-cycle counters in the fused kernel

    def compile_one(src: str) -> Path:
        obj = OBJ_DIR / (Path(src).stem + ".o")
        cmd = [nvcc, *ARCH, *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}\n")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB.with_suffix(".so.tmp")
    cmd = [nvcc, *ARCH, "-shared", "-o", str(tmp), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    p = build_lib(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
