"""Build libmmd.so (the C-ABI CUDA library) in-tree for sm_100a.

    python multimodal-misinformation-detection_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so lands next to the Python host layer
(mmd_retrieval/libmmd.so) so that it travels with the repo snapshot to the GPU box.

Provenance: the library is rebuilt whenever the SHA-256 over csrc/*, include/*, the compiler flags and the nvcc version
differs from the one recorded next to the binary (libmmd.so.buildinfo) -- not when an mtime says so -- and the same string
is compiled into the binary (mmd_build_info()), so a bench line can say exactly which sources produced its numbers.
Concurrent builders (every rank of a torchrun job importing the package at once) serialise on a file lock and write
through per-process temporary names.
"""
from __future__ import annotations

import concurrent.futures as cf
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
INCLUDE = HERE.parent / "include"
OUT_DIR = HERE / "mmd_retrieval"
LIB = Path(os.environ["MMD_BUILD_OUT"]).resolve() if os.environ.get("MMD_BUILD_OUT") else OUT_DIR / "libmmd.so"   # developer builds
INFO = LIB.with_name(LIB.name + ".buildinfo")
OBJ_DIR = HERE / "build"

SOURCES = ["api.cu", "normalize.cu", "topk_fused.cu", "topk_merge.cu", "rescore.cu", "dedupe.cu", "exchange.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "--expt-relaxed-constexpr", "-DMMD_BUILDING"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _nvcc_version(nvcc: str) -> str:
    out = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout
    for line in out.splitlines():
        if "release" in line:
            return line.split("release")[-1].strip().replace(" ", "")
    return "unknown"


def _extra_flags() -> list:
    """Developer builds: MMD_STATS=1 (per-tile timeline in the fused kernel), MMD_DEFINES="A=1,B=2" (extra -D switches)."""
    flags = ["-DMMD_STATS"] if os.environ.get("MMD_STATS") else []
    flags += [f"-D{d}" for d in os.environ.get("MMD_DEFINES", "").split(",") if d]
    return flags


def source_hash(nvcc_version: str) -> str:
    """SHA-256 over every file that determines the binary: sources, headers, flags, compiler version."""
    h = hashlib.sha256()
    files = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h")))
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(ARCH + NVCC_FLAGS + _extra_flags() + SOURCES).encode())
    h.update(nvcc_version.encode())
    return h.hexdigest()[:16]


def build_info_string(nvcc: str) -> str:
    ver = _nvcc_version(nvcc)
    return f"src={source_hash(ver)} nvcc={ver} arch=sm_100a flags={'_'.join(f.lstrip('-') for f in NVCC_FLAGS[:3] + _extra_flags())}"


def up_to_date(nvcc: str | None = None) -> bool:
    if not LIB.exists() or not INFO.exists():
        return False
    try:
        return INFO.read_text().strip() == build_info_string(nvcc or _nvcc())
    except Exception:  # noqa: BLE001
        return False


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    if not force and up_to_date(nvcc):
        return LIB
    OBJ_DIR.mkdir(exist_ok=True)
    with open(OBJ_DIR / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)                       # one builder at a time; the others find it up to date
        try:
            if not force and up_to_date(nvcc):
                return LIB
            info = build_info_string(nvcc)
            extra = (["-Xptxas", "-v"] if verbose else []) + _extra_flags() + [f'-DMMD_BUILD_INFO="{info}"']
            tag = f".{os.getpid()}"

            def compile_one(src: str) -> Path:
                obj = OBJ_DIR / (Path(src).stem + tag + ".o")
                only = ["-DMMD_BUILD_INFO=\"" + info + "\""] if src == "api.cu" else []
                flags = [f for f in extra if not f.startswith("-DMMD_BUILD_INFO")] + only
                cmd = [nvcc, *ARCH, *NVCC_FLAGS, *flags, "-c", str(CSRC / src), "-o", str(obj)]
                r = subprocess.run(cmd, capture_output=True, text=True)
                if verbose or r.returncode != 0:
                    sys.stderr.write(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}\n")
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
                return obj

            with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
                objs = list(ex.map(compile_one, SOURCES))
            tmp = LIB.with_suffix(f".so{tag}.tmp")
            cmd = [nvcc, *ARCH, "-shared", "-o", str(tmp), *map(str, objs)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            for o in objs:
                o.unlink(missing_ok=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("link failed")
            os.replace(tmp, LIB)
            INFO.write_text(info + "\n")
            return LIB
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    p = build_lib(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
    print(INFO.read_text().strip())
