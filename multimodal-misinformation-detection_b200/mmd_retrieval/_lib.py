"""ctypes binding of libmmd.so (the C ABI declared in include/mmd_retrieval.h).

There is no CPU fallback: if the shared library is missing and cannot be built, or the CUDA device is not
sm_100, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["MMD_LIB_PATH"]) if os.environ.get("MMD_LIB_PATH") else _HERE / "libmmd.so"   # override: developer builds

MMD_OK = 0
SRC_F32, SRC_F16, SRC_BF16 = 0, 1, 2
OP_BF16, OP_F16, OP_E4M3, OP_BF16X3 = 0, 1, 2, 3
SIDE_QUERY, SIDE_CORPUS = 0, 1

#: every symbol include/mmd_retrieval.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mmd_abi_version": (C.c_int, []),
    "mmd_last_error": (C.c_char_p, []),
    "mmd_device_check": (C.c_int, []),
    "mmd_prepared_layout": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mmd_normalize_cast": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mmd_topk_max_k": (C.c_int, []),
    "mmd_topk_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int]),
    "mmd_topk_scores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mmd_topk_scores_shared": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p), C.c_int,
                                         C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_void_p]),
    "mmd_scores_dense": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p,
                                   C.c_int64, C.c_void_p]),
    "mmd_topk_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "mmd_rescore": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                              C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "mmd_topk_merge_pairs": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "mmd_scatter_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_void_p]),
    "mmd_rescore_pairs": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                    C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int,
                                    C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_void_p]),
    "mmd_normalize_cast_segment": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_float,
                                             C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "mmd_rescore_joint": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                    C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mmd_rescore_joint_pairs": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                          C.c_int64, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_void_p]),
    "mmd_sharded_candidates": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p), C.c_int,
                                         C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "mmd_zero_u32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "mmd_exchange_rescore": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                       C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int64, C.c_int64, C.POINTER(C.c_void_p), C.c_int,
                                       C.c_int, C.c_int64, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p,
                                       C.c_void_p]),
    "mmd_exchange_finish": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p]),
    "mmd_dedupe_scores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "mmd_launch_count": (C.c_int64, []),
    "mmd_build_info": (C.c_char_p, []),
    "mmd_profile_enable": (C.c_int, [C.c_int]),
    "mmd_profile_collect": (C.c_int, [C.POINTER(C.c_float), C.c_int]),
}

_lock = threading.Lock()
_lib = None


class MmdError(RuntimeError):
    """A libmmd call failed (status code + mmd_last_error text)."""


def _try_build() -> None:
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mmd_build", _HERE.parent / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_lib()


def load():
    """Load (building first if the .so is absent) and return the ctypes library handle."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if os.environ.get("MMD_LIB_PATH") or os.environ.get("MMD_NO_AUTOBUILD"):
            if not LIB_PATH.exists():
                raise MmdError(f"{LIB_PATH} is missing and auto-build is off; there is no CPU fallback")
        else:
            # build_lib() is a no-op when the recorded source hash matches; it rebuilds a missing OR stale library
            # (one builder at a time under a file lock: every rank of a torchrun job lands here at once)
            try:
                _try_build()
            except Exception as e:  # noqa: BLE001
                if not LIB_PATH.exists():
                    raise MmdError(f"{LIB_PATH} is missing and building it failed ({e}); there is no CPU fallback") from e
                import warnings
                warnings.warn(f"libmmd.so may be stale: rebuilding failed ({e})")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            if os.environ.get("MMD_LIB_PARTIAL") and not hasattr(lib, name):
                continue                 # developer A/B runs against an older build of the library
            fn = getattr(lib, name)  # AttributeError here = header and library disagree
            fn.restype = res
            fn.argtypes = args
        if lib.mmd_abi_version() != 1:
            raise MmdError(f"libmmd ABI version {lib.mmd_abi_version()} != 1")
        _lib = lib
        return _lib


def build_info() -> str:
    """Provenance string compiled into the loaded binary (source hash, nvcc version, arch)."""
    return load().mmd_build_info().decode("utf-8", "replace")


def last_error() -> str:
    return load().mmd_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != MMD_OK:
        raise MmdError(f"{what} failed with status {rc}: {last_error()}")
