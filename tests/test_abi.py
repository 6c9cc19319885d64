"""CPU: the C-ABI library loads and exports every symbol include/mmd_retrieval.h declares; argument
validation and the no-GPU / no-fallback behaviour (no compute is attempted without a GPU)."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT
from mmd_retrieval import _lib

HEADER = os.path.join(ROOT, "include", "mmd_retrieval.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MMD_API\s+[\w\s\*]+?\b(mmd_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_symbols()
    assert len(names) >= 14
    assert set(names) == set(_lib.SIGNATURES), "include/mmd_retrieval.h and mmd_retrieval/_lib.py disagree"


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(str(_lib.LIB_PATH)) if _lib.LIB_PATH.exists() else _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"libmmd.so does not export {name}"
    assert _lib.load().mmd_abi_version() == 1
    assert _lib.load().mmd_topk_max_k() == 120


def test_prepared_layout():
    from mmd_retrieval import ops
    assert ops.prepared_layout("bf16", 768) == (768, 1536)
    assert ops.prepared_layout("fp16", 2048) == (2048, 4096)
    assert ops.prepared_layout("bf16", 100) == (104, 208)          # padded to 16 bytes
    assert ops.prepared_layout("fp8", 100) == (112, 112)
    assert ops.prepared_layout("fp32", 768) == (6 * 768, 6 * 768 * 2)   # 3 bf16 limbs, 6 cross terms
    lib = _lib.load()
    assert lib.mmd_prepared_layout(99, 768, None, None) == -1
    assert "op_dtype" in _lib.last_error()


def test_argument_validation_without_touching_the_gpu():
    lib = _lib.load()
    assert lib.mmd_normalize_cast(None, 0, -1, 768, 768, 1, 1e-12, 0, 0, None, None, None) == -1
    assert lib.mmd_normalize_cast(None, 0, 10, 768, 768, 1, 1e-12, 0, 0, None, None, None) == -1   # null buffers
    assert lib.mmd_topk_scores(None, None, 0, 4, 4, 8, 0, 0, None, None, None, 0, None) == -1       # k = 0
    assert lib.mmd_topk_scores(None, None, 0, 4, 2 ** 31, 8, 1, 0, None, None, None, 0, None) == -1  # N too large
    assert lib.mmd_topk_merge(None, None, 0, 4, 4, 4, None, None, None) == -1
    assert lib.mmd_rescore(None, 0, 8, None, None, 0, 8, None, 4, 4, 8, None, 2000, 0, 4, None, None, None) == -1
    assert lib.mmd_topk_workspace_bytes(16384, 1000000, 768, 0, 18) > 0
    assert lib.mmd_topk_workspace_bytes(16384, 1000000, 768, 0, 500) == 0      # beyond the fused K limit
    assert lib.mmd_launch_count() >= 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a host WITHOUT a GPU")
def test_no_gpu_means_error_not_fallback():
    import mmd_retrieval as m
    lib = _lib.load()
    assert lib.mmd_device_check() == -2
    assert "no CPU fallback" in _lib.last_error()
    q, c = torch.randn(4, 16), torch.randn(32, 16)
    with pytest.raises(m.MmdError):
        m.topk(q, c, 3)
    with pytest.raises(m.MmdError):
        m.semantic_search(q, c, top_k=3)
    with pytest.raises(m.MmdError):
        m.prepare_corpus(c, device="cpu")
    with pytest.raises(m.MmdError):
        m.ImageCorpus(feature_dict={"a": torch.ones(8)}).retrieve_similar_features(torch.ones(1, 8), 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multimodal-misinformation-detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"
