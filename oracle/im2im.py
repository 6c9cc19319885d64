"""The reference's image retrieval on CPU.  TEST INFRASTRUCTURE ONLY.

Two things live here:

1. `load_reference_module()` imports the reference's OWN, unmodified `src/evidence/im2im_retrieval.py` from
   /root/reference (present only in the build container).  The module needs `matplotlib` (absent: stubbed with
   empty modules) and its classes download ResNet-50 weights in `__init__` (no network: instances are made with
   `object.__new__` and given a feature dict + a stub extractor).  `reference_retrieve(...)` then runs the
   reference's own `ImageSimilarity.similarity` (im2im_retrieval.py:38-42) and
   `ImageCorpus.retrieve_similar_images` (im2im_retrieval.py:80-106).  oracle/make_golden.py uses this to produce
   tests/golden/im2im_*.npz, which PIN the restatement below and the CUDA path.

2. A restatement that needs nothing but torch, for the GPU box where /root/reference does not exist:
     similarity(f1, f2)                      nn.CosineSimilarity(dim=1, eps=1e-6) on two vectors   (:38-42)
     retrieve_similar(query, feature_dict)   loop over the dict, full sort, first-of-each-score dedupe (:80-106)
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Dict, List, Tuple

import torch

REFERENCE_ROOT = os.environ.get("MMD_REFERENCE_ROOT", "/root/reference")
IMAGE_EPS = 1e-6


# ---------------------------------------------------------------- 1. the reference's own code
def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "evidence", "im2im_retrieval.py"))


def load_reference_module():
    if not reference_available():
        raise RuntimeError(f"reference checkout not found under {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:  # noqa: BLE001  (absent in this image)
                sys.modules[name] = types.ModuleType(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("src.evidence.im2im_retrieval")


class _StubExtractor:
    """Stands in for the ResNet-50 encoder: 'paths' are keys into a dict of precomputed features."""

    def __init__(self, ref_similarity_cls, features: Dict[str, torch.Tensor]):
        self._impl = object.__new__(ref_similarity_cls)   # no __init__: no weight download
        self._features = features

    def extract_features(self, path):
        return self._features[path]

    def similarity(self, f1, f2):
        return type(self._impl).similarity(self._impl, f1, f2)   # the reference's own method


def reference_retrieve(feature_dict: Dict[str, torch.Tensor], query_features: Dict[str, torch.Tensor], top_k: int
                       ) -> Dict[str, List[Tuple[str, float]]]:
    """Run the reference's unmodified retrieve_similar_images for every query key."""
    mod = load_reference_module()
    corpus = object.__new__(mod.ImageCorpus)
    corpus.feature_corpus_path = None
    corpus.feature_dict = feature_dict
    corpus.feature_extractor = _StubExtractor(mod.ImageSimilarity, query_features)
    return {qk: mod.ImageCorpus.retrieve_similar_images(corpus, qk, top_k=top_k) for qk in query_features}


def reference_similarity(f1: torch.Tensor, f2: torch.Tensor) -> float:
    mod = load_reference_module()
    return mod.ImageSimilarity.similarity(object.__new__(mod.ImageSimilarity), f1, f2)


# ---------------------------------------------------------------- 2. restatement (runs anywhere)
def similarity(f1: torch.Tensor, f2: torch.Tensor) -> float:
    cos = torch.nn.CosineSimilarity(dim=1, eps=IMAGE_EPS)
    return cos(f1.unsqueeze(0), f2.unsqueeze(0)).item()


def retrieve_similar(query: torch.Tensor, feature_dict: Dict[str, torch.Tensor], top_k: int = 50
                     ) -> List[Tuple[str, float]]:
    scores = {name: similarity(query, feat) for name, feat in feature_dict.items()}
    ranked = sorted(scores.items(), key=lambda kv: kv[1], reverse=True)
    seen, kept = set(), []
    for name, score in ranked:
        if score not in seen:
            seen.add(score)
            kept.append((name, score))
        if len(kept) == top_k:
            break
    return kept


def retrieve_similar_batched(queries: torch.Tensor, feature_dict: Dict[str, torch.Tensor], top_k: int = 50
                             ) -> List[List[Tuple[str, float]]]:
    """Same result as retrieve_similar for every row of `queries`, but one fp32 matmul instead of Q*N python
    iterations (checked equal to the loop in tests/test_oracle.py); this is what bench.py times as the CPU baseline."""
    keys = list(feature_dict.keys())
    c = torch.stack([feature_dict[k].float() for k in keys])
    qn = queries.float() / queries.float().norm(dim=1, keepdim=True).clamp_min(IMAGE_EPS)
    cn = c / c.norm(dim=1, keepdim=True).clamp_min(IMAGE_EPS)
    s = qn @ cn.T
    vals, idx = torch.sort(s, dim=1, descending=True, stable=True)
    out = []
    for r in range(s.shape[0]):
        seen, kept = set(), []
        for v, i in zip(vals[r].tolist(), idx[r].tolist()):
            if v not in seen:
                seen.add(v)
                kept.append((keys[i], v))
            if len(kept) == top_k:
                break
        out.append(kept)
    return out
