"""Drop-in for `sentence_transformers.util.semantic_search` (boundary B1 of SURVEY.md section 8b).

The reference calls it once per claim (src/evidence/text2text_retrieval.py:56-64,
src/evidence/experiment_text.py:25-33):

    hits = util.semantic_search(question_embedding, self.train_embeddings, top_k=top_k * 5)[0]

Same signature, same return type (`list[list[{"corpus_id": int, "score": float}]]`, one inner list per
query, min(top_k, N) hits sorted by score descending, corpus_id = 0-based corpus row).  The whole batch
goes through the fused CUDA path in one call; `query_chunk_size` / `corpus_chunk_size` are accepted and
ignored (the kernel tiles internally; results do not depend on chunking).  The prepared (normalised, cast,
HBM-resident) corpus is cached per corpus tensor, because the reference passes the same corpus on every
call and upstream re-normalises it every time.
"""
from __future__ import annotations

import weakref
from typing import Callable, Dict, List, Optional, Union

import torch

from . import _lib, ops


def cos_sim(a, b) -> torch.Tensor:
    """cos_sim(a, b)[i, j] = cosine(a[i], b[j])  (upstream util.cos_sim: F.normalize both, then mm)."""
    return ops.dense_scores(a, b, metric="cos", dtype="fp32")


def dot_score(a, b) -> torch.Tensor:
    """dot_score(a, b)[i, j] = <a[i], b[j]>  (upstream util.dot_score)."""
    return ops.dense_scores(a, b, metric="dot", dtype="fp32")


_METRIC_OF = {cos_sim: "cos", dot_score: "dot", "cos_sim": "cos", "dot_score": "dot", "cos": "cos", "dot": "dot"}

# prepared corpora keyed by the identity + version of the caller's tensor
_cache: Dict[tuple, ops.PreparedCorpus] = {}
_cache_refs: Dict[tuple, weakref.ref] = {}
_CACHE_MAX = 8


def _cache_key(t: torch.Tensor, dtype: str, metric: str, device) -> tuple:
    return (id(t), t.data_ptr(), tuple(t.shape), t.dtype, t._version, dtype, metric, str(device))


def _prepared_for(corpus, dtype: str, metric: str, device) -> ops.PreparedCorpus:
    if isinstance(corpus, ops.PreparedCorpus):
        return corpus
    if not isinstance(corpus, torch.Tensor):
        return ops.prepare_corpus(corpus, dtype=dtype, metric=metric, device=device)
    key = _cache_key(corpus, dtype, metric, device)
    hit = _cache.get(key)
    if hit is not None and _cache_refs[key]() is corpus:
        return hit
    pc = ops.prepare_corpus(corpus, dtype=dtype, metric=metric, device=device)
    if len(_cache) >= _CACHE_MAX:
        old = next(iter(_cache))
        _cache.pop(old, None)
        _cache_refs.pop(old, None)
    _cache[key] = pc
    _cache_refs[key] = weakref.ref(corpus, lambda _r, k=key: (_cache.pop(k, None), _cache_refs.pop(k, None)))
    return pc


def clear_cache() -> None:
    _cache.clear()
    _cache_refs.clear()


def semantic_search(query_embeddings, corpus_embeddings, query_chunk_size: int = 100, corpus_chunk_size: int = 500000,
                    top_k: int = 10, score_function: Union[Callable, str] = cos_sim, *, dtype: str = "bf16",
                    device: Optional[Union[str, torch.device]] = None) -> List[List[Dict[str, Union[int, float]]]]:
    """Cosine (or dot-product) top-k search of every query against the whole corpus on the GPU."""
    del query_chunk_size, corpus_chunk_size
    try:
        metric = _METRIC_OF[score_function]
    except (KeyError, TypeError):
        raise _lib.MmdError("semantic_search: score_function must be this module's cos_sim or dot_score "
                            "(arbitrary Python score functions would need a CPU path; there is none)") from None
    if device is None and isinstance(corpus_embeddings, torch.Tensor) and corpus_embeddings.is_cuda:
        device = corpus_embeddings.device
    pc = _prepared_for(corpus_embeddings, dtype, metric, device)
    scores, idx = ops.topk(query_embeddings, pc, top_k, dense_fallback=True)     # top_k > 120: dense pass + device sort
    s_host = scores.cpu().tolist()
    i_host = idx.cpu().tolist()
    return [[{"corpus_id": int(i), "score": float(s)} for s, i in zip(srow, irow) if i >= 0]
            for srow, irow in zip(s_host, i_host)]
