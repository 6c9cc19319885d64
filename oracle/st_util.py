"""Restatement of the third-party arithmetic behind the reference's text retrieval.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference calls `sentence_transformers.util.semantic_search` (pinned
sentence-transformers==3.3.1, /root/reference/requirements.txt:15) at
  src/evidence/text2text_retrieval.py:56-58, 61-63   and   src/evidence/experiment_text.py:25-27, 30-32
but the package is neither vendored under /root/reference nor installed in this image (no network), and
the reference has no tests or golden vectors for this boundary.  What follows restates the published
upstream algorithm:

  normalize_embeddings(x) = F.normalize(x, p=2, dim=1)            -> x / max(||x||_2, 1e-12)
  cos_sim(a, b)           = mm(normalize(a), normalize(b).T)       (1-D inputs are unsqueezed)
  dot_score(a, b)         = mm(a, b.T)
  semantic_search(q, c, query_chunk_size=100, corpus_chunk_size=500000, top_k=10, score_function=cos_sim):
      ndarray / list inputs become tensors; 1-D queries are unsqueezed; queries move to the corpus device;
      for every (query chunk, corpus chunk) block: scores -> torch.topk(min(top_k, block width), sorted=False)
      -> python lists -> per-query min-heap of (score, corpus_id) capped at top_k (push, then pushpop);
      finally each heap becomes a list of {"corpus_id", "score"} sorted by score descending.

Outputs follow the input dtype (fp16 in the reference, text2text_retrieval.py:44,53).  Tie order among equal
scores is whatever topk/heap/sort leave: unspecified upstream, so tests compare through oracle/exact.py's
near-tie classifier rather than position by position.
"""
from __future__ import annotations

import heapq
from typing import Callable, Dict, List, Union

import numpy as np
import torch


def _as_batch(x) -> torch.Tensor:
    if isinstance(x, (np.ndarray, np.generic)):
        x = torch.from_numpy(np.asarray(x))
    elif isinstance(x, (list, tuple)):
        x = torch.stack([torch.as_tensor(v) for v in x])
    elif not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    return x.unsqueeze(0) if x.dim() == 1 else x


def normalize_embeddings(embeddings: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.normalize(embeddings, p=2, dim=1)


def cos_sim(a, b) -> torch.Tensor:
    a, b = _as_batch(a), _as_batch(b)
    return torch.mm(normalize_embeddings(a), normalize_embeddings(b).transpose(0, 1))


def dot_score(a, b) -> torch.Tensor:
    a, b = _as_batch(a), _as_batch(b)
    return torch.mm(a, b.transpose(0, 1))


def semantic_search(query_embeddings, corpus_embeddings, query_chunk_size: int = 100, corpus_chunk_size: int = 500000,
                    top_k: int = 10, score_function: Callable = cos_sim) -> List[List[Dict[str, Union[int, float]]]]:
    queries = _as_batch(query_embeddings)
    corpus = _as_batch(corpus_embeddings)
    if corpus.device != queries.device:
        queries = queries.to(corpus.device)
    n_q, n_c = queries.shape[0], corpus.shape[0]
    heaps: List[list] = [[] for _ in range(n_q)]
    for q0 in range(0, n_q, query_chunk_size):
        q_block = queries[q0:q0 + query_chunk_size]
        for c0 in range(0, n_c, corpus_chunk_size):
            block = score_function(q_block, corpus[c0:c0 + corpus_chunk_size])
            vals, cols = torch.topk(block, min(top_k, block.shape[1]), dim=1, largest=True, sorted=False)
            vals, cols = vals.cpu().tolist(), cols.cpu().tolist()
            for r in range(len(vals)):
                heap = heaps[q0 + r]
                for col, score in zip(cols[r], vals[r]):
                    entry = (score, c0 + col)
                    if len(heap) < top_k:
                        heapq.heappush(heap, entry)
                    else:
                        heapq.heappushpop(heap, entry)
    out = []
    for heap in heaps:
        hits = [{"corpus_id": cid, "score": score} for score, cid in heap]
        out.append(sorted(hits, key=lambda h: h["score"], reverse=True))
    return out
