"""TEST INFRASTRUCTURE ONLY -- CPU oracle for joint image+text score fusion (BASELINE.json configs[3]).

The reference has no fused score: text evidence (src/evidence/text2text_retrieval.py:49-120) and image evidence
(src/evidence/im2im_retrieval.py:80-106) are retrieved separately, and the only merge it knows is the concatenation
+ descending sort of two hit lists (text2text_retrieval.py:97-110).  This file therefore states the fusion the
B200 path implements -- fused(q, c) = sum_m w_m * cos(q_m, c_m), every cos computed exactly like the
single-modality oracles (F.normalize semantics, per-norm clamp eps) -- in float64, so that the CUDA path can be
checked against it.  PARITY UNPINNED: no reference output exists for this configuration.
Nothing outside tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import exact


def fused_scores(queries: Sequence[torch.Tensor], corpora: Sequence[torch.Tensor], weights: Sequence[float],
                 metric: str = "cos", eps: float = 1e-12) -> torch.Tensor:
    """[Q, N] float64 matrix of sum_m w_m * score_m."""
    total = None
    for q, c, w in zip(queries, corpora, weights):
        s = exact.exact_scores(q, c, metric, eps) * float(w)
        total = s if total is None else total + s
    return total


def fused_topk(queries, corpora, weights, k: int, metric: str = "cos", eps: float = 1e-12) -> Tuple[torch.Tensor, torch.Tensor]:
    """(scores float64 [Q,k'], rows int64 [Q,k']) ordered by (score descending, row ascending)."""
    full = fused_scores(queries, corpora, weights, metric, eps)
    k = min(k, full.shape[1])
    order = torch.argsort(full, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(full, 1, order), order


def concat_and_sort(hits_a, hits_b, top_k: int):
    """The reference's own merge of two hit lists: concatenate, sort by score descending, keep the first of every
    distinct score until top_k are kept (src/evidence/text2text_retrieval.py:97-118)."""
    merged = sorted(list(hits_a) + list(hits_b), key=lambda t: t[1], reverse=True)
    seen, out = set(), []
    for key, score in merged:
        if score not in seen:
            seen.add(score)
            out.append((key, score))
        if len(out) == top_k:
            break
    return out
