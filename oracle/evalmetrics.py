"""Restatement of the reference's retrieval evaluation tails.  TEST INFRASTRUCTURE ONLY.

  dedupe_first_of_each_score   src/evidence/im2im_retrieval.py:94-104, src/evidence/text2text_retrieval.py:105-118
  ... with gold exemption      src/evidence/experiment_image.py:41-50,  src/evidence/experiment_text.py:79-87
  hits_at_k                    src/evidence/experiment_image.py:52-61,  src/evidence/experiment_text.py:89-104

The scripts themselves cannot be imported (experiment_image.py:5 uses a non-package import and Windows
path separators, experiment_text.py needs sentence_transformers), so the logic is restated on (key, score)
lists and exercised on synthetic planted-positive data.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple


def dedupe_first_of_each_score(ranked: Sequence[Tuple[object, float]], top_k: int,
                               gold: Optional[Callable[[object], bool]] = None) -> List[Tuple[object, float]]:
    unique_scores, kept = set(), []
    for key, score in ranked:
        if score not in unique_scores or (gold is not None and gold(key)):
            unique_scores.add(score)
            kept.append((key, score))
        if len(kept) == top_k:
            break
    return kept


def hits_at_k(per_query_keys: Sequence[Sequence[object]], gold_keys: Sequence[object],
              k_values: Sequence[int] = (1, 2, 5, 10)) -> Dict[int, float]:
    hits = {k: 0 for k in k_values}
    for keys, gold in zip(per_query_keys, gold_keys):
        for k in k_values:
            if gold in list(keys)[:k]:
                hits[k] += 1
    return {k: hits[k] / max(len(gold_keys), 1) for k in k_values}


def image_eval(scores_full, corpus_keys: Sequence[object], gold_keys: Sequence[object],
               k_values: Sequence[int] = (1, 2, 5, 10)) -> Dict[int, float]:
    """experiment_image.py:12-61 on a precomputed [Q,N] score matrix: full sort, dedupe with gold exemption, hits@k."""
    import torch
    top_k = max(k_values)
    order = torch.sort(scores_full, dim=1, descending=True, stable=True)
    vals, idx = order.values.tolist(), order.indices.tolist()
    lists = []
    for q in range(len(vals)):
        ranked = [(corpus_keys[i], float(s)) for s, i in zip(vals[q], idx[q])]
        kept = dedupe_first_of_each_score(ranked, top_k, gold=lambda key, q=q: key == gold_keys[q])
        lists.append([k for k, _ in kept])
    return hits_at_k(lists, gold_keys, k_values)
