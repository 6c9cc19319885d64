"""Host-side tails of the reference's retrieval entry points: distinct-score dedupe and hits@k.

These are a few comparisons per query on the already-reduced top-K' list, so they stay in Python exactly
as in the reference; the heavy part (scoring + top-K') is the CUDA path in ops.py.

  dedupe_by_score            src/evidence/im2im_retrieval.py:94-104, src/evidence/text2text_retrieval.py:105-118
  dedupe_by_score(gold=...)  src/evidence/experiment_image.py:41-50, src/evidence/experiment_text.py:79-87
  hits_at_k                  src/evidence/experiment_image.py:52-61, src/evidence/experiment_text.py:89-104
"""
from __future__ import annotations

from typing import Callable, Dict, Hashable, Iterable, List, Optional, Sequence, Tuple


def dedupe_by_score(ranked: Iterable[Tuple[Hashable, float]], top_k: int,
                    is_gold: Optional[Callable[[Hashable], bool]] = None) -> List[Tuple[Hashable, float]]:
    """Walk a (key, score) list sorted by score descending; keep the first entry of every distinct score
    (and, in the eval variant, any entry for which is_gold(key) holds) until top_k entries are kept."""
    seen = set()
    kept: List[Tuple[Hashable, float]] = []
    if top_k <= 0:
        return kept
    for key, score in ranked:
        if (score not in seen) or (is_gold is not None and is_gold(key)):
            seen.add(score)
            kept.append((key, score))
        if len(kept) == top_k:
            break
    return kept


def hits_at_k(retrieved: Sequence[Sequence[Hashable]], gold: Sequence[Hashable],
              k_values: Sequence[int] = (1, 2, 5, 10)) -> Dict[int, float]:
    """Fraction of queries whose gold key is among their first k retrieved keys, for every k."""
    hits = {k: 0 for k in k_values}
    for keys, g in zip(retrieved, gold):
        for k in k_values:
            if g in list(keys)[:k]:
                hits[k] += 1
    n = max(len(gold), 1)
    return {k: v / n for k, v in hits.items()}
