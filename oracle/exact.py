"""Batched ground truth and the near-tie classifier.  TEST INFRASTRUCTURE ONLY.

exact_scores / exact_topk compute the path's contract -- L2-normalise (x / max(||x||, eps)), score every
query against every corpus row, ordered top-k -- in float64 (or any torch dtype), with the path's tie rule
(score descending, corpus row ascending).  They restate, in batched form,
  * sentence_transformers.util.cos_sim + topk (call sites src/evidence/text2text_retrieval.py:56-64), eps = 1e-12
  * nn.CosineSimilarity(dim=1, eps=1e-6) per pair + sorted()   (src/evidence/im2im_retrieval.py:38-42, 84-92)
and are themselves checked against oracle/st_util.py and oracle/im2im.py (tests/test_oracle.py).

compare_topk implements the parity rule of BASELINE.json: "top-K index sets must match exactly except at
score near-ties": a row of the device result differs legitimately from the oracle only if every row it
swapped in / out has an oracle score within `tie_tol` of the oracle's K-th score.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch


def normalize_rows(x: torch.Tensor, eps: float, dtype=torch.float64) -> torch.Tensor:
    x = x.to(dtype)
    return x / x.norm(dim=1, keepdim=True).clamp_min(eps)


def round_operand(x: torch.Tensor, op: Optional[str]) -> torch.Tensor:
    """Round normalised fp32 rows the way the device's operand cast does (None = keep)."""
    if op is None or op == "fp32":
        return x
    if op == "bf16":
        return x.float().to(torch.bfloat16).to(x.dtype)
    if op == "fp16":
        return x.float().to(torch.float16).to(x.dtype)
    if op == "fp8":
        return ((x.float() * 256.0).to(torch.float8_e4m3fn).to(x.dtype)) / 256.0
    raise ValueError(op)


def exact_scores(q: torch.Tensor, c: torch.Tensor, metric: str = "cos", eps: float = 1e-12, dtype=torch.float64,
                 operand: Optional[str] = None) -> torch.Tensor:
    """[Q,N] scores.  operand="bf16"/"fp16"/"fp8": normalise in fp32 like the device, round the operands,
    then take the dot products in `dtype` -- the value the tensor-core pass computes up to accumulation order."""
    if metric == "cos":
        if operand is None:
            qn, cn = normalize_rows(q, eps, dtype), normalize_rows(c, eps, dtype)
        else:
            qn = round_operand(normalize_rows(q, eps, torch.float32), operand).to(dtype)
            cn = round_operand(normalize_rows(c, eps, torch.float32), operand).to(dtype)
    elif metric == "dot":
        qn, cn = round_operand(q.float(), operand).to(dtype), round_operand(c.float(), operand).to(dtype)
    else:
        raise ValueError(metric)
    return qn @ cn.T


def ordered_topk(scores: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k of each row by (score descending, column ascending)."""
    k = min(k, scores.shape[1])
    # stable sort on descending score keeps ascending column order among equals
    vals, idx = torch.sort(scores, dim=1, descending=True, stable=True)
    return vals[:, :k].contiguous(), idx[:, :k].contiguous()


def exact_topk(q, c, k: int, metric: str = "cos", eps: float = 1e-12, dtype=torch.float64, operand: Optional[str] = None):
    return ordered_topk(exact_scores(q, c, metric, eps, dtype, operand), k)


@dataclass
class TopkComparison:
    rows: int
    identical_sets: int          # rows whose index SET equals the oracle's
    identical_order: int         # rows whose index LIST equals the oracle's
    excused_rows: int            # rows that differ only by near-ties
    violations: int              # rows with a genuine mismatch
    max_score_err: float         # max |device score - oracle score of the same (row, index)|
    max_rel_score_err: float

    @property
    def ok(self) -> bool:
        return self.violations == 0


def compare_topk(dev_scores: torch.Tensor, dev_idx: torch.Tensor, oracle_full: torch.Tensor, k: int,
                 tie_tol: float) -> TopkComparison:
    """dev_* : [Q,k] device result; oracle_full : [Q,N] oracle scores (float64)."""
    dev_scores, dev_idx = dev_scores.cpu().double(), dev_idx.cpu().long()
    n_q, n = oracle_full.shape
    k = min(k, n)
    o_vals, o_idx = ordered_topk(oracle_full, k)
    same_set = same_order = excused = bad = 0
    kth = o_vals[:, k - 1]
    for r in range(n_q):
        d, o = dev_idx[r, :k].tolist(), o_idx[r].tolist()
        if d == o:
            same_order += 1
            same_set += 1
            continue
        ds, os_ = set(d), set(o)
        if len(ds) != k or min(d) < 0 or max(d) >= n:
            bad += 1
            continue
        if ds == os_:
            same_set += 1
            # order differs: every adjacent inversion must be between near-equal oracle scores
            sc = oracle_full[r, d]
            if bool(((sc[1:] - sc[:-1]) > tie_tol).any()):
                bad += 1
            else:
                excused += 1
            continue
        swapped = list(ds ^ os_)
        if bool(((oracle_full[r, swapped] - kth[r]).abs() <= tie_tol).all()):
            excused += 1
        else:
            bad += 1
    picked = torch.gather(oracle_full, 1, dev_idx[:, :k].clamp(0, n - 1))
    err = (dev_scores[:, :k] - picked).abs()
    rel = err / picked.abs().clamp_min(1e-30)
    return TopkComparison(n_q, same_set, same_order, excused, bad, float(err.max()) if err.numel() else 0.0,
                          float(rel.max()) if rel.numel() else 0.0)


def recall_at_k(dev_idx: torch.Tensor, oracle_full: torch.Tensor, k: int) -> float:
    """Fraction of the oracle's top-k rows (by (score desc, row asc)) that the device lists contain -- the bar for the lossy
    fp8 candidate pass, for which BASELINE.json states no tolerance."""
    _, o_idx = ordered_topk(oracle_full, k)
    d = dev_idx.cpu().long()[:, :k].tolist()
    hit = sum(len(set(a) & set(b)) for a, b in zip(d, o_idx.tolist()))
    return hit / float(o_idx.numel())
