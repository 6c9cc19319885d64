"""Joint image+text evidence retrieval with top-K score fusion (BASELINE.json configs[3]).

    score(claim, evidence) = sum_m w_m * cos(q_m, c_m)        m over modalities (text, image, ...)

The reference retrieves text and image evidence separately (src/evidence/text2text_retrieval.py:49-120,
src/evidence/im2im_retrieval.py:80-106) and its closest thing to a fusion is the concatenation + sort of two hit
lists (text2text_retrieval.py:97-118); there is no reference implementation of a fused score.  Here the
modalities of one corpus row are laid side by side along the contraction axis (each segment L2-normalised by K1 on
its own, the query segments additionally scaled by w_m), so that ONE pass of the fused tensor-core top-K kernel
ranks by the weighted sum, and the exact fp32 re-score recomputes the weighted sum from the original embeddings.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib, ops


@dataclass
class JointCorpus:
    """Prepared multi-modality corpus: rows = [segment_0 | segment_1 | ...] operand tiles resident in HBM."""
    rows: torch.Tensor                        # uint8 [N, row_bytes]
    inv_norms: List[Optional[torch.Tensor]]   # per modality f32 [N] (None for metric="dot")
    sources: List[Optional[torch.Tensor]]     # per modality caller embeddings on the device (exact re-score)
    dims: List[int]                           # per modality embedding width
    seg_bytes: List[int]                      # per modality bytes of one prepared segment
    weights: List[float]
    n: int
    op: str
    metric: str
    eps: float
    idx_offset: int = 0

    @property
    def device(self) -> torch.device:
        return self.rows.device

    @property
    def dim(self) -> int:
        """Contraction length in operand elements (what mmd_topk_scores takes as `dim`)."""
        return sum(self.seg_bytes) // (1 if self.op == "fp8" else 2)

    @property
    def source(self):
        return None if any(s is None for s in self.sources) else self.sources


def _segment_layout(op: str, dims: Sequence[int]) -> Tuple[List[int], int]:
    seg = [ops.prepared_layout(op, d)[1] for d in dims]
    return seg, sum(seg)


def _cast_segments(mats: Sequence[torch.Tensor], op: str, side: int, normalize: bool, eps: float, scales: Sequence[float],
                   seg_bytes: Sequence[int], row_bytes: int, out=None):
    """out = (rows buffer uint8 [>= rows, row_bytes], [inv_norm buffer f32 [>= rows] per modality]): caller-owned
    destinations, no allocation."""
    rows = mats[0].shape[0]
    dev = mats[0].device
    inv_bufs = None
    if out is None:
        out = torch.empty((rows, row_bytes), dtype=torch.uint8, device=dev)
    else:
        out, inv_bufs = out
        assert out.shape[1] == row_bytes and out.shape[0] >= rows and out.is_contiguous()
    invs = []
    lib = _lib.load()
    off = 0
    for m, (x, sb, sc) in enumerate(zip(mats, seg_bytes, scales)):
        inv = torch.empty((rows,), dtype=torch.float32, device=dev) if inv_bufs is None else inv_bufs[m]
        if rows:
            with torch.cuda.device(dev):
                rc = lib.mmd_normalize_cast_segment(ops._ptr(x), ops._SRC_DTYPE[x.dtype], rows, x.shape[1], x.stride(0), int(normalize),
                                                    float(eps), float(sc), ops._OP_DTYPE[op], side, C.c_void_p(out.data_ptr() + off),
                                                    row_bytes, ops._ptr(inv), ops._stream_ptr(dev))
            _lib.check(rc, "mmd_normalize_cast_segment")
        invs.append(inv)
        off += sb
    return out, invs


def prepare_joint(corpora: Sequence, weights: Optional[Sequence[float]] = None, dtype: str = "bf16", metric: str = "cos",
                  eps: float = ops.DEFAULT_EPS, keep_source: bool = True, device=None, idx_offset: int = 0) -> JointCorpus:
    """corpora: one [N, D_m] embedding matrix per modality (same N).  weights default to 1/len(corpora)."""
    if dtype not in ("bf16", "fp16", "fp8"):
        raise ValueError("joint retrieval supports dtype bf16, fp16 or fp8")
    if metric not in ops.METRICS:
        raise ValueError(f"metric must be one of {ops.METRICS}")
    if not 1 <= len(corpora) <= 4:
        raise ValueError("1 to 4 modalities")
    w = [1.0 / len(corpora)] * len(corpora) if weights is None else [float(x) for x in weights]
    if len(w) != len(corpora) or any(x < 0 for x in w):
        raise ValueError("one non-negative weight per modality")
    if dtype == "fp8":
        # e4m3 operands carry a fixed 2^8 scale sized for |x| <= 1 (unit-norm rows; the query side is also multiplied by w_m):
        # 256 * 1.75 = 448 is the largest finite e4m3 value
        if metric != "cos":
            raise ValueError("dtype='fp8' needs metric='cos' (unit-norm operands)")
        if any(x > 1.75 for x in w):
            raise ValueError("dtype='fp8' needs modality weights <= 1.75 (the scaled query operands would saturate e4m3)")
    if device is not None:
        dev = torch.device(device)
    elif isinstance(corpora[0], torch.Tensor) and corpora[0].is_cuda:
        dev = corpora[0].device
    elif torch.cuda.is_available():
        dev = torch.device("cuda", torch.cuda.current_device())
    else:
        raise _lib.MmdError("no CUDA device is available; the retrieval path has no CPU fallback")
    ops._require_cuda(dev)
    mats = [ops._as_rows(c, dev) for c in corpora]
    n = mats[0].shape[0]
    if any(m.shape[0] != n for m in mats):
        raise ValueError("every modality must have the same number of corpus rows")
    dims = [m.shape[1] for m in mats]
    seg, row_bytes = _segment_layout(dtype, dims)
    rows, invs = _cast_segments(mats, dtype, _lib.SIDE_CORPUS, metric == "cos", eps, [1.0] * len(mats), seg, row_bytes)
    return JointCorpus(rows=rows, inv_norms=invs if metric == "cos" else [None] * len(mats),
                       sources=mats if keep_source else [None] * len(mats), dims=dims, seg_bytes=seg, weights=w, n=n, op=dtype,
                       metric=metric, eps=eps, idx_offset=idx_offset)


def _rescore_joint(qs: Sequence[torch.Tensor], q_invs, jc: JointCorpus, cand: torch.Tensor, k_out: int):
    lib = _lib.load()
    dev = jc.device
    n_seg = len(qs)
    n_queries, k_in = cand.shape
    scores = torch.empty((n_queries, k_out), dtype=torch.float32, device=dev)
    idx = torch.empty((n_queries, k_out), dtype=torch.int32, device=dev)
    if n_queries == 0:
        return scores, idx
    vp = C.c_void_p * n_seg
    ip = C.c_int * n_seg
    lp = C.c_int64 * n_seg
    fp = C.c_float * n_seg
    null_or = lambda t: C.c_void_p(0 if t is None else t.data_ptr())  # noqa: E731
    with torch.cuda.device(dev):
        rc = lib.mmd_rescore_joint(
            n_seg, vp(*[null_or(q) for q in qs]), ip(*[ops._SRC_DTYPE[q.dtype] for q in qs]), lp(*[q.stride(0) for q in qs]),
            vp(*[null_or(t) for t in q_invs]), vp(*[null_or(s) for s in jc.sources]),
            ip(*[ops._SRC_DTYPE[s.dtype] for s in jc.sources]), lp(*[s.stride(0) if s.shape[0] else d for s, d in zip(jc.sources, jc.dims)]),
            vp(*[null_or(t) for t in jc.inv_norms]), ip(*jc.dims), fp(*jc.weights), n_queries, jc.n, ops._ptr(cand), k_in,
            jc.idx_offset, k_out, ops._ptr(scores), ops._ptr(idx), ops._stream_ptr(dev))
    _lib.check(rc, "mmd_rescore_joint")
    return scores, idx


def topk_joint(queries: Sequence, corpus: JointCorpus, k: int, rescore_exact: Optional[bool] = None,
               overfetch: Optional[int] = None, index_dtype: torch.dtype = torch.int64) -> Tuple[torch.Tensor, torch.Tensor]:
    """queries: one [Q, D_m] matrix per modality.  Returns (fused scores f32 [Q,k'], corpus rows [Q,k']),
    k' = min(k, N), ordered by (fused score descending, row ascending)."""
    if k <= 0:
        raise ValueError("k must be positive")
    jc = corpus
    if len(queries) != len(jc.dims):
        raise ValueError(f"expected {len(jc.dims)} query modalities, got {len(queries)}")
    qs = [ops._as_rows(q, jc.device) for q in queries]
    n_queries = qs[0].shape[0]
    for q, d in zip(qs, jc.dims):
        if q.shape[1] != d:
            raise RuntimeError(f"query dim {q.shape[1]} does not match corpus dim {d}")
        if q.shape[0] != n_queries:
            raise ValueError("every modality must have the same number of queries")
    k_eff = min(k, jc.n)
    if k_eff == 0:
        return (torch.empty((n_queries, 0), dtype=torch.float32, device=jc.device),
                torch.empty((n_queries, 0), dtype=index_dtype, device=jc.device))
    if k_eff > ops.max_k():
        raise _lib.MmdError(f"k={k_eff} exceeds the fused top-k limit {ops.max_k()}")
    do_rescore = (jc.source is not None) if rescore_exact is None else bool(rescore_exact)
    if do_rescore and jc.source is None:
        raise ValueError("rescore_exact=True needs a JointCorpus built with keep_source=True")
    kprime = k_eff
    if do_rescore:
        kprime = ops.overfetch_for(k_eff, jc.n) if overfetch is None else max(k_eff, min(int(overfetch), jc.n, ops.max_k()))
    q_rows, q_invs = _cast_segments(qs, jc.op, _lib.SIDE_QUERY, jc.metric == "cos", jc.eps, jc.weights, jc.seg_bytes,
                                    sum(jc.seg_bytes))
    scores, idx = ops.topk_prepared(q_rows, n_queries, jc, kprime)
    if do_rescore:
        scores, idx = _rescore_joint(qs, q_invs if jc.metric == "cos" else [None] * len(qs), jc, idx, k_eff)
    if index_dtype != torch.int32:
        idx = idx.to(index_dtype)
    return scores, idx
