"""Raw tensor API of the retrieval hot path (boundary B3 of SURVEY.md section 8b).

    topk(queries[Q,D], corpus[N,D] | PreparedCorpus, k, metric, dtype, eps) -> (scores f32 [Q,k'], indices [Q,k'])

with k' = min(k, N), scores sorted descending, ties by ascending corpus row.  It replaces, for a whole batch
of queries at once, what the reference does one query at a time:

  * sentence_transformers.util.semantic_search (normalise -> mm -> topk -> heap merge), called from
    src/evidence/text2text_retrieval.py:56-64 and src/evidence/experiment_text.py:25-33;
  * the per-pair nn.CosineSimilarity loop + full sort of src/evidence/im2im_retrieval.py:84-92 and
    src/evidence/experiment_image.py:25-33.

PyTorch is used for device memory and streams only; all arithmetic happens in libmmd.so (CUDA, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import functools
import math
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple, Union

import torch

from . import _lib

_SRC_DTYPE = {torch.float32: _lib.SRC_F32, torch.float16: _lib.SRC_F16, torch.bfloat16: _lib.SRC_BF16}
_OP_DTYPE = {"bf16": _lib.OP_BF16, "fp16": _lib.OP_F16, "fp8": _lib.OP_E4M3, "fp32": _lib.OP_BF16X3}
METRICS = ("cos", "dot")

#: default clamp of the norm: F.normalize's 1e-12 (text path); the image path passes nn.CosineSimilarity's 1e-6
DEFAULT_EPS = 1e-12


def _stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(device: torch.device) -> None:
    if device.type != "cuda":
        raise _lib.MmdError(f"the retrieval path runs on a CUDA (sm_100) device only, got {device}; no CPU fallback")


def _as_rows(x, device: Optional[torch.device] = None) -> torch.Tensor:
    """list / ndarray / tensor -> 2-D tensor with unit inner stride on `device` (fp32/fp16/bf16)."""
    if not isinstance(x, torch.Tensor):
        import numpy as np
        arr = np.asarray(x)
        if not arr.flags.writeable:          # e.g. a read-only memory map: torch needs a writable buffer
            arr = arr.copy()
        x = torch.as_tensor(arr)
    if x.dim() == 1:
        x = x.unsqueeze(0)
    if x.dim() != 2:
        raise ValueError(f"expected [rows, dim] embeddings, got shape {tuple(x.shape)}")
    if x.dtype not in _SRC_DTYPE:
        x = x.to(torch.float32)
    if device is not None and x.device != device:
        x = x.to(device, non_blocking=True)
    if x.stride(1) != 1 and x.numel() > 0:
        x = x.contiguous()
    return x


@functools.lru_cache(maxsize=None)
def prepared_layout(op: str, dim: int) -> Tuple[int, int]:
    """(contraction length in operand elements, bytes per prepared row)."""
    kdim, row_bytes = C.c_int64(0), C.c_int64(0)
    _lib.check(_lib.load().mmd_prepared_layout(_OP_DTYPE[op], dim, C.byref(kdim), C.byref(row_bytes)), "mmd_prepared_layout")
    return kdim.value, row_bytes.value


def normalize_cast(x: torch.Tensor, op: str, side: int, normalize: bool, eps: float,
                   out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """K1: rows of `x` -> (prepared operand rows uint8 [rows, row_bytes], inv_norm f32 [rows]).
    out = (rows buffer, inv_norm buffer): caller-owned destinations (first `rows` rows are written), no allocation."""
    _require_cuda(x.device)
    rows, dim = x.shape
    _, row_bytes = prepared_layout(op, dim)
    if out is None:
        out = torch.empty((rows, row_bytes), dtype=torch.uint8, device=x.device)
        inv = torch.empty((rows,), dtype=torch.float32, device=x.device)
    else:
        out, inv = out
        assert out.dtype == torch.uint8 and out.shape[0] >= rows and out.shape[1] == row_bytes and out.is_contiguous()
        assert inv.dtype == torch.float32 and inv.shape[0] >= rows
    if rows:
        lib = _lib.load()
        with torch.cuda.device(x.device):
            rc = lib.mmd_normalize_cast(_ptr(x), _SRC_DTYPE[x.dtype], rows, dim, x.stride(0), int(normalize), float(eps),
                                        _OP_DTYPE[op], side, _ptr(out), _ptr(inv), _stream_ptr(x.device))
        _lib.check(rc, "mmd_normalize_cast")
    return out, inv


@dataclass
class PreparedCorpus:
    """Corpus rows normalised and cast once, resident in HBM, ready for any number of query batches.

    (The reference re-normalises the whole corpus inside every semantic_search call,
    src/evidence/text2text_retrieval.py:56-63.)"""
    rows: torch.Tensor            # uint8 [N, row_bytes] operand tiles
    inv_norm: Optional[torch.Tensor]   # f32 [N] (None for metric="dot")
    source: Optional[torch.Tensor]     # caller's embeddings on the device (for the exact re-score), or None
    n: int
    dim: int
    op: str
    metric: str
    eps: float
    idx_offset: int = 0           # global row of local row 0 (row-sharded corpora)

    @property
    def device(self) -> torch.device:
        return self.rows.device


def prepare_corpus(corpus, dtype: str = "bf16", metric: str = "cos", eps: float = DEFAULT_EPS, keep_source: bool = True,
                   device: Optional[Union[str, torch.device]] = None, idx_offset: int = 0) -> PreparedCorpus:
    if dtype not in _OP_DTYPE:
        raise ValueError(f"dtype must be one of {sorted(_OP_DTYPE)}, got {dtype!r}")
    if metric not in METRICS:
        raise ValueError(f"metric must be one of {METRICS}, got {metric!r}")
    if dtype == "fp8" and metric == "dot":
        # the e4m3 operands carry a fixed 2^8 scale that assumes unit-norm rows (|x| <= 1); un-normalised embeddings would
        # saturate at 448 or underflow and the candidate selection would silently be garbage
        raise ValueError("dtype='fp8' needs metric='cos' (unit-norm operands); use bf16 / fp16 / fp32 for inner-product scoring")
    if device is not None:
        dev = torch.device(device)
    elif isinstance(corpus, torch.Tensor) and corpus.is_cuda:
        dev = corpus.device
    elif torch.cuda.is_available():
        dev = torch.device("cuda", torch.cuda.current_device())
    else:
        raise _lib.MmdError("no CUDA device is available; the retrieval path has no CPU fallback")
    _require_cuda(dev)
    c = _as_rows(corpus, dev)
    rows, inv = normalize_cast(c, dtype, _lib.SIDE_CORPUS, metric == "cos", eps)
    return PreparedCorpus(rows=rows, inv_norm=inv if metric == "cos" else None, source=c if keep_source else None,
                          n=c.shape[0], dim=c.shape[1], op=dtype, metric=metric, eps=eps, idx_offset=idx_offset)


@functools.lru_cache(maxsize=None)
def max_k() -> int:
    return int(_lib.load().mmd_topk_max_k())


def overfetch_for(k: int, n: int) -> int:
    """Candidates kept by the low-precision pass when an exact re-score follows: k + max(8, k/2), but never so
    many that the per-row candidate buffer of the fused kernel (128 entries for lists longer than 32) is left
    with fewer than 24 free slots between two compactions.

    CAVEAT (k > 69): the cap of 104 shrinks the over-fetch -- at the reference's largest list, k = 100
    (experiment_text.py:26), only 4 spare candidates are re-scored.  bf16 operand noise is ~1e-4 absolute; on corpora large
    enough that the rank-k score spacing falls below that (~1e-4 at rank 100 of a 1M-row Gaussian corpus; the reference's
    own corpora hold 7.5k / 35k rows, spacing ~1e-3) a true top-k row whose bf16 score ranks below K' is never re-scored.
    tests/test_gpu_fullsize.py states the measured recall@100 on 1M rows; dtype="fp32" (6-term bf16 split) is exact."""
    return max(1, min(n, max(k, min(k + max(8, k // 2), 104)), max_k()))


def topk_prepared(q_rows: torch.Tensor, n_queries: int, corpus: PreparedCorpus, k: int,
                  shared_thr: Optional[Tuple[int, Sequence[int]]] = None,
                  pair_dst: Optional[Tuple[Sequence[int], int]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """K2+K3 (+K3b): fused contraction + top-k over prepared operands -> (scores f32 [Q,k], idx i32 [Q,k]).
    shared_thr = (device pointer of this rank's threshold array, device pointers of every rank's array): row-sharded
    corpora share their pruning thresholds across GPUs (mmd_topk_scores_shared); pair_dst = (device pointers, pair offset):
    the strip merge also stores the packed list into those buffers (every rank's gather buffer)."""
    lib = _lib.load()
    dev = corpus.device
    scores = torch.empty((n_queries, k), dtype=torch.float32, device=dev)
    idx = torch.empty((n_queries, k), dtype=torch.int32, device=dev)
    if n_queries == 0:
        return scores, idx
    op = _OP_DTYPE[corpus.op]
    with torch.cuda.device(dev):
        # (inside the device guard: the schedule the workspace is sized for depends on the SM count of the device that runs it)
        ws_bytes = int(lib.mmd_topk_workspace_bytes(n_queries, max(corpus.n, 1), corpus.dim, op, k))
        ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=dev)
        if shared_thr is None:
            rc = lib.mmd_topk_scores(_ptr(q_rows), _ptr(corpus.rows), op, n_queries, corpus.n, corpus.dim, k, corpus.idx_offset,
                                     _ptr(scores), _ptr(idx), _ptr(ws), ws_bytes, _stream_ptr(dev))
        else:
            local, everyone = shared_thr
            arr = (C.c_void_p * len(everyone))(*[C.c_void_p(int(p)) for p in everyone])
            dsts, off = pair_dst if pair_dst is not None else ((), 0)
            darr = (C.c_void_p * max(len(dsts), 1))(*[C.c_void_p(int(p)) for p in dsts])
            rc = lib.mmd_topk_scores_shared(_ptr(q_rows), _ptr(corpus.rows), op, n_queries, corpus.n, corpus.dim, k, corpus.idx_offset,
                                            _ptr(scores), _ptr(idx), _ptr(ws), ws_bytes, C.c_void_p(int(local)), arr, len(everyone),
                                            darr, len(dsts), int(off), _stream_ptr(dev))
    _lib.check(rc, "mmd_topk_scores")
    return scores, idx


def rescore(q: torch.Tensor, q_inv: Optional[torch.Tensor], corpus: PreparedCorpus, cand_idx: torch.Tensor,
            k_out: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """K5: exact fp32 cosine / dot of the candidates from the original embeddings, re-ranked."""
    lib = _lib.load()
    dev = corpus.device
    n_queries, k_in = cand_idx.shape
    scores = torch.empty((n_queries, k_out), dtype=torch.float32, device=dev)
    idx = torch.empty((n_queries, k_out), dtype=torch.int32, device=dev)
    if n_queries == 0:
        return scores, idx
    src = corpus.source
    with torch.cuda.device(dev):
        rc = lib.mmd_rescore(_ptr(q), _SRC_DTYPE[q.dtype], q.stride(0), _ptr(q_inv), _ptr(src), _SRC_DTYPE[src.dtype],
                             src.stride(0) if src.shape[0] else corpus.dim, _ptr(corpus.inv_norm), n_queries, corpus.n, corpus.dim,
                             _ptr(cand_idx), k_in, corpus.idx_offset, k_out, _ptr(scores), _ptr(idx), _stream_ptr(dev))
    _lib.check(rc, "mmd_rescore")
    return scores, idx


def rescore_pairs(q: torch.Tensor, q_inv: Optional[torch.Tensor], corpus: PreparedCorpus, cand_idx: torch.Tensor, k_out: int,
                  dst_ptrs: Sequence[int], dst_offset_pairs: int = 0) -> None:
    """K5 with packed output: the k_out best re-scored candidates of every query go, as {score bits, row} int32
    pairs, to every device pointer in `dst_ptrs` at pair offset dst_offset_pairs + q * k_out (see mmd_rescore_pairs)."""
    lib = _lib.load()
    dev = corpus.device
    n_queries, k_in = cand_idx.shape
    if n_queries == 0:
        return
    src = corpus.source
    arr = (C.c_void_p * len(dst_ptrs))(*[C.c_void_p(int(p)) for p in dst_ptrs])
    with torch.cuda.device(dev):
        rc = lib.mmd_rescore_pairs(_ptr(q), _SRC_DTYPE[q.dtype], q.stride(0), _ptr(q_inv), _ptr(src), _SRC_DTYPE[src.dtype],
                                   src.stride(0) if src.shape[0] else corpus.dim, _ptr(corpus.inv_norm), n_queries, corpus.n,
                                   corpus.dim, _ptr(cand_idx), k_in, corpus.idx_offset, k_out, arr, len(dst_ptrs),
                                   int(dst_offset_pairs), _stream_ptr(dev))
    _lib.check(rc, "mmd_rescore_pairs")


def _ordered_bound(scores: torch.Tensor, raw_scale: float) -> torch.Tensor:
    """K-th best scores f32 [Q] -> the kernel's threshold encoding (order-preserving uint32 of the next float below, as
    int64): what publish_threshold() writes for a bound learnt on the device."""
    raw = (scores.float() * raw_scale + 0.0).contiguous()
    u = raw.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    ordered = torch.where(u >= 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000)
    ordered = torch.where(torch.isfinite(raw), ordered - 1, torch.zeros_like(ordered))      # -inf (short list): no bound
    return ordered


def topk_prepared_phased(q_rows: torch.Tensor, n_queries: int, corpus: PreparedCorpus, k: int, phases: int
                         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """topk_prepared over `phases` contiguous row ranges of the corpus, one fused launch each, with the pruning
    thresholds carried over AND tightened between launches: after every phase the lists found so far are merged and the
    merged K-th best becomes every query's bound for the next phase.  Inside one launch a CTA only ever learns the K-th
    best of single strips; the merged bound is an order of magnitude tighter on large corpora, which keeps the epilogue's
    collect pass rare -- it matters where the contraction is fast and the lists long (fp8, K = 100: BASELINE configs[4])."""
    dev = corpus.device
    phases = max(1, min(int(phases), max(1, corpus.n // 65536)))
    if phases == 1:
        return topk_prepared(q_rows, n_queries, corpus, k)
    raw_scale = 65536.0 if corpus.op == "fp8" else 1.0
    thr = torch.zeros((n_queries,), dtype=torch.int32, device=dev)
    thr_ptr = (thr.data_ptr(), [thr.data_ptr()])
    best_s = best_i = None
    for ph in range(phases):
        lo, hi = (corpus.n * ph) // phases, (corpus.n * (ph + 1)) // phases
        sub = PreparedCorpus(rows=corpus.rows[lo:hi], inv_norm=None, source=None, n=hi - lo, dim=corpus.dim, op=corpus.op,
                             metric=corpus.metric, eps=corpus.eps, idx_offset=corpus.idx_offset + lo)
        s, i = topk_prepared(q_rows, n_queries, sub, k, shared_thr=thr_ptr)
        if best_s is None:
            best_s, best_i = s, i
        else:
            best_s, best_i = merge_topk(torch.stack([best_s, s]), torch.stack([best_i, i]), k)
        if ph + 1 < phases:
            bound = _ordered_bound(best_s[:, k - 1], raw_scale)
            cur = thr.to(torch.int64) & 0xFFFFFFFF
            thr.copy_(torch.maximum(cur, bound).to(torch.int32))        # wraps into the same 32 bits
    return best_s, best_i


def topk_candidates(queries, pc: PreparedCorpus, k: int, overfetch: Optional[int] = None,
                    shared_thr: Optional[Tuple[int, Sequence[int]]] = None, pair_dst: Optional[Tuple[Sequence[int], int]] = None,
                    phases: int = 1):
    """First half of the default path: K1 on the queries + fused tensor-core top-K' (K' = over-fetched k).
    Returns (q rows on the device, q inv_norm or None, raw scores f32 [Q,K'], candidate rows i32 [Q,K'])."""
    q = _as_rows(queries, pc.device)
    if q.shape[1] != pc.dim:
        raise RuntimeError(f"query dim {q.shape[1]} does not match corpus dim {pc.dim}")
    k_eff = min(k, pc.n)
    kprime = overfetch_for(k_eff, pc.n) if overfetch is None else max(k_eff, min(int(overfetch), pc.n, max_k()))
    q_rows, q_inv = normalize_cast(q, pc.op, _lib.SIDE_QUERY, pc.metric == "cos", pc.eps)
    if phases > 1 and shared_thr is None and pc.n > 0:
        scores, idx = topk_prepared_phased(q_rows, q.shape[0], pc, max(kprime, 1), phases)
    else:
        scores, idx = topk_prepared(q_rows, q.shape[0], pc, max(kprime, 1), shared_thr=shared_thr if pc.n > 0 else None,
                                    pair_dst=pair_dst if (pc.n > 0 and shared_thr is not None) else None)
    return q, (q_inv if pc.metric == "cos" else None), scores, idx


# Over-fetch factor (candidates re-scored per result row, K'/k) a low-precision candidate pass needs before its exact
# re-score returns the fp32 top-k with recall ~1: the operand noise moves a score by sigma ~ 1.8e-3 (e4m3) / 1.4e-4 (bf16)
# on unit-norm 768-d embeddings, and the k-th best must stay inside the candidate list.  Lists are capped at 120 entries
# (104 with re-score), so long lists reach the factor by SPLITTING the corpus instead: see auto_splits().
_OVERFETCH_TARGET = {"fp8": 4.0, "bf16": 1.5, "fp16": 1.5}
_MIN_SPLIT_ROWS = 131072


def auto_splits(op: str, k: int, kprime: int, n_rows: int) -> int:
    """Independent sub-searches (contiguous row ranges, each with its own pruning bounds, its own K' candidates and its own
    exact re-score; the exact lists are merged) needed to bring the effective over-fetch V * K' / k up to the target of the
    operand type.  1 for the short lists of the default path (k = 10: K' = 18); 2 for bf16 at k = 100 (K' = 104); 3 / 4 for
    fp8 at k = 10 / 100.  A candidate pass can only lose a true top-k row whose low-precision score ranks below K' WITHIN
    ITS SPLIT, and a split holds only ~k/V of the true top-k (4 spare candidates at k = 100)."""
    target = _OVERFETCH_TARGET.get(op)
    if target is None or k <= 0 or kprime <= 0:
        return 1
    want = int(math.ceil(target * k / kprime - 0.05))
    return max(1, min(want, 8, n_rows // _MIN_SPLIT_ROWS))


def _slice_corpus(pc: PreparedCorpus, lo: int, hi: int) -> PreparedCorpus:
    return PreparedCorpus(rows=pc.rows[lo:hi], inv_norm=None if pc.inv_norm is None else pc.inv_norm[lo:hi],
                          source=None if pc.source is None else pc.source[lo:hi], n=hi - lo, dim=pc.dim, op=pc.op,
                          metric=pc.metric, eps=pc.eps, idx_offset=pc.idx_offset + lo)


def topk(queries, corpus, k: int, metric: str = "cos", dtype: str = "bf16", eps: Optional[float] = None,
         rescore_exact: Optional[bool] = None, overfetch: Optional[int] = None,
         index_dtype: torch.dtype = torch.int64, dense_fallback: bool = False, phases: int = 1,
         splits: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k most similar corpus rows for every query.

    queries : [Q,D] (or [D]) tensor / ndarray / list, any device (moved to the corpus device).
    corpus  : [N,D] embeddings (prepared on the fly) or a PreparedCorpus (prepared once, reused).
    returns : scores f32 [Q, min(k,N)] descending, indices [Q, min(k,N)] (rows of `corpus`, + idx_offset).

    rescore_exact (default: on when the corpus kept its source embeddings) over-fetches candidates with the
    tensor-core pass and recomputes their scores in fp32 from the original embeddings.
    phases > 1: the corpus is swept in that many launches with the merged K-th best carried between them as pruning bound
    (see topk_prepared_phased; for long lists over large corpora).
    splits (with the exact re-score; default auto_splits()): independent sub-searches over contiguous row ranges, each
    re-scored exactly, exact lists merged -- the effective over-fetch of long lists / fp8 candidates.
    dense_fallback: k beyond the fused selection's limit (120) is served by the dense tensor-core contraction plus a
    device-side stable sort of the score rows (query chunks of <= 1 GB of scores) -- off the hot path, for drop-in
    completeness only (the reference's largest list is top_k*10 = 100, experiment_text.py:26).
    """
    if k <= 0:
        raise ValueError("k must be positive")
    if isinstance(corpus, PreparedCorpus):
        pc = corpus
    else:
        pc = prepare_corpus(corpus, dtype=dtype, metric=metric, eps=DEFAULT_EPS if eps is None else eps)
    q = _as_rows(queries, pc.device)
    if q.shape[1] != pc.dim:
        raise RuntimeError(f"query dim {q.shape[1]} does not match corpus dim {pc.dim}")
    n_queries = q.shape[0]
    k_eff = min(k, pc.n)
    if k_eff == 0:
        return (torch.empty((n_queries, 0), dtype=torch.float32, device=pc.device),
                torch.empty((n_queries, 0), dtype=index_dtype, device=pc.device))
    if k_eff > max_k():
        if not dense_fallback:
            raise _lib.MmdError(f"k={k_eff} exceeds the fused top-k limit {max_k()} (pass dense_fallback=True for a ranked dense pass)")
        return _topk_dense(q, pc, k_eff, rescore_exact, index_dtype)
    do_rescore = (pc.source is not None) if rescore_exact is None else bool(rescore_exact)
    if do_rescore and pc.source is None:
        raise ValueError("rescore_exact=True needs a PreparedCorpus built with keep_source=True")
    if do_rescore:
        kprime = overfetch_for(k_eff, pc.n) if overfetch is None else max(k_eff, min(int(overfetch), pc.n, max_k()))
        n_splits = auto_splits(pc.op, k_eff, kprime, pc.n) if splits is None else max(1, min(int(splits), pc.n // max(k_eff, 1)))
        if n_splits > 1:
            parts_s, parts_i = [], []
            for v in range(n_splits):
                sub = _slice_corpus(pc, (pc.n * v) // n_splits, (pc.n * (v + 1)) // n_splits)
                qd, q_inv, _, cand = topk_candidates(q, sub, k_eff, overfetch, phases=phases)
                ps, pi = rescore(qd, q_inv, sub, cand, k_eff)
                parts_s.append(ps)
                parts_i.append(pi)
            scores, idx = merge_topk(torch.stack(parts_s), torch.stack(parts_i), k_eff)
        else:
            q, q_inv, _, cand = topk_candidates(q, pc, k_eff, overfetch, phases=phases)
            scores, idx = rescore(q, q_inv, pc, cand, k_eff)
    else:
        q_rows, _ = normalize_cast(q, pc.op, _lib.SIDE_QUERY, pc.metric == "cos", pc.eps)
        scores, idx = topk_prepared_phased(q_rows, n_queries, pc, k_eff, phases) if phases > 1 else \
            topk_prepared(q_rows, n_queries, pc, k_eff)
    if index_dtype != torch.int32:
        idx = idx.to(index_dtype)
    return scores, idx


def _topk_dense(q: torch.Tensor, pc: PreparedCorpus, k: int, rescore_exact: Optional[bool], index_dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """k > max_k(): dense scores (the same tensor-core contraction) + stable descending sort per query chunk; the k
    best are re-scored exactly when the corpus kept its source (lists of up to 1024)."""
    n_queries = q.shape[0]
    do_rescore = ((pc.source is not None) if rescore_exact is None else bool(rescore_exact)) and k <= 1024
    kprime = min(pc.n, k + max(8, k // 8), 1024) if do_rescore else k
    chunk = max(1, (1 << 28) // max(pc.n, 1))
    out_s = torch.empty((n_queries, k), dtype=torch.float32, device=pc.device)
    out_i = torch.empty((n_queries, k), dtype=torch.int32, device=pc.device)
    q_inv_all = None
    if do_rescore and pc.metric == "cos":
        _, q_inv_all = normalize_cast(q, pc.op, _lib.SIDE_QUERY, True, pc.eps)
    for lo in range(0, n_queries, chunk):
        sub = q[lo:lo + chunk]
        dense = dense_scores(sub, pc)
        s_sorted, i_sorted = torch.sort(dense, dim=1, descending=True, stable=True)
        s_top, i_top = s_sorted[:, :kprime].contiguous(), (i_sorted[:, :kprime] + pc.idx_offset).to(torch.int32).contiguous()
        if do_rescore:
            s_top, i_top = rescore(sub, None if q_inv_all is None else q_inv_all[lo:lo + chunk], pc, i_top, k)
        out_s[lo:lo + chunk], out_i[lo:lo + chunk] = s_top[:, :k], i_top[:, :k]
    return out_s, (out_i if index_dtype == torch.int32 else out_i.to(index_dtype))


def dense_scores(queries, corpus, metric: str = "cos", dtype: str = "bf16", eps: Optional[float] = None) -> torch.Tensor:
    """Full [Q,N] fp32 score matrix from the same tensor-core contraction (small shapes only)."""
    pc = corpus if isinstance(corpus, PreparedCorpus) else prepare_corpus(
        corpus, dtype=dtype, metric=metric, eps=DEFAULT_EPS if eps is None else eps, keep_source=False)
    q = _as_rows(queries, pc.device)
    if q.shape[1] != pc.dim:
        raise RuntimeError(f"query dim {q.shape[1]} does not match corpus dim {pc.dim}")
    out = torch.empty((q.shape[0], pc.n), dtype=torch.float32, device=pc.device)
    if out.numel() == 0:
        return out
    q_rows, _ = normalize_cast(q, pc.op, _lib.SIDE_QUERY, pc.metric == "cos", pc.eps)
    with torch.cuda.device(pc.device):
        rc = _lib.load().mmd_scores_dense(_ptr(q_rows), _ptr(pc.rows), _OP_DTYPE[pc.op], q.shape[0], pc.n, pc.dim, _ptr(out),
                                          out.stride(0), _stream_ptr(pc.device))
    _lib.check(rc, "mmd_scores_dense")
    return out


def merge_topk(scores: torch.Tensor, idx: torch.Tensor, k_out: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """K4: merge [parts, Q, k_in] partial lists (e.g. all-gathered per-rank results) -> [Q, k_out]."""
    _require_cuda(scores.device)
    parts, n_queries, k_in = scores.shape
    scores = scores.contiguous().float()
    idx = idx.contiguous().to(torch.int32)
    out_s = torch.empty((n_queries, k_out), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((n_queries, k_out), dtype=torch.int32, device=scores.device)
    if n_queries:
        with torch.cuda.device(scores.device):
            rc = _lib.load().mmd_topk_merge(_ptr(scores), _ptr(idx), parts, n_queries, k_in, k_out, _ptr(out_s), _ptr(out_i),
                                            _stream_ptr(scores.device))
        _lib.check(rc, "mmd_topk_merge")
    return out_s, out_i


def merge_pairs(pairs: torch.Tensor, k_out: int, n_queries: Optional[int] = None, k_in: Optional[int] = None,
                parts_sorted: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """K4 over packed lists -> (scores f32, rows i32) [Q, k_out].  pairs int32: either [parts, Q, k_in, 2] (dense), or
    [parts, cap, 2] with the first Q * k_in pairs of every part in use (pass n_queries and k_in; parts are `cap` apart)."""
    _require_cuda(pairs.device)
    assert pairs.dtype == torch.int32 and pairs.shape[-1] == 2 and pairs.is_contiguous()
    if pairs.dim() == 4:
        parts, n_queries, k_in, _ = pairs.shape
        stride = n_queries * k_in
    else:
        parts, stride, _ = pairs.shape
        assert n_queries is not None and k_in is not None and n_queries * k_in <= stride
    out_s = torch.empty((n_queries, k_out), dtype=torch.float32, device=pairs.device)
    out_i = torch.empty((n_queries, k_out), dtype=torch.int32, device=pairs.device)
    if n_queries:
        with torch.cuda.device(pairs.device):
            rc = _lib.load().mmd_topk_merge_pairs(_ptr(pairs), parts, stride, n_queries, k_in, k_out, int(parts_sorted), _ptr(out_s),
                                                  _ptr(out_i), _stream_ptr(pairs.device))
        _lib.check(rc, "mmd_topk_merge_pairs")
    return out_s, out_i


def merge_pairs_at(region: torch.Tensor, part_stride: int, k_out: int, n_queries: int, k_in: int,
                   parts_sorted: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """merge_pairs over a window of a larger gather buffer: `region` is a (non-contiguous) view whose first element is the
    first pair of part 0; consecutive parts are part_stride pairs apart."""
    _require_cuda(region.device)
    parts = region.shape[0]
    out_s = torch.empty((n_queries, k_out), dtype=torch.float32, device=region.device)
    out_i = torch.empty((n_queries, k_out), dtype=torch.int32, device=region.device)
    if n_queries:
        with torch.cuda.device(region.device):
            rc = _lib.load().mmd_topk_merge_pairs(_ptr(region), parts, part_stride, n_queries, k_in, k_out, int(parts_sorted), _ptr(out_s),
                                                  _ptr(out_i), _stream_ptr(region.device))
        _lib.check(rc, "mmd_topk_merge_pairs")
    return out_s, out_i


def scatter_pairs(scores: torch.Tensor, idx: torch.Tensor, dst_ptrs: Sequence[int], dst_offset_pairs: int = 0) -> None:
    """Pack ranked lists into {score bits, row} pairs and store them to every device pointer in dst_ptrs."""
    _require_cuda(scores.device)
    n_queries, k = scores.shape
    if n_queries == 0:
        return
    assert scores.dtype == torch.float32 and idx.dtype == torch.int32 and scores.is_contiguous() and idx.is_contiguous()
    arr = (C.c_void_p * len(dst_ptrs))(*[C.c_void_p(int(p)) for p in dst_ptrs])
    with torch.cuda.device(scores.device):
        rc = _lib.load().mmd_scatter_pairs(_ptr(scores), _ptr(idx), n_queries, k, arr, len(dst_ptrs), int(dst_offset_pairs),
                                           _stream_ptr(scores.device))
    _lib.check(rc, "mmd_scatter_pairs")


# ---------------------------------------------------------------------------------------------- row-sharded step, stage by stage
def _vp(ptrs: Sequence[int]):
    return (C.c_void_p * max(len(ptrs), 1))(*[C.c_void_p(int(p)) for p in ptrs])


def sharded_candidates(q_rows: torch.Tensor, n_queries: int, corpus, k_loc: int, raw_s: torch.Tensor, raw_i: torch.Tensor,
                       ws: torch.Tensor, thr_all: Sequence[int], own: int, share_thr: bool, gather_ptrs: Sequence[int],
                       pair_offset: int, pair_width: int, arrive_ptrs: Sequence[int] = (), sync_ptr: int = 0,
                       reset_thr: bool = True, merge_stream=None) -> None:
    """Stage C of a sharded step (mmd_sharded_candidates) on the current stream: fused top-k_loc over this rank's shard; the
    strip merge (on merge_stream, if given) stores the list, padded to pair_width, at pair pair_offset + q * pair_width into
    every gather buffer and raises the arrive flags.  corpus: PreparedCorpus or JointCorpus (rows / op / n / dim / idx_offset
    are used)."""
    dev = corpus.device
    with torch.cuda.device(dev):
        rc = _lib.load().mmd_sharded_candidates(
            _ptr(q_rows), C.c_void_p(corpus.rows.data_ptr() if corpus.n else 0), _OP_DTYPE[corpus.op], n_queries, corpus.n, corpus.dim,
            k_loc, corpus.idx_offset, _ptr(raw_s), _ptr(raw_i), _ptr(ws), ws.numel(),
            C.c_void_p(int(thr_all[own]) if thr_all else 0), _vp(thr_all), len(thr_all) if share_thr else 0, int(reset_thr),
            _vp(gather_ptrs), len(gather_ptrs), int(pair_offset), int(pair_width),
            _vp(arrive_ptrs), len(arrive_ptrs), C.c_void_p(int(sync_ptr)), _stream_ptr(dev),
            C.c_void_p(0 if merge_stream is None else merge_stream.cuda_stream))
    _lib.check(rc, "mmd_sharded_candidates")


def zero_u32(ptr: int, n: int, device: torch.device) -> None:
    """n 32-bit words at device pointer `ptr` <- 0 on the current stream (memset node, no kernel)."""
    with torch.cuda.device(device):
        _lib.check(_lib.load().mmd_zero_u32(C.c_void_p(int(ptr)), int(n), _stream_ptr(device)), "mmd_zero_u32")


def segment_tables(q_mats: Sequence[torch.Tensor], row0: int, q_invs: Sequence[Optional[torch.Tensor]], c_srcs: Sequence[torch.Tensor],
                   c_invs: Sequence[Optional[torch.Tensor]], dims: Sequence[int], weights: Sequence[float]):
    """ctypes tables of the multi-modality re-score (argument block shared by mmd_rescore_joint and mmd_exchange_rescore) for
    the query rows starting at row0."""
    n_seg = len(dims)
    vp, ip, lp, fp = C.c_void_p * n_seg, C.c_int * n_seg, C.c_int64 * n_seg, C.c_float * n_seg
    return (n_seg,
            vp(*[C.c_void_p(m.data_ptr() + row0 * m.stride(0) * m.element_size()) for m in q_mats]),
            ip(*[_SRC_DTYPE[m.dtype] for m in q_mats]), lp(*[m.stride(0) for m in q_mats]),
            vp(*[C.c_void_p(0 if t is None else t.data_ptr()) for t in q_invs]),
            vp(*[C.c_void_p(s.data_ptr() if s.shape[0] else 0) for s in c_srcs]),
            ip(*[_SRC_DTYPE[s.dtype] for s in c_srcs]), lp(*[s.stride(0) if s.shape[0] else d for s, d in zip(c_srcs, dims)]),
            vp(*[C.c_void_p(0 if t is None else t.data_ptr()) for t in c_invs]), ip(*dims), fp(*weights))


def exchange_rescore(gather_ptr: int, parts: int, part_stride: int, n_queries: int, k_in: int, kc: int, tables, n_local: int,
                     idx_offset: int, resc_ptrs: Sequence[int], own: int, device: torch.device, wait_ptr: int = 0, n_wait: int = 0,
                     arrive_ptrs: Sequence[int] = (), sync_ptr: int = 0) -> None:
    """Stage X (mmd_exchange_rescore) on the current stream."""
    with torch.cuda.device(device):
        rc = _lib.load().mmd_exchange_rescore(
            C.c_void_p(int(gather_ptr)), parts, int(part_stride), n_queries, k_in, kc, *tables, n_local, idx_offset,
            _vp(resc_ptrs), len(resc_ptrs), own, 0, C.c_void_p(int(wait_ptr)), n_wait, _vp(arrive_ptrs), len(arrive_ptrs),
            C.c_void_p(int(sync_ptr)), _stream_ptr(device))
    _lib.check(rc, "mmd_exchange_rescore")


def exchange_finish(resc_ptr: int, n_queries: int, kc: int, k_out: int, out_s_ptr: int, out_i_ptr: int, idx64: bool,
                    device: torch.device, wait_ptr: int = 0, n_wait: int = 0, sync_ptr: int = 0) -> None:
    """Stage F (mmd_exchange_finish) on the current stream: out_s f32 [Q,k_out], out_i i64 / i32 [Q,k_out] at the given pointers."""
    with torch.cuda.device(device):
        rc = _lib.load().mmd_exchange_finish(C.c_void_p(int(resc_ptr)), n_queries, kc, k_out, C.c_void_p(int(out_s_ptr)),
                                             C.c_void_p(int(out_i_ptr)), int(idx64), C.c_void_p(int(wait_ptr)), n_wait,
                                             C.c_void_p(int(sync_ptr)), _stream_ptr(device))
    _lib.check(rc, "mmd_exchange_finish")


def dedupe_scores(scores: torch.Tensor, idx: torch.Tensor, top_k: int, gold_idx: Optional[torch.Tensor] = None
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Distinct-score filter on the device: ranked (scores f32, idx i32) [Q, k_in] -> (scores, idx) [Q, top_k] padded with
    (-inf, -1) and the number of entries kept per query.  gold_idx i32 [Q] (optional) exempts each query's gold row."""
    _require_cuda(scores.device)
    n_queries, k_in = scores.shape
    scores = scores.contiguous().float()
    idx = idx.contiguous().to(torch.int32)
    gold = None if gold_idx is None else gold_idx.to(device=scores.device, dtype=torch.int32).contiguous()
    out_s = torch.empty((n_queries, top_k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((n_queries, top_k), dtype=torch.int32, device=scores.device)
    out_n = torch.empty((n_queries,), dtype=torch.int32, device=scores.device)
    if n_queries:
        with torch.cuda.device(scores.device):
            rc = _lib.load().mmd_dedupe_scores(_ptr(scores), _ptr(idx), _ptr(gold), n_queries, k_in, top_k, _ptr(out_s), _ptr(out_i),
                                               _ptr(out_n), _stream_ptr(scores.device))
        _lib.check(rc, "mmd_dedupe_scores")
    return out_s, out_i, out_n


def profile_enable(on: bool) -> None:
    _lib.load().mmd_profile_enable(int(on))


def profile_collect(cap: int = 512):
    buf = (C.c_float * cap)()
    n = _lib.load().mmd_profile_collect(buf, cap)
    return [float(buf[i]) for i in range(n)]


def launch_count() -> int:
    return int(_lib.load().mmd_launch_count())
