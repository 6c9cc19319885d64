"""Generate tests/golden/*.npz.  TEST INFRASTRUCTURE ONLY; run in the build container:

    python -m oracle.make_golden

im2im_*.npz   produced by the REFERENCE'S OWN code (src/evidence/im2im_retrieval.py imported unmodified through
              oracle/im2im.py's shim): inputs + the (row, score) lists retrieve_similar_images returned.
t2t_*.npz     produced by oracle/st_util.py, the restatement of sentence_transformers.util.semantic_search
              (the third-party package is absent, so these vectors are NOT pinned by the reference; they freeze the
              restatement so that it cannot drift silently).
Inputs are stored in the files (fp16-representable values, so the files stay small) -- nothing is re-derived from a
seed at test time.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _image_like(gen, rows, dim):
    # post-ReLU average-pooled CNN features: non-negative, many exact zeros (dataset_search.ipynb:282-284)
    x = torch.relu(torch.randn(rows, dim, generator=gen))
    return x.half().float()   # fp16-representable so the npz can store fp16 losslessly


def make_im2im(name: str, seed: int, n_corpus: int, n_query: int, dim: int, top_k: int, n_dups: int, n_zero: int):
    from oracle import im2im
    gen = torch.Generator().manual_seed(seed)
    corpus = _image_like(gen, n_corpus, dim)
    for d in range(n_dups):          # duplicated evidence images: identical features -> identical scores
        corpus[n_corpus - 1 - d] = corpus[d]
    for z in range(n_zero):          # degenerate all-zero feature rows (norm clamp path)
        corpus[n_corpus // 2 + z] = 0.0
    queries = _image_like(gen, n_query, dim)
    queries[0] = corpus[3] * 2.0     # an exact (scaled) match: cosine 1
    corpus, queries = corpus.half().float(), queries.half().float()   # what the npz stores, losslessly
    feature_dict = {f"c{idx:05d}": corpus[idx].clone() for idx in range(n_corpus)}
    qdict = {f"q{idx:03d}": queries[idx].clone() for idx in range(n_query)}
    res = im2im.reference_retrieve(feature_dict, qdict, top_k)
    rows = np.full((n_query, top_k), -1, dtype=np.int32)
    scores = np.full((n_query, top_k), np.nan, dtype=np.float64)
    for qi, qk in enumerate(qdict):
        for j, (ck, sc) in enumerate(res[qk]):
            rows[qi, j] = int(ck[1:])
            scores[qi, j] = sc
    pair = np.array([im2im.reference_similarity(queries[i], corpus[i]) for i in range(min(n_query, 8))], dtype=np.float64)
    tiny = torch.full((dim,), 1e-8)
    edge = np.array([im2im.reference_similarity(tiny, tiny), im2im.reference_similarity(torch.zeros(dim), torch.ones(dim))])
    np.savez_compressed(os.path.join(GOLDEN, name), corpus=corpus.numpy().astype(np.float16),
                        queries=queries.numpy().astype(np.float16), top_k=np.int32(top_k), rows=rows, scores=scores,
                        pair_scores=pair, edge_scores=edge)
    print(name, "rows", rows.shape, "first list", rows[0, :5], scores[0, :3])


def make_t2t(name: str, seed: int, n_corpus: int, n_query: int, dim: int, top_k: int, dtype: torch.dtype, chunk: int):
    from oracle import st_util
    gen = torch.Generator().manual_seed(seed)
    corpus = torch.randn(n_corpus, dim, generator=gen).half()
    queries = torch.randn(n_query, dim, generator=gen).half()
    for i in range(min(n_query, n_corpus // 4)):     # planted positives
        corpus[4 * i] = (queries[i].float() + 0.5 * torch.randn(dim, generator=gen)).half()
    hits = st_util.semantic_search(queries.to(dtype), corpus.to(dtype), top_k=top_k, corpus_chunk_size=chunk)
    rows = np.array([[h["corpus_id"] for h in hl] for hl in hits], dtype=np.int32)
    scores = np.array([[h["score"] for h in hl] for hl in hits], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, name), corpus=corpus.numpy(), queries=queries.numpy(), top_k=np.int32(top_k),
                        rows=rows, scores=scores, dtype=str(dtype), corpus_chunk_size=np.int32(chunk))
    print(name, rows.shape, rows[0, :5], scores[0, :3])


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    sys.path.insert(0, os.path.dirname(HERE))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    make_im2im("im2im_a.npz", seed=11, n_corpus=300, n_query=6, dim=2048, top_k=10, n_dups=5, n_zero=2)
    make_im2im("im2im_b.npz", seed=12, n_corpus=64, n_query=3, dim=2048, top_k=50, n_dups=10, n_zero=1)
    make_t2t("t2t_fp32.npz", seed=21, n_corpus=700, n_query=16, dim=768, top_k=5, dtype=torch.float32, chunk=256)
    make_t2t("t2t_fp16.npz", seed=22, n_corpus=700, n_query=16, dim=768, top_k=25, dtype=torch.float16, chunk=500000)


if __name__ == "__main__":
    main()
