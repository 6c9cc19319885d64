"""Drop-in for the reference's text-evidence search (boundary B1 of SURVEY.md section 8b, caller side).

    SemanticSimilarity.search(query, top_k) -> list[(id_str, score)]        src/evidence/text2text_retrieval.py:49-120

The reference's search is: bi-encode the claim -> util.semantic_search against the train and the test corpus with
top_k*5 each (re-normalising both corpora every call) -> cross-encoder re-rank -> map rows to ids -> concatenate ->
sort -> keep the first entry of every distinct score until top_k.  Only the two semantic_search calls are on the
hot path; the encoders are NOT (they are plugged in: any object with `encode(text) -> Tensor[D]` / any callable
`rerank(query, texts) -> scores`).  Here both corpora are prepared once (normalised, cast, HBM-resident) and a whole
batch of claims is searched with two fused CUDA calls; the tail stays host Python as in the reference.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple, Union

import torch

from . import ops
from .postfilter import dedupe_by_score, hits_at_k


def _decode(i) -> str:
    return i.decode("utf-8") if isinstance(i, bytes) else str(i)


class SemanticSimilarity:
    """Text-evidence retrieval over the train + test evidence corpora.

    train_embeddings / test_embeddings: [N,768] tensors (fp16 in the reference, text2text_retrieval.py:44) or
    PreparedCorpus objects (e.g. from corpus_io.load_text_corpus); *_ids: their row-aligned ids (b"train_17" ...).
    bi_encoder: object with encode(str) -> Tensor[D] (optional: embeddings can be passed to search directly).
    cross_encoder: callable (query, list_of_texts) -> list of scores, or an object with a batched predict(list of [query,
    text] pairs) (mmd_retrieval.cross_encoder.BatchedCrossEncoder; a sentence-transformers CrossEncoder), with train_texts /
    test_texts giving the text of every corpus row (the reference's `evidence_enriched` column); without it the bi-encoder
    cosine is the final score.
    """

    OVERFETCH = 5     # top_k * 5 per corpus, text2text_retrieval.py:57,62

    def __init__(self, train_embeddings, train_ids: Sequence, test_embeddings, test_ids: Sequence, bi_encoder=None,
                 cross_encoder: Optional[Callable] = None, train_texts: Optional[Sequence[str]] = None,
                 test_texts: Optional[Sequence[str]] = None, dtype: str = "bf16", device=None):
        prep = lambda c: c if isinstance(c, ops.PreparedCorpus) else ops.prepare_corpus(c, dtype=dtype, device=device)  # noqa: E731
        self.train, self.test = prep(train_embeddings), prep(test_embeddings)
        self.train_ids, self.test_ids = list(train_ids), list(test_ids)
        if len(self.train_ids) != self.train.n or len(self.test_ids) != self.test.n:
            raise ValueError("ids must be row-aligned with the embeddings")
        self.bi_encoder, self.cross_encoder = bi_encoder, cross_encoder
        self.train_texts, self.test_texts = train_texts, test_texts

    def _embed(self, queries) -> torch.Tensor:
        if isinstance(queries, str):
            queries = [queries]
        if isinstance(queries, (list, tuple)) and queries and isinstance(queries[0], str):
            if self.bi_encoder is None:
                raise RuntimeError("no bi-encoder attached (the sentence encoder is outside the retrieval path); pass embeddings")
            return torch.stack([torch.as_tensor(self.bi_encoder.encode(t)).reshape(-1).float() for t in queries])
        return ops._as_rows(queries)

    def search_batch(self, queries, top_k: int, query_texts: Optional[Sequence[str]] = None, overfetch: Optional[int] = None,
                     gold_ids: Optional[Sequence[str]] = None) -> List[List[Tuple[str, float]]]:
        """One deduped [(id, score)] list per claim.  queries: strings (needs the bi-encoder) or embeddings [Q,D].
        overfetch: hits per corpus = top_k * overfetch (5 in search, 10 in the evaluation script); gold_ids[q]: id that
        the distinct-score filter must keep for claim q even if its score repeats (experiment_text.py:80)."""
        if isinstance(queries, str):
            queries = [queries]
        if query_texts is None and isinstance(queries, (list, tuple)) and queries and isinstance(queries[0], str):
            query_texts = list(queries)
        emb = self._embed(queries)
        k_each = top_k * (self.OVERFETCH if overfetch is None else overfetch)
        lists = []
        for corpus in (self.train, self.test):
            s, i = ops.topk(emb, corpus, k_each, dense_fallback=True)
            lists.append((s.cpu().tolist(), i.cpu().tolist()))
        # Re-rank (text2text_retrieval.py:67-95).  A cross-encoder with a batched `predict(pairs)` (mmd_retrieval.cross_encoder.
        # BatchedCrossEncoder, or a sentence-transformers CrossEncoder) scores the pairs of ALL claims and both corpora in one
        # call; a plain callable (query, texts) -> scores is called per claim and corpus as the reference does.
        rerank = self.cross_encoder is not None and query_texts is not None
        batched_scores = None
        if rerank and hasattr(self.cross_encoder, "predict"):
            pairs, where = [], []
            for ci, ((s_host, i_host), texts) in enumerate(zip(lists, (self.train_texts, self.test_texts))):
                if texts is None:
                    continue
                for qi in range(emb.shape[0]):
                    for pos, row in enumerate(i_host[qi]):
                        if row >= 0:
                            pairs.append((query_texts[qi], texts[row]))
                            where.append((ci, qi, pos))
            flat = self.cross_encoder.predict(pairs) if pairs else []
            batched_scores = {w: float(c) for w, c in zip(where, flat)}
        out = []
        for qi in range(emb.shape[0]):
            results: List[Tuple[str, float]] = []
            for ci, ((s_host, i_host), ids, texts) in enumerate(zip(lists, (self.train_ids, self.test_ids), (self.train_texts, self.test_texts))):
                hits = [(row, score) for score, row in zip(s_host[qi], i_host[qi]) if row >= 0]
                if rerank and texts is not None:
                    if batched_scores is not None:
                        cross = [batched_scores[(ci, qi, pos)] for pos, row in enumerate(i_host[qi]) if row >= 0]
                    else:
                        cross = self.cross_encoder(query_texts[qi], [texts[row] for row, _ in hits])
                    hits = sorted(((row, float(c)) for (row, _), c in zip(hits, cross)), key=lambda t: t[1], reverse=True)[:k_each]
                results += [(_decode(ids[row]), score) for row, score in hits]
            ranked = sorted(results, key=lambda t: t[1], reverse=True)
            gold = None if gold_ids is None else (lambda key, g=gold_ids[qi]: key == g)
            out.append(dedupe_by_score(ranked, top_k, gold))
        return out

    def search(self, query: Union[str, torch.Tensor], top_k: int) -> List[Tuple[str, float]]:
        """The reference's signature: one claim in, [(id, score)] out (text2text_retrieval.py:49)."""
        return self.search_batch([query] if isinstance(query, str) else query, top_k)[0]


def calculate_topk_accuracy_text_retrieval(similarity: SemanticSimilarity, queries, k_values: Sequence[int] = (1, 2, 5, 10),
                                           query_texts: Optional[Sequence[str]] = None,
                                           gold_ids: Optional[Sequence[str]] = None) -> dict:
    """hits@k of every claim's paired evidence, the reference's text-retrieval evaluation
    (src/evidence/experiment_text.py:11-106): top_k*10 hits from the train and from the test corpus, optional
    cross-encoder re-rank, concatenation, descending sort, distinct-score filter that always keeps the gold evidence
    `test_{query_id}`, then membership of the gold id in the first k entries.  All claims are searched in two fused GPU
    calls instead of one call pair per claim."""
    n = len(queries) if not isinstance(queries, torch.Tensor) else (1 if queries.dim() == 1 else queries.shape[0])
    gold = [f"test_{i}" for i in range(n)] if gold_ids is None else list(gold_ids)
    top_k = max(k_values)
    lists = similarity.search_batch(queries, top_k, query_texts=query_texts, overfetch=10, gold_ids=gold)
    return hits_at_k([[key for key, _ in lst] for lst in lists], gold, k_values)
