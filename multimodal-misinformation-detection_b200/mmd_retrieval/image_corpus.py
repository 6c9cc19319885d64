"""Drop-in for the reference's image retrieval classes (boundary B2 of SURVEY.md section 8b).

  ImageSimilarity.similarity(f1, f2) -> float                       src/evidence/im2im_retrieval.py:38-42
  ImageCorpus.retrieve_similar_images(query_image_path, top_k=50)   src/evidence/im2im_retrieval.py:80-106
      -> list[(key, score)], score descending, first entry of each distinct score, at most top_k entries
  calculate_topk_accuracy_image_retrieval(...)                      src/evidence/experiment_image.py:7-63

The reference scores the query against every corpus entry in a Python loop (one nn.CosineSimilarity call
per pair) and sorts all N scores.  Here the feature dict is stacked once into an HBM-resident prepared
corpus; a query (or a batch of queries) is one fused CUDA top-K' call, and only the dedupe walk over the
K' survivors stays on the host.  The ResNet-50 feature extractor is NOT part of the path: any object with
`extract_features(path) -> Tensor[D]` can be plugged in, exactly as `self.feature_extractor` in the reference.
"""
from __future__ import annotations

import pickle
from typing import Callable, Dict, Hashable, List, Optional, Sequence, Tuple

import torch

from . import ops
from .postfilter import dedupe_by_score, hits_at_k

IMAGE_EPS = 1e-6   # nn.CosineSimilarity(dim=1, eps=1e-6), im2im_retrieval.py:40


class ImageSimilarity:
    """Pairwise cosine with the reference's per-norm clamp (eps=1e-6), computed on the GPU."""

    def __init__(self, feature_extractor=None, dtype: str = "fp32"):
        self._extractor = feature_extractor
        self._dtype = dtype

    def extract_features(self, image_stream):
        if self._extractor is None:
            raise RuntimeError("no feature extractor attached (the image encoder is outside the retrieval path)")
        return self._extractor.extract_features(image_stream)

    def similarity(self, features1: torch.Tensor, features2: torch.Tensor) -> float:
        s = ops.dense_scores(features1.reshape(1, -1), features2.reshape(1, -1), metric="cos", dtype=self._dtype,
                             eps=IMAGE_EPS)
        return s.item()


class ImageCorpus:
    """Image-feature evidence corpus with GPU top-K retrieval.

    feature_dict keeps the reference's format: dict[key -> 1-D feature tensor] in insertion order
    (im2im_retrieval.py:51-62); insertion order is the corpus row order, which makes the tie rule
    (equal scores -> earlier entry first) the same as the reference's stable sort.
    """

    def __init__(self, feature_corpus_path: Optional[str] = None, feature_dict: Optional[Dict[Hashable, torch.Tensor]] = None,
                 feature_extractor=None, dtype: str = "bf16", device=None):
        self.feature_corpus_path = feature_corpus_path
        if feature_dict is None:
            feature_dict = self.load_features() if feature_corpus_path is not None else {}
        self.feature_dict = feature_dict if feature_dict is not None else {}
        self.feature_extractor = feature_extractor if feature_extractor is not None else ImageSimilarity()
        self.dtype = dtype
        self._device = device
        self._keys: List[Hashable] = []
        self._prepared: Optional[ops.PreparedCorpus] = None
        self._prepared_len = -1

    # -- corpus container (same on-disk format as the reference) ---------------------------------
    def load_features(self):
        try:
            with open(self.feature_corpus_path, "rb") as f:
                return pickle.load(f)
        except (EOFError, pickle.UnpicklingError):
            print("Warning: Pickle file is empty or corrupted. Initializing empty feature dict.")
            return {}

    def save_features(self):
        with open(self.feature_corpus_path, "wb") as f:
            pickle.dump(self.feature_dict, f)

    def add_image(self, image_path):
        self.feature_dict[image_path] = self.feature_extractor.extract_features(image_path)
        self._prepared = None
        if self.feature_corpus_path is not None:
            self.save_features()

    # -- prepared device corpus -------------------------------------------------------------------
    def prepared(self) -> ops.PreparedCorpus:
        if self._prepared is None or self._prepared_len != len(self.feature_dict):
            self._keys = list(self.feature_dict.keys())
            if self._keys:
                feats = torch.stack([torch.as_tensor(v).reshape(-1).float() for v in self.feature_dict.values()])
            else:
                feats = torch.empty((0, 1), dtype=torch.float32)
            self._prepared = ops.prepare_corpus(feats, dtype=self.dtype, metric="cos", eps=IMAGE_EPS, device=self._device)
            self._prepared_len = len(self.feature_dict)
        return self._prepared

    # -- retrieval --------------------------------------------------------------------------------
    def retrieve_similar_features(self, query_features, top_k: int = 50,
                                  is_gold: Optional[Callable[[int, Hashable], bool]] = None,
                                  gold_rows: Optional[Sequence[int]] = None) -> List[List[Tuple[Hashable, float]]]:
        """Batched form: query_features [Q,D] (or [D]) -> one deduped (key, score) list per query.

        The distinct-score walk (im2im_retrieval.py:94-104) runs on the device over the K' over-fetched candidates
        (mmd_dedupe_scores); only the final top_k entries per query come back to the host.  gold_rows[q] (corpus row of
        query q's gold evidence, -1 = none) enables the evaluation's gold exemption (experiment_image.py:41-50) on the
        device; an arbitrary `is_gold(q, key)` predicate is honoured by a host walk instead."""
        pc = self.prepared()
        q = ops._as_rows(query_features, pc.device)
        n_queries = q.shape[0]
        if pc.n == 0 or top_k <= 0:
            return [[] for _ in range(n_queries)]
        out: List[Optional[List[Tuple[Hashable, float]]]] = [None] * n_queries
        pending = list(range(n_queries))
        gold_t = None
        if gold_rows is not None and is_gold is None:
            gold_t = torch.as_tensor(list(gold_rows), dtype=torch.int32, device=pc.device)
        fetch = min(pc.n, max(top_k + 8, 2 * top_k), ops.max_k())
        while pending:
            whole = len(pending) == n_queries
            sub = q if whole else q[pending]
            if fetch >= pc.n and pc.n > ops.max_k():
                # duplicate-heavy corpus exhausted the fused K limit: rank the dense score row on the device
                dense = ops.dense_scores(sub, pc.source, metric="cos", dtype="fp32", eps=IMAGE_EPS)
                scores, idx = torch.sort(dense, dim=1, descending=True, stable=True)
                idx = idx.to(torch.int32)
            else:
                scores, idx = ops.topk(sub, pc, fetch, index_dtype=torch.int32)
            n_ranked = scores.shape[1]
            if is_gold is None:
                gsub = None if gold_t is None else (gold_t if whole else gold_t[pending])
                ks, ki, kn = ops.dedupe_scores(scores, idx, min(top_k, n_ranked), gsub)
                s_host, i_host, n_host = ks.cpu().tolist(), ki.cpu().tolist(), kn.cpu().tolist()
                kept_lists = [[(self._keys[i], s) for s, i in zip(s_host[pos][:n_host[pos]], i_host[pos][:n_host[pos]])]
                              for pos in range(len(pending))]
            else:
                s_host, i_host = scores.cpu().tolist(), idx.cpu().tolist()
                kept_lists = []
                for pos, qi in enumerate(pending):
                    ranked = [(self._keys[i], s) for s, i in zip(s_host[pos], i_host[pos]) if i >= 0]
                    kept_lists.append(dedupe_by_score(ranked, top_k, lambda key, qi=qi: is_gold(qi, key)))
            still = []
            for pos, qi in enumerate(pending):
                if len(kept_lists[pos]) == top_k or n_ranked >= pc.n:
                    out[qi] = kept_lists[pos]
                else:
                    still.append(qi)
            pending = still
            if pending:
                fetch = pc.n if fetch >= ops.max_k() else min(pc.n, ops.max_k(), fetch * 2)
        return out  # type: ignore[return-value]

    def retrieve_similar_images(self, query_image_path, top_k: int = 50) -> List[Tuple[Hashable, float]]:
        query_features = self.feature_extractor.extract_features(query_image_path)
        return self.retrieve_similar_features(torch.as_tensor(query_features).reshape(1, -1), top_k)[0]


def calculate_topk_accuracy_image_retrieval(image_corpus: ImageCorpus, query_features, gold_keys: Sequence[Hashable],
                                            k_values: Sequence[int] = (1, 2, 5, 10)) -> Dict[int, float]:
    """hits@k of each query's gold evidence key, with the reference's gold exemption in the dedupe
    (src/evidence/experiment_image.py:41-61).  query_features [Q,D]; gold_keys[q] = key of the paired evidence."""
    top_k = max(k_values)
    image_corpus.prepared()
    row_of = {key: r for r, key in enumerate(image_corpus._keys)}
    lists = image_corpus.retrieve_similar_features(query_features, top_k, gold_rows=[row_of.get(g, -1) for g in gold_keys])
    return hits_at_k([[k for k, _ in lst] for lst in lists], list(gold_keys), k_values)
