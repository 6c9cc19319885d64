"""CPU oracle for the evidence-retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it, and there only as the checker or the timed CPU
baseline.  The product path (multimodal-misinformation-detection_b200/) never imports this package and
has no CPU fallback.

Modules
  st_util.py     restatement of sentence-transformers==3.3.1 `util.semantic_search / cos_sim / dot_score /
                 normalize_embeddings` (third-party, NOT vendored in the reference; PARITY UNPINNED -- see its header)
  im2im.py       restatement of the reference's own image retrieval (src/evidence/im2im_retrieval.py) plus a
                 shim that imports and runs the reference's unmodified code when /root/reference is present
                 (PINNED: tests/golden/im2im_*.npz were produced by the reference's code, see make_golden.py)
  exact.py       batched float64 ground truth (normalise -> matmul -> ordered top-k) and the near-tie classifier
  evalmetrics.py restatement of the hits@k evaluation of src/evidence/experiment_{image,text}.py
  fusion.py      float64 statement of the joint image+text score sum_m w_m * cos_m (no reference counterpart: PARITY
                 UNPINNED) and a restatement of the reference's two-list concat + sort (text2text_retrieval.py:97-118)
  make_golden.py generator of tests/golden/*.npz (run in the build container, where /root/reference exists)
"""
