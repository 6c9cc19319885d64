"""CPU: the oracle against the golden vectors (produced by the reference's own code) and against itself."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import evalmetrics, exact, im2im, st_util


@pytest.mark.parametrize("name", ["im2im_a.npz", "im2im_b.npz"])
def test_im2im_restatement_matches_reference_golden(name):
    """oracle/im2im.py's restatement == the reference's retrieve_similar_images (golden made by the reference)."""
    g = load_golden(name)
    fd = {f"c{i:05d}": g["corpus_t"][i] for i in range(g["corpus_t"].shape[0])}
    top_k = int(g["top_k"])
    for qi in range(g["queries_t"].shape[0]):
        got = im2im.retrieve_similar(g["queries_t"][qi], fd, top_k)
        want_rows = [r for r in g["rows"][qi].tolist() if r >= 0]
        assert [int(k[1:]) for k, _ in got] == want_rows
        np.testing.assert_allclose([s for _, s in got], g["scores"][qi][: len(got)], rtol=0, atol=1e-6)
    batched = im2im.retrieve_similar_batched(g["queries_t"], fd, top_k)
    for qi, lst in enumerate(batched):
        want_rows = [r for r in g["rows"][qi].tolist() if r >= 0]
        got_rows = [int(k[1:]) for k, _ in lst]
        # the matmul formulation may reorder entries whose fp32 scores differ by rounding only
        assert set(got_rows) == set(want_rows) or _near_tie_only(got_rows, want_rows, g, qi)
        np.testing.assert_allclose(sorted(s for _, s in lst), sorted(g["scores"][qi][: len(lst)]), rtol=0, atol=5e-6)


def _near_tie_only(got_rows, want_rows, g, qi, tol=5e-6):
    full = exact.exact_scores(g["queries_t"][qi:qi + 1], g["corpus_t"], eps=1e-6)[0]
    diff = list(set(got_rows) ^ set(want_rows))
    kth = full[want_rows[-1]]
    return bool(((full[diff] - kth).abs() <= tol).all())


def test_pairwise_similarity_matches_reference_golden():
    g = load_golden("im2im_a.npz")
    for i, want in enumerate(g["pair_scores"]):
        assert im2im.similarity(g["queries_t"][i], g["corpus_t"][i]) == pytest.approx(want, abs=1e-7)
    dim = g["corpus_t"].shape[1]
    # per-norm clamp edge cases recorded from the reference: cos(1e-8*1, 1e-8*1) and cos(0, 1)
    tiny = torch.full((dim,), 1e-8)
    assert im2im.similarity(tiny, tiny) == pytest.approx(g["edge_scores"][0], rel=1e-6)
    assert im2im.similarity(torch.zeros(dim), torch.ones(dim)) == g["edge_scores"][1] == 0.0
    e = exact.exact_scores(tiny[None], tiny[None], eps=1e-6)[0, 0].item()
    assert e == pytest.approx(g["edge_scores"][0], rel=1e-4)


@pytest.mark.parametrize("name,dtype", [("t2t_fp32.npz", torch.float32), ("t2t_fp16.npz", torch.float16)])
def test_semantic_search_restatement_frozen(name, dtype):
    g = load_golden(name)
    q = torch.from_numpy(g["queries"]).to(dtype)
    c = torch.from_numpy(g["corpus"]).to(dtype)
    hits = st_util.semantic_search(q, c, top_k=int(g["top_k"]), corpus_chunk_size=int(g["corpus_chunk_size"]))
    rows = np.array([[h["corpus_id"] for h in hl] for hl in hits])
    scores = np.array([[h["score"] for h in hl] for hl in hits])
    if dtype == torch.float32:
        np.testing.assert_array_equal(rows, g["rows"])
    np.testing.assert_allclose(scores, g["scores"], rtol=0, atol=2e-3 if dtype == torch.float16 else 1e-6)


def test_semantic_search_contract():
    gen = torch.Generator().manual_seed(0)
    c = torch.randn(50, 16, generator=gen)
    q = torch.randn(16, generator=gen)                      # 1-D query is unsqueezed
    hits = st_util.semantic_search(q, c, top_k=7)
    assert len(hits) == 1 and len(hits[0]) == 7
    assert [h["score"] for h in hits[0]] == sorted((h["score"] for h in hits[0]), reverse=True)
    hits = st_util.semantic_search(q.numpy(), c.numpy(), top_k=100)   # ndarray inputs, top_k > N
    assert len(hits[0]) == 50
    # chunking does not change the result
    a = st_util.semantic_search(c[:9], c, top_k=5, query_chunk_size=2, corpus_chunk_size=7)
    b = st_util.semantic_search(c[:9], c, top_k=5)
    assert [[h["corpus_id"] for h in x] for x in a] == [[h["corpus_id"] for h in x] for x in b]
    d = st_util.semantic_search(c[:3], c, top_k=3, score_function=st_util.dot_score)
    full = c[:3] @ c.T
    assert [h["corpus_id"] for h in d[1]] == torch.topk(full[1], 3).indices.tolist()


def test_exact_agrees_with_restatements():
    gen = torch.Generator().manual_seed(3)
    q, c = torch.randn(20, 64, generator=gen), torch.randn(300, 64, generator=gen)
    c[7] = 0
    hits = st_util.semantic_search(q, c, top_k=10)
    vals, idx = exact.exact_topk(q, c, 10, eps=1e-12)
    assert [[h["corpus_id"] for h in hl] for hl in hits] == idx.tolist()
    np.testing.assert_allclose([[h["score"] for h in hl] for hl in hits], vals.numpy(), atol=1e-6)
    cmp = exact.compare_topk(vals.float(), idx, exact.exact_scores(q, c), 10, tie_tol=0.0)
    assert cmp.ok and cmp.identical_order == 20 and cmp.max_score_err < 1e-6


def test_compare_topk_flags_real_mismatches_and_excuses_near_ties():
    full = torch.tensor([[0.9, 0.5, 0.5 + 1e-9, 0.1, 0.0]], dtype=torch.float64)
    ok = exact.compare_topk(torch.tensor([[0.9, 0.5]]), torch.tensor([[0, 1]]), full, 2, tie_tol=1e-6)
    assert ok.ok and ok.excused_rows == 1
    bad = exact.compare_topk(torch.tensor([[0.9, 0.1]]), torch.tensor([[0, 3]]), full, 2, tie_tol=1e-6)
    assert not bad.ok
    dup = exact.compare_topk(torch.tensor([[0.9, 0.9]]), torch.tensor([[0, 0]]), full, 2, tie_tol=1e-6)
    assert not dup.ok


def test_ordered_topk_tie_rule():
    s = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0]])
    v, i = exact.ordered_topk(s, 4)
    assert i.tolist() == [[1, 2, 4, 3]]
    v, i = exact.ordered_topk(s, 10)
    assert i.shape == (1, 5)


def test_dedupe_and_hits_restatement():
    ranked = [("a", 0.9), ("b", 0.9), ("c", 0.8), ("d", 0.8), ("e", 0.7)]
    assert evalmetrics.dedupe_first_of_each_score(ranked, 3) == [("a", 0.9), ("c", 0.8), ("e", 0.7)]
    assert evalmetrics.dedupe_first_of_each_score(ranked, 2) == [("a", 0.9), ("c", 0.8)]
    assert evalmetrics.dedupe_first_of_each_score(ranked, 3, gold=lambda k: k == "d") == [("a", 0.9), ("c", 0.8), ("d", 0.8)]
    assert evalmetrics.dedupe_first_of_each_score(ranked, 10) == [("a", 0.9), ("c", 0.8), ("e", 0.7)]
    h = evalmetrics.hits_at_k([["x", "g"], ["g", "y"], ["z", "w"]], ["g", "g", "g"], (1, 2))
    assert h == {1: 1 / 3, 2: 2 / 3}


def test_image_eval_planted_positive():
    gen = torch.Generator().manual_seed(5)
    c = torch.relu(torch.randn(200, 64, generator=gen))
    q = c[:10] + 0.05 * torch.randn(10, 64, generator=gen)
    c[100:105] = c[0:5]                               # duplicates of gold rows: identical scores
    keys = [f"k{i}" for i in range(200)]
    acc = evalmetrics.image_eval(exact.exact_scores(q, c, eps=1e-6), keys, [f"k{i}" for i in range(10)])
    assert acc[1] == 1.0 and acc[10] == 1.0


def test_fusion_oracle_consistency():
    """oracle/fusion.py: weights (1, 0) reduce to the single-modality oracle; the fused matrix is linear in the weights;
    concat_and_sort restates the reference's two-list merge (text2text_retrieval.py:97-118)."""
    from oracle import exact, fusion
    gen = torch.Generator().manual_seed(5)
    qs = [torch.randn(7, 16, generator=gen), torch.randn(7, 24, generator=gen)]
    cs = [torch.randn(50, 16, generator=gen), torch.randn(50, 24, generator=gen)]
    a = fusion.fused_scores(qs, cs, (1.0, 0.0))
    assert torch.allclose(a, exact.exact_scores(qs[0], cs[0]))
    mix = fusion.fused_scores(qs, cs, (0.3, 0.7))
    assert torch.allclose(mix, 0.3 * exact.exact_scores(qs[0], cs[0]) + 0.7 * exact.exact_scores(qs[1], cs[1]))
    s, i = fusion.fused_topk(qs, cs, (0.5, 0.5), 5)
    assert bool((s[:, :-1] >= s[:, 1:]).all()) and tuple(i.shape) == (7, 5)
    merged = fusion.concat_and_sort([("train_1", 0.9), ("train_2", 0.5)], [("test_3", 0.9), ("test_4", 0.7)], 3)
    assert merged == [("train_1", 0.9), ("test_4", 0.7), ("train_2", 0.5)]
