"""CPU: host-side logic of the reference-facing adapters (no kernels are launched)."""
import pytest
import torch

import mmd_retrieval as m
from mmd_retrieval import ops, sharded
from mmd_retrieval.postfilter import dedupe_by_score, hits_at_k
from oracle import evalmetrics


def test_dedupe_matches_reference_semantics():
    ranked = [("a", 0.9), ("b", 0.9), ("c", 0.8), ("d", 0.8), ("e", 0.7)]
    for k in (0, 1, 2, 3, 10):
        assert dedupe_by_score(ranked, k) == evalmetrics.dedupe_first_of_each_score(ranked, k) if k else dedupe_by_score(ranked, k) == []
    gold = lambda key: key == "d"   # noqa: E731
    assert dedupe_by_score(ranked, 3, gold) == evalmetrics.dedupe_first_of_each_score(ranked, 3, gold)
    assert dedupe_by_score([], 5) == []


def test_hits_at_k_matches_oracle():
    lists = [["x", "g0"], ["g1", "y"], ["z", "w"], []]
    gold = ["g0", "g1", "g2", "g3"]
    assert hits_at_k(lists, gold, (1, 2, 5)) == evalmetrics.hits_at_k(lists, gold, (1, 2, 5))


@pytest.mark.parametrize("n,world", [(10, 1), (10, 3), (7, 8), (1000000, 8), (0, 4), (100000000, 8)])
def test_shard_bounds_partition(n, world):
    spans = [sharded.shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def test_overfetch_policy():
    assert ops.overfetch_for(10, 1_000_000) == 18
    assert ops.overfetch_for(5, 10_000) == 13
    assert ops.overfetch_for(100, 10_000_000) == 104      # leaves 24 free slots in the 128-entry candidate buffer
    assert ops.overfetch_for(120, 10_000_000) == 120
    assert ops.overfetch_for(10, 12) == 12
    assert ops.overfetch_for(1, 1) == 1


def test_as_rows_accepts_lists_arrays_and_1d():
    import numpy as np
    assert ops._as_rows([1.0, 2.0, 3.0]).shape == (1, 3)
    assert ops._as_rows(np.ones((4, 5), dtype=np.float64)).dtype == torch.float32
    assert ops._as_rows(torch.ones(4, 5, dtype=torch.float16)).dtype == torch.float16
    t = torch.ones(6, 8)[:, ::2]
    assert ops._as_rows(t).stride(1) == 1
    with pytest.raises(ValueError):
        ops._as_rows(torch.ones(2, 3, 4))


def test_semantic_search_rejects_python_score_functions():
    with pytest.raises(m.MmdError):
        m.semantic_search(torch.ones(1, 4), torch.ones(3, 4), score_function=lambda a, b: a @ b.T)


def test_bad_arguments():
    with pytest.raises(ValueError):
        m.prepare_corpus(torch.ones(3, 4), dtype="int8")
    with pytest.raises(ValueError):
        m.prepare_corpus(torch.ones(3, 4), metric="l2")
    with pytest.raises(ValueError):
        m.topk(torch.ones(1, 4), torch.ones(3, 4), 0)


def test_ordered_bound_matches_the_kernel_threshold_encoding():
    """ops._ordered_bound must produce what the device's publish_threshold() writes: float_to_ordered(x + 0.0) - 1
    (csrc/common.cuh), and no bound (0) for -inf, i.e. for lists that are not full yet."""
    import struct

    def device_encoding(f):
        u = struct.unpack("<I", struct.pack("<f", f + 0.0))[0]
        return ((~u) & 0xFFFFFFFF if u & 0x80000000 else u | 0x80000000) - 1

    vals = [0.5, -0.25, 0.0, -0.0, 1e-30, -1e-30, 123.0, -7.5, 3.4e38]
    got = ops._ordered_bound(torch.tensor(vals), 1.0).tolist()
    assert got == [device_encoding(v) for v in vals]
    assert ops._ordered_bound(torch.tensor([float("-inf")]), 1.0).tolist() == [0]
    # order preserving: a larger score gives a larger bound; the fp8 path's 2^16 accumulator scale is applied first
    s = torch.tensor([-2.0, -1.0, -0.5, 0.0, 0.25, 1.0])
    b = ops._ordered_bound(s, 65536.0)
    assert bool((b[1:] > b[:-1]).all()) and b.tolist() == [device_encoding(float(v) * 65536.0) for v in s]
