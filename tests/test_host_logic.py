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


def test_fp8_needs_unit_norm_operands():
    """e4m3 operands carry a fixed 2^8 scale sized for |x| <= 1: inner-product scoring of un-normalised rows and modality
    weights above 1.75 are refused up front -- before any device is touched."""
    with pytest.raises(ValueError, match="fp8"):
        m.prepare_corpus(torch.ones(3, 16), dtype="fp8", metric="dot")
    with pytest.raises(ValueError, match="fp8"):
        m.prepare_joint([torch.ones(3, 16), torch.ones(3, 16)], weights=(0.5, 0.5), dtype="fp8", metric="dot")
    with pytest.raises(ValueError, match="1.75"):
        m.prepare_joint([torch.ones(3, 16), torch.ones(3, 16)], weights=(2.0, 0.5), dtype="fp8")


class _FakeShard:
    """Just enough of a PreparedCorpus for the host-side planning code (no device memory)."""
    source = object()
    n, dim, op, metric, eps, idx_offset = 125_000, 768, "bf16", "cos", 1e-12, 0
    rows = torch.empty((0, 1536), dtype=torch.uint8)
    device = torch.device("cpu")


def test_sharded_step_planning():
    """Sub-batch split (whole 256-query tiles, sums to Q), exchanged list widths, and the layout of the symmetric ring."""
    sc = sharded.ShardedCorpus(None, 1_000_000, 0, _shard=_FakeShard())
    assert sc._sub_sizes(16384) == [16384] and sc._sub_sizes(1) == [1] and sc._sub_sizes(0) == []
    sc.sub_batches = 2
    assert sc._sub_sizes(16384) == [8192, 8192] and sc._sub_sizes(300) == [256, 44] and sc._sub_sizes(256) == [256]
    sc.sub_batches = 5
    for q in (1, 255, 257, 1000, 16384, 65536):
        sizes = sc._sub_sizes(q)
        assert sum(sizes) == q and all(s % 256 == 0 for s in sizes[:-1]) and len(sizes) <= 5
    # (k_glob, K' every rank exchanges, global candidate list): a single rank re-scores all K'
    assert sc._widths(10) == (10, 18, 18) and sc._widths(100) == (100, 104, 104)
    sc.world, sc._max_local, sc._min_local, sc.n_total = 4, 11, 10, 41           # 41 rows over 4 ranks
    assert sc._widths(10) == (10, 11, 18)                                          # lists shorter than the global candidate list
    sc.n_total, sc._max_local = 5, 2
    assert sc._widths(10) == (5, 2, 5)


def test_compaction_target_leaves_room_for_the_pending_appends():
    """Mirror of compaction_target() in csrc/topk_fused.cu: a compaction must leave at most CAP - 16 entries (8 slots for the
    group of appends that triggered it) whenever K' allows, and never fewer than K'."""
    def target(cap, kprime):
        slack = max((cap - kprime) // 3, 8)
        want, most = kprime + slack, cap - 16
        return want if want < most else (most if most > kprime else kprime)
    for cap, ks in ((64, range(1, 33)), (128, range(33, 121))):
        for k in ks:
            t = target(cap, k)
            assert k <= t and (t <= cap - 16 or t == k)
    assert target(64, 18) == 33 and target(128, 104) == 112 and target(128, 120) == 120


def test_auto_splits_rule():
    """ops.auto_splits: sub-searches needed to reach the over-fetch target of the operand type (no GPU involved)."""
    import importlib
    ops = importlib.import_module("mmd_retrieval.ops")
    big = 10_000_000
    assert ops.auto_splits("bf16", 10, 18, big) == 1            # the default path is untouched
    assert ops.auto_splits("bf16", 100, 104, big) == 2
    assert ops.auto_splits("fp16", 100, 104, big) == 2
    assert ops.auto_splits("fp8", 10, 18, big) == 3
    assert ops.auto_splits("fp8", 100, 104, big) == 4
    assert ops.auto_splits("fp32", 100, 100, big) == 1          # exact operands: nothing to recover
    assert ops.auto_splits("fp8", 100, 104, 300_000) == 2       # never below 131072 rows per split
    assert ops.auto_splits("fp8", 100, 104, 60_000) == 1
    assert ops.auto_splits("fp8", 5, 13, big) == 2


def test_quantile_level_bound_is_a_valid_lower_bound():
    """The merged bound of the fused kernel (csrc/topk_fused.cu, "Merged bounds"), restated on the host: units publish their
    score at rank ceil(K'/2^j) into slot (strip mod 2^j) of level j (max per slot); the bound is the max over levels of the min
    over a level's slots, an empty slot voiding its level.  Whatever subset of strips has finished, in whatever order, and
    however many equal scores there are, at least K' corpus rows must score at or above the bound."""
    import random
    rng = random.Random(7)
    for trial in range(300):
        kprime = rng.choice([1, 2, 5, 18, 33, 104, 120])
        n_strips = rng.randint(1, 40)
        levels = 0
        while levels < 4 and (2 << levels) <= n_strips:
            levels += 1
        quant = rng.choice([None, 0.05, 0.5])                      # coarse scores: plenty of ties
        strips = []
        for _ in range(n_strips):
            rows = rng.randint(1, 400)
            sc = [rng.gauss(0.0, 1.0) for _ in range(rows)]
            if quant:
                sc = [round(x / quant) * quant for x in sc]
            strips.append(sorted(sc, reverse=True))
        slots = {(j, s): None for j in range(1, levels + 1) for s in range(1 << j)}
        order = list(range(n_strips))
        rng.shuffle(order)
        finished = []
        for strip in order[: rng.randint(1, n_strips)]:
            finished.append(strip)
            lst = strips[strip][:kprime]                           # a unit keeps at most K' rows of its strip
            for j in range(1, levels + 1):
                rank = -(-kprime // (1 << j))
                if len(lst) >= rank:
                    key = (j, strip & ((1 << j) - 1))
                    slots[key] = lst[rank - 1] if slots[key] is None else max(slots[key], lst[rank - 1])
            best = None
            for j in range(1, levels + 1):
                vals = [slots[(j, s)] for s in range(1 << j)]
                if all(v is not None for v in vals):
                    best = min(vals) if best is None else max(best, min(vals))
            if best is not None:
                at_or_above = sum(1 for st in finished for x in strips[st] if x >= best)
                assert at_or_above >= kprime, (trial, kprime, n_strips, levels, best, at_or_above)


def _ordered(x: float) -> int:
    import struct
    u = struct.unpack("<I", struct.pack("<f", x + 0.0))[0]
    return (~u) & 0xFFFFFFFF if u & 0x80000000 else u | 0x80000000


def _select_bound(ords, kprime, keep_hi):
    """csrc/topk_fused.cu select_batch, restated: 4-ary search on the order-preserving score word for a bound with between
    kprime and keep_hi entries above it.  Returns (bound, count) or None when equal scores straddle the kprime-th best."""
    n = len(ords)
    lo, hi, c_lo = min(ords) - 1, max(ords), n
    if n <= keep_hi or n < kprime:
        return lo, n
    for _ in range(40):
        d = hi - lo
        q = d >> 2
        t1 = lo + q if q else lo + (d >> 1)
        t2 = t1 + q if q else t1
        t3 = t2 + q if q else t1
        c1, c2, c3 = (sum(1 for o in ords if o > t) for t in (t1, t2, t3))
        if t1 == lo:
            return None
        if c3 >= kprime:
            lo, c_lo = t3, c3
        elif c2 >= kprime:
            lo, c_lo, hi = t2, c2, t3
        elif c1 >= kprime:
            lo, c_lo, hi = t1, c1, t2
        else:
            hi = t1
        if c_lo <= keep_hi:
            return lo, c_lo
    return None


def test_selection_compaction_keeps_the_best_kprime():
    """The compaction of long lists never drops one of a row's kprime best entries, leaves at most keep_hi, and reports
    rows it cannot split (ties around the kprime-th best) instead of guessing."""
    import random
    rng = random.Random(11)
    fallbacks = 0
    for trial in range(400):
        kprime = rng.choice([33, 64, 100, 104, 120])
        n = rng.randint(kprime, 127)
        keep_hi = min(kprime + max(1, (124 - kprime) // 8), 124)
        quant = rng.choice([None, None, 0.01, 0.2])
        scores = [rng.gauss(0.1, 0.04) * rng.choice([1.0, 1.0, -1.0]) for _ in range(n)]
        if quant:
            scores = [round(x / quant) * quant for x in scores]
        ords = [_ordered(x) for x in scores]
        assert all((a < b) == (_ordered(a) < _ordered(b)) for a, b in zip(scores, scores[1:]))      # the mapping preserves order
        res = _select_bound(ords, kprime, keep_hi)
        if res is None:
            fallbacks += 1
            ranked = sorted(ords, reverse=True)
            assert ranked[kprime - 1] == ranked[min(keep_hi, n - 1)]            # only a tie run longer than the window fails
            continue
        bound, count = res
        kept = [o for o in ords if o > bound]
        assert len(kept) == count and (kprime <= count <= keep_hi or count == n <= keep_hi)
        assert sorted(kept, reverse=True)[:kprime] == sorted(ords, reverse=True)[:kprime]
    assert fallbacks < 200
