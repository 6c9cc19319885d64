"""GPU (B200): the CUDA path, called through the C ABI (ctypes -> libmmd.so), against the CPU oracle.

Tolerances (BASELINE.json north_star):
  * index sets identical to the oracle except at score near-ties (oracle.exact.compare_topk);
  * scores within 1e-5 relative for the fp32 configuration and for every result that went through the exact
    re-score (the default); raw bf16 tensor-core scores within 1e-3 relative on image-like (non-negative) features,
    and within 1e-3 of the unit-norm scale on near-orthogonal Gaussian pairs (operand rounding is relative to
    ||q||*||c|| = 1, not to a score that is itself ~0.1 -- see DESIGN.md "Numerics").
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import evalmetrics, exact, im2im, st_util

pytestmark = pytest.mark.gpu

FP32_RTOL = 1e-5
BF16_RTOL = 1e-3


@pytest.fixture(scope="module")
def m():
    import mmd_retrieval
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    assert mmd_retrieval.LIB_PATH.exists(), "libmmd.so must be built in-tree"
    return mmd_retrieval


def _data(kind, rows, dim, seed):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, dim, generator=gen)
    return torch.relu(x) if kind == "image" else x


def _device_operands(m, x, op, side, eps, normalize=True):
    """The operand values K1 actually produced for `x` (float64 [rows, dim]).  K1 itself is checked against
    F.normalize in test_normalize_cast_*; feeding ITS output to the oracle makes the K2/K3 checks exact with
    respect to the operands (only fp32 accumulation order remains)."""
    rows, _ = m.normalize_cast(x.cuda(), op, side, normalize, eps)
    dim = x.shape[1]
    if op == "bf16":
        return rows.view(torch.bfloat16)[:, :dim].double().cpu()
    if op == "fp16":
        return rows.view(torch.float16)[:, :dim].double().cpu()
    if op == "fp8":
        return rows.view(torch.float8_e4m3fn)[:, :dim].double().cpu() / 256.0
    kd, _ = m.ops.prepared_layout("fp32", dim)
    seg = rows.view(torch.bfloat16).double().cpu().view(x.shape[0], 6, kd // 6)[:, :, :dim]
    order = (2, 0, 1, 1, 0, 0) if side == 0 else (0, 2, 1, 0, 1, 0)
    limb = {order[s_]: seg[:, s_] for s_ in range(6)}
    return limb[0] + limb[1] + limb[2]


def _device_scores(m, q, c, op, eps, metric="cos"):
    qv = _device_operands(m, q, op, 0, eps, metric == "cos")
    cv = _device_operands(m, c, op, 1, eps, metric == "cos")
    return qv @ cv.T


# ------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("dim", [768, 2048, 100, 8, 4100])
@pytest.mark.parametrize("src", [torch.float32, torch.float16, torch.bfloat16])
def test_normalize_cast_bf16(m, dim, src):
    from mmd_retrieval import _lib
    x = _data("text", 333, dim, dim).to(src)
    x[5] = 0                                           # zero row -> clamp path, output 0
    x[6] = 1e-20 if src == torch.float32 else x[6]     # norm below eps
    rows, inv = m.normalize_cast(x.cuda(), "bf16", _lib.SIDE_CORPUS, True, 1e-12)
    got = rows.view(torch.bfloat16).float().cpu()
    want32 = torch.nn.functional.normalize(x.float(), p=2, dim=1, eps=1e-12)
    want = want32.to(torch.bfloat16).float()
    assert got.shape[1] >= dim and torch.count_nonzero(got[:, dim:]) == 0          # K padding is zero
    d = (got[:, :dim] - want).abs()
    # identical up to the last-place rounding of the norm: at most one bf16 ulp, on a vanishing fraction
    assert float((d > 0).float().mean()) < 2e-3
    assert bool((d <= want.abs() * 2 ** -7 + 1e-30).all())
    assert torch.count_nonzero(got[5]) == 0
    want_inv = 1.0 / x.float().norm(dim=1).clamp_min(1e-12)
    torch.testing.assert_close(inv.cpu(), want_inv, rtol=1e-6, atol=0)


def test_normalize_cast_layouts(m):
    from mmd_retrieval import _lib
    x = _data("text", 64, 200, 1)
    xn = torch.nn.functional.normalize(x, p=2, dim=1, eps=1e-12)
    # fp16 operands
    r16, _ = m.normalize_cast(x.cuda(), "fp16", _lib.SIDE_QUERY, True, 1e-12)
    assert float((r16.view(torch.float16).float().cpu()[:, :200] - xn.half().float()).abs().max()) <= 2 ** -11
    # fp8 operands carry a fixed 2^8 scale
    r8, _ = m.normalize_cast(x.cuda(), "fp8", _lib.SIDE_CORPUS, True, 1e-12)
    got8 = r8.view(torch.float8_e4m3fn).float().cpu()[:, :200] / 256
    want8 = (xn * 256).to(torch.float8_e4m3fn).float() / 256
    assert float((got8 != want8).float().mean()) < 2e-3
    # fp32 configuration: three bf16 limbs reconstruct the fp32 value; limb order differs per side
    kd, rb = m.ops.prepared_layout("fp32", 200)
    dpad = kd // 6
    for side, order in ((_lib.SIDE_QUERY, (2, 0, 1, 1, 0, 0)), (_lib.SIDE_CORPUS, (0, 2, 1, 0, 1, 0))):
        r, _ = m.normalize_cast(x.cuda(), "fp32", side, True, 1e-12)
        seg = r.view(torch.bfloat16).float().cpu().view(64, 6, dpad)[:, :, :200]
        limb = {order[s]: seg[:, s] for s in range(6)}
        rec = limb[0] + limb[1] + limb[2]
        assert float((rec - xn).abs().max()) <= 2 ** -24
        for s in range(6):
            assert torch.equal(seg[:, s], limb[order[s]])
    # metric="dot": no normalisation; strided source rows
    big = _data("text", 32, 512, 2).cuda()
    view = big[:, :256]
    rd, inv = m.normalize_cast(view, "bf16", _lib.SIDE_CORPUS, False, 1e-12)
    assert torch.equal(rd.view(torch.bfloat16).float().cpu(), view.cpu().to(torch.bfloat16).float())
    assert torch.equal(inv.cpu(), torch.ones(32))


# ------------------------------------------------------------------------------------------ K2 (dense)
@pytest.mark.parametrize("op", ["bf16", "fp16", "fp8", "fp32"])
@pytest.mark.parametrize("shape", [(128, 256, 64), (300, 1000, 200), (129, 257, 768), (37, 41, 2048)])
def test_dense_scores_match_oracle_on_device_operands(m, op, shape):
    q_n, c_n, dim = shape
    q, c = _data("text", q_n, dim, 1), _data("text", c_n, dim, 2)
    got = m.dense_scores(q.cuda(), c.cuda(), metric="cos", dtype=op).cpu().double()
    want = _device_scores(m, q, c, op, 1e-12)
    # identical operands; fp32 tensor-core accumulation vs float64
    assert float((got - want).abs().max()) <= 2e-6
    if op == "fp32":
        # and against the un-rounded float64 ground truth: 1e-5 relative (+ fp32 rounding noise at unit scale)
        truth = exact.exact_scores(q, c, "cos", 1e-12)
        assert bool(((got - truth).abs() <= FP32_RTOL * truth.abs() + 3e-7).all())
    else:
        # operand rounding itself stays inside the bound the format implies (relative to ||q||*||c|| = 1)
        truth = exact.exact_scores(q, c, "cos", 1e-12)
        bound = {"bf16": 2e-3, "fp16": 3e-4, "fp8": 5e-2}[op]   # D = 64 rows have few, large elements
        assert float((got - truth).abs().max()) <= bound


def test_dense_dot_metric_and_pairwise_similarity(m):
    q, c = _data("text", 10, 96, 3), _data("text", 20, 96, 4)
    got = m.dense_scores(q.cuda(), c.cuda(), metric="dot", dtype="fp32").cpu().double()
    want = q.double() @ c.double().T
    assert float(((got - want).abs() / want.abs().clamp_min(1.0)).max()) <= FP32_RTOL
    g = load_golden("im2im_a.npz")
    sim = m.ImageSimilarity()
    for i, want_s in enumerate(g["pair_scores"]):
        assert sim.similarity(g["queries_t"][i], g["corpus_t"][i]) == pytest.approx(want_s, rel=FP32_RTOL)
    dim = g["corpus_t"].shape[1]
    assert sim.similarity(torch.zeros(dim), torch.ones(dim)) == 0.0
    tiny = torch.full((dim,), 1e-8)
    assert sim.similarity(tiny, tiny) == pytest.approx(g["edge_scores"][0], rel=1e-3)   # per-norm clamp, eps=1e-6


# ------------------------------------------------------------------------------------------ K2+K3 fused top-K
CASES = [
    # kind,  Q,    N,     D,   k,  op
    ("text", 1000, 10000, 768, 5, "fp32"),      # config 1 shape
    ("text", 300, 5000, 768, 10, "bf16"),
    ("image", 256, 6000, 2048, 10, "bf16"),     # config 2 shape, scaled down
    ("text", 64, 20000, 768, 100, "bf16"),      # eval over-fetch (experiment_text.py:26)
    ("text", 40, 3000, 512, 25, "fp16"),
    ("text", 129, 257, 64, 7, "bf16"),          # one past every tile edge
    ("text", 127, 255, 100, 7, "bf16"),         # D not a multiple of 8
    ("text", 1, 41256, 768, 50, "bf16"),        # one query, the reference's usage pattern
    ("text", 33, 9000, 768, 10, "bf16"),        # small batches stage only the 32-row query groups that exist (1, 2, 3 groups)
    ("text", 64, 9000, 768, 10, "fp8"),
    ("image", 70, 5000, 2048, 10, "bf16"),
]


@pytest.mark.parametrize("kind,q_n,c_n,dim,k,op", CASES)
def test_topk_raw_matches_operand_rounded_oracle(m, kind, q_n, c_n, dim, k, op):
    """No re-score: the fused kernel's own selection vs the float64 oracle fed the operands K1 produced."""
    q, c = _data(kind, q_n, dim, 10), _data(kind, c_n, dim, 11)
    eps = 1e-6 if kind == "image" else 1e-12
    pc = m.prepare_corpus(c.cuda(), dtype=op, eps=eps, keep_source=False)
    s, i = m.topk(q.cuda(), pc, k, rescore_exact=False)
    assert s.dtype == torch.float32 and i.dtype == torch.int64 and tuple(s.shape) == (q_n, min(k, c_n))
    full = _device_scores(m, q, c, op, eps)
    cmp = exact.compare_topk(s, i, full, k, tie_tol=2e-6)
    assert cmp.ok, cmp
    assert cmp.max_score_err <= 2e-6, cmp
    assert bool((s[:, :-1] >= s[:, 1:]).all())


@pytest.mark.parametrize("kind,q_n,c_n,dim,k,op", CASES)
def test_topk_default_path_matches_fp32_reference(m, kind, q_n, c_n, dim, k, op):
    """Default path (tensor-core selection + exact re-score) vs the float64 ground truth AND the restated
    reference semantic_search (fp32): identical index sets up to near-ties, scores within 1e-5 relative."""
    q, c = _data(kind, q_n, dim, 10), _data(kind, c_n, dim, 11)
    eps = 1e-6 if kind == "image" else 1e-12
    s, i = m.topk(q.cuda(), m.prepare_corpus(c.cuda(), dtype=op, eps=eps), k)
    full = exact.exact_scores(q, c, "cos", eps)
    cmp = exact.compare_topk(s, i, full, k, tie_tol=2e-6)
    assert cmp.ok, cmp
    assert cmp.max_rel_score_err <= FP32_RTOL, cmp
    if kind == "text" and q_n <= 300:
        hits = st_util.semantic_search(q, c, top_k=k)
        ref_idx = torch.tensor([[h["corpus_id"] for h in hl] for hl in hits])
        ref_s = torch.tensor([[h["score"] for h in hl] for hl in hits])
        cmp2 = exact.compare_topk(ref_s, ref_idx, full, k, tie_tol=2e-6)
        assert cmp2.ok
        agree = sum(set(a) == set(b) for a, b in zip(ref_idx.tolist(), i.cpu().tolist()))
        assert agree >= q_n - cmp.excused_rows - cmp2.excused_rows


def test_raw_bf16_scores_within_1e3_relative_on_image_features(m):
    q, c = _data("image", 128, 2048, 20), _data("image", 4000, 2048, 21)
    s, i = m.topk(q.cuda(), m.prepare_corpus(c.cuda(), dtype="bf16", eps=1e-6, keep_source=False), 10, rescore_exact=False)
    full = exact.exact_scores(q, c, "cos", 1e-6)
    picked = torch.gather(full, 1, i.cpu())
    assert float(((s.cpu().double() - picked).abs() / picked.abs()).max()) <= BF16_RTOL
    # Gaussian (near-orthogonal) pairs: error is relative to ||q||*||c|| = 1
    q, c = _data("text", 128, 768, 22), _data("text", 4000, 768, 23)
    s, i = m.topk(q.cuda(), m.prepare_corpus(c.cuda(), dtype="bf16", keep_source=False), 10, rescore_exact=False)
    picked = torch.gather(exact.exact_scores(q, c), 1, i.cpu())
    assert float((s.cpu().double() - picked).abs().max()) <= BF16_RTOL


# ------------------------------------------------------------------------------------------ edge cases
def test_edge_cases(m):
    # K > N, N = 1, Q = 1, K = 1
    c = _data("text", 7, 64, 30)
    q = _data("text", 5, 64, 31)
    s, i = m.topk(q.cuda(), c.cuda(), 10)
    assert tuple(s.shape) == (5, 7)
    assert torch.equal(i.cpu(), exact.exact_topk(q, c, 10)[1])
    s, i = m.topk(q[:1].cuda(), c[:1].cuda(), 3)
    assert tuple(i.shape) == (1, 1) and int(i) == 0
    s, i = m.topk(q[0].cuda(), c.cuda(), 1)                       # 1-D query is unsqueezed
    assert tuple(i.shape) == (1, 1) and int(i) == int(exact.exact_topk(q[:1], c, 1)[1])
    # empty query batch / empty corpus
    s, i = m.topk(torch.empty(0, 64).cuda(), c.cuda(), 3)
    assert tuple(s.shape) == (0, 3)
    s, i = m.topk(q.cuda(), torch.empty(0, 64).cuda(), 3)
    assert tuple(s.shape) == (5, 0)
    # zero query row: every score is 0 -> ties -> ascending rows
    z = torch.zeros(1, 64)
    s, i = m.topk(z.cuda(), c.cuda(), 4)
    assert i.cpu().tolist() == [[0, 1, 2, 3]] and torch.count_nonzero(s) == 0
    # duplicate corpus rows: equal scores come back in ascending row order (the reference's stable sort)
    c2 = _data("text", 600, 64, 32)
    c2[500] = c2[3]
    c2[100] = c2[3]
    for rescore in (False, True):
        s, i = m.topk(c2[3:4].cuda() * 3.0, m.prepare_corpus(c2.cuda(), dtype="bf16"), 5, rescore_exact=rescore)
        assert i.cpu().tolist()[0][:3] == [3, 100, 500]
        assert float(s[0, 0]) == float(s[0, 1]) == float(s[0, 2])
    # all rows identical: pure tie-breaking across tiles and strips
    same = torch.ones(3000, 64)
    s, i = m.topk(torch.ones(2, 64).cuda(), same.cuda(), 20, rescore_exact=False)
    assert i.cpu().tolist() == [list(range(20))] * 2
    # the largest supported K, and one beyond
    big = _data("text", 2000, 64, 33)
    s, i = m.topk(q.cuda(), m.prepare_corpus(big.cuda(), keep_source=False), 120)
    assert exact.compare_topk(s, i, _device_scores(m, q, big, "bf16", 1e-12), 120, 2e-6).ok
    with pytest.raises(m.MmdError):
        m.topk(q.cuda(), big.cuda(), 121)
    # ... unless the dense fallback is asked for (semantic_search does): exact lists of any length
    s, i = m.topk(q.cuda(), big.cuda(), 300, dense_fallback=True)
    assert tuple(i.shape) == (5, 300) and exact.compare_topk(s, i, exact.exact_scores(q, big), 300, tie_tol=2e-6).ok
    hits = m.semantic_search(q[:2], big, top_k=1500)
    assert len(hits[0]) == 1500 and hits[0][0]["corpus_id"] == int(exact.exact_topk(q[:1], big, 1)[1])
    with pytest.raises(RuntimeError):
        m.topk(torch.ones(2, 32).cuda(), big.cuda(), 3)           # dim mismatch, like torch.mm upstream
    # inner-product metric
    s, i = m.topk(q.cuda(), m.prepare_corpus(big.cuda(), dtype="fp32", metric="dot"), 6)
    want_s, want_i = exact.exact_topk(q, big, 6, metric="dot")
    assert torch.equal(i.cpu(), want_i)
    assert float(((s.cpu().double() - want_s).abs() / want_s.abs()).max()) <= FP32_RTOL


# ------------------------------------------------------------------------------------------ K3b / K4 / K5
def test_merge_kernel(m):
    gen = torch.Generator().manual_seed(40)
    for parts, n_q, k_in, k_out in [(2, 50, 10, 10), (8, 33, 100, 100), (3, 7, 5, 12), (37, 9, 18, 18), (1, 4, 4, 2),
                                      (40, 5, 120, 64), (3, 6, 200, 130), (64, 3, 100, 100)]:
        s = torch.randn(parts, n_q, k_in, generator=gen).round(decimals=1)       # plenty of equal scores
        idx = torch.stack([torch.stack([torch.randperm(1000, generator=gen)[:k_in] for _ in range(n_q)]) + 1000 * p
                           for p in range(parts)]).to(torch.int32)
        idx[0, 0, -2:] = -1
        s[0, 0, -2:] = float("-inf")
        got_s, got_i = m.merge_topk(s.cuda(), idx.cuda(), k_out)
        flat_s = s.permute(1, 0, 2).reshape(n_q, -1)
        flat_i = idx.permute(1, 0, 2).reshape(n_q, -1).long()
        for r in range(n_q):
            cand = sorted(((-float(a), int(b)) for a, b in zip(flat_s[r], flat_i[r]) if b >= 0))[:k_out]
            want_i = [b for _, b in cand] + [-1] * (k_out - len(cand))
            want_s = [-a for a, _ in cand] + [float("-inf")] * (k_out - len(cand))
            assert got_i[r].cpu().tolist() == want_i
            assert got_s[r].cpu().tolist() == want_s


def test_sharded_equals_unsharded(m):
    """Row-sharding invariance on one GPU: per-shard top-K with global offsets + K4 merge == unsharded top-K."""
    q, c = _data("text", 200, 768, 50), _data("text", 9001, 768, 51)
    want_s, want_i = m.topk(q.cuda(), c.cuda(), 10)
    from mmd_retrieval.sharded import shard_bounds
    ss, ii = [], []
    for r in range(3):
        lo, hi = shard_bounds(c.shape[0], 3, r)
        pc = m.prepare_corpus(c[lo:hi].cuda(), idx_offset=lo)
        s, i = m.topk(q.cuda(), pc, 10, index_dtype=torch.int32)
        assert int(i.min()) >= lo and int(i.max()) < hi
        ss.append(s)
        ii.append(i)
    s, i = m.merge_topk(torch.stack(ss), torch.stack(ii), 10)
    assert torch.equal(i.long(), want_i) and torch.equal(s, want_s)


@pytest.mark.parametrize("op,k", [("bf16", 10), ("fp8", 100)])
def test_phased_sweep_equals_single_launch(m, op, k):
    """Several launches over contiguous row ranges with the merged K-th best carried between them as pruning bound give
    the same lists as one launch (the bound only removes rows that cannot make the final list)."""
    q, c = _data("text", 200, 768, 66), _data("text", 300_000, 768, 67)
    c[299_999] = c[3]
    q[0] = c[3] * 2
    pc = m.prepare_corpus(c.cuda(), dtype=op, keep_source=False)
    want_s, want_i = m.topk(q.cuda(), pc, k, rescore_exact=False)
    for phases in (2, 4):
        s, i = m.topk(q.cuda(), pc, k, rescore_exact=False, phases=phases)
        assert torch.equal(i, want_i) and torch.equal(s, want_s), phases
    pc2 = m.prepare_corpus(c.cuda(), dtype=op)
    s, i = m.topk(q.cuda(), pc2, k, phases=3)
    s1, i1 = m.topk(q.cuda(), pc2, k)
    assert torch.equal(i, i1) and torch.equal(s, s1)


def test_cuda_graph_replay_single_gpu(m):
    """ShardedCorpus.capture on one GPU: the whole step (K1, fused top-K', strip merge, re-score) replayed from a CUDA
    graph gives the same bits as the eager call, for fresh query batches copied into the static input."""
    q, c = _data("text", 300, 768, 62), _data("text", 9000, 768, 63)
    sc = m.ShardedCorpus(c.cuda(), 9000, 0)
    want_s, want_i = sc.topk(q.cuda(), 10)
    graphed = sc.capture(q.cuda(), 10)
    for rep in range(3):
        qq = q if rep != 1 else q.flip(0)
        s, i = graphed(qq.cuda())
        torch.cuda.synchronize()
        ws, wi = (want_s, want_i) if rep != 1 else (want_s.flip(0), want_i.flip(0))
        assert torch.equal(i, wi) and torch.equal(s, ws)


def test_packed_pairs_exchange_layout(m):
    """The row-sharded path's exchange format on one GPU: every 'rank' re-scores its shard straight into its slot of a
    [world, Q, k] gather buffer of {score bits, global row} pairs (two destination buffers at once, as with
    peer-mapped buffers), then K4 merges the packed lists: identical to the unsharded result."""
    from mmd_retrieval import ops
    from mmd_retrieval.sharded import shard_bounds
    q, c = _data("text", 130, 768, 60), _data("text", 7001, 768, 61)
    c[7000] = c[2]                                               # equal scores on different shards
    k, world = 10, 3
    want_s, want_i = m.topk(q.cuda(), c.cuda(), k)
    bufs = [torch.zeros((world, q.shape[0], k, 2), dtype=torch.int32, device="cuda") for _ in range(2)]
    for r in range(world):
        lo, hi = shard_bounds(c.shape[0], world, r)
        pc = m.prepare_corpus(c[lo:hi].cuda(), idx_offset=lo)
        qd, q_inv, _, cand = ops.topk_candidates(q.cuda(), pc, k)
        ops.rescore_pairs(qd, q_inv, pc, cand, k, [b.data_ptr() for b in bufs], dst_offset_pairs=r * q.shape[0] * k)
    assert torch.equal(bufs[0], bufs[1])
    s, i = ops.merge_pairs(bufs[0], k)
    assert torch.equal(i.long(), want_i) and torch.equal(s, want_s)
    # a shard smaller than k pads its slot with (-inf, -1)
    pc = m.prepare_corpus(c[:4].cuda())
    qd, q_inv, _, cand = ops.topk_candidates(q.cuda(), pc, k)
    one = torch.zeros((1, q.shape[0], k, 2), dtype=torch.int32, device="cuda")
    ops.rescore_pairs(qd, q_inv, pc, cand, k, [one.data_ptr()])
    assert bool((one[0, :, 4:, 1] == -1).all()) and bool((one[0, :, :4, 1] >= 0).all())
    assert bool(torch.isinf(one[0, :, 4:, 0].view(torch.float32)).all())


def test_global_rescore_flow_with_shared_thresholds_emulated_on_one_gpu(m):
    """The production multi-GPU step, stage by stage, with three 'ranks' emulated on one device: per-shard tensor-core
    pass whose pruning thresholds are SHARED between the shards (mmd_topk_scores_shared: a bound learnt on one shard
    prunes the others) and whose strip merge stores the raw candidates straight into every rank's gather buffer ->
    strided merge of the raw lists -> each shard re-scores only the global candidates it owns into the second region ->
    final merge.  Must equal the unsharded result bit for bit."""
    from mmd_retrieval import ops, _lib
    from mmd_retrieval.sharded import shard_bounds
    q, c = _data("text", 300, 768, 64), _data("text", 20011, 768, 65)
    c[20010] = c[7]                                                # equal scores on different shards
    q[0] = c[7] * 1.5
    k, world = 10, 3
    want_s, want_i = m.topk(q.cuda(), c.cuda(), k)
    n_q, kp = q.shape[0], ops.overfetch_for(k, 6671)
    cap = n_q * (kp + k)
    gather = [torch.zeros((world, cap, 2), dtype=torch.int32, device="cuda") for _ in range(world)]     # one buffer per "rank"
    thr = [torch.zeros((n_q,), dtype=torch.int32, device="cuda") for _ in range(world)]
    shards, qd, q_inv = [], None, None
    for r in range(world):                                          # stage 1 on every rank (sequentially here)
        lo, hi = shard_bounds(c.shape[0], world, r)
        pc = m.prepare_corpus(c[lo:hi].cuda(), idx_offset=lo)
        shards.append(pc)
        qd, q_inv, raw_s, cand = ops.topk_candidates(q.cuda(), pc, k, overfetch=kp,
                                                     shared_thr=(thr[r].data_ptr(), [t.data_ptr() for t in thr]),
                                                     pair_dst=([g.data_ptr() for g in gather], r * cap))
        assert cand.shape[1] == kp
        # every bound this shard learnt was published to EVERY rank's threshold array
        assert int((thr[0] != 0).sum()) == n_q and all(torch.equal(t, thr[0]) for t in thr[1:])
        # the raw list arrived, packed, in slot r of every rank's buffer
        assert torch.equal(gather[0][r, :n_q * kp, 1].view(n_q, kp), cand) and torch.equal(gather[0], gather[world - 1])
    for g in gather[1:]:
        assert torch.equal(g, gather[0])
    _, cand_glob = ops.merge_pairs(gather[0], kp, n_queries=n_q, k_in=kp, parts_sorted=True)
    for r in range(world):                                          # stage 2: every rank re-scores what it owns
        ops.rescore_pairs(qd, q_inv, shards[r], cand_glob, k, [g.data_ptr() for g in gather], dst_offset_pairs=r * cap + n_q * kp)
    s, i = ops.merge_pairs_at(gather[1][:, n_q * kp:, :], cap, k, n_q, k, parts_sorted=True)
    assert torch.equal(i.long(), want_i) and torch.equal(s, want_s)
    assert i[0, :2].tolist() == [7, 20010]


@pytest.mark.parametrize("k,n_rows,joint", [(10, 20011, False), (100, 9001, False), (10, 41, False), (10, 7001, True)])
def test_three_stage_exchange_emulated_on_one_gpu(m, k, n_rows, joint):
    """The sharded step of mmd_retrieval.sharded, stage by stage (C: mmd_sharded_candidates, X: mmd_exchange_rescore,
    F: mmd_exchange_finish), with three 'ranks' run one after the other on one device and NO flags (n_wait = n_arrive = 0:
    kernels that wait for one another must not share a GPU).  Every rank's gather / re-score buffer is its own allocation,
    written by all ranks -- as with peer-mapped buffers.  Must equal the unsharded result bit for bit."""
    from mmd_retrieval import ops, joint as jt, _lib
    from mmd_retrieval.sharded import shard_bounds
    world, n_q = 3, 300
    if joint:
        dims, weights = (256, 128), (0.6, 0.4)
        cs = [_data("text", n_rows, d, 70 + j) for j, d in enumerate(dims)]
        qs = [_data("text", n_q, d, 80 + j) for j, d in enumerate(dims)]
        want_s, want_i = m.topk_joint([x.cuda() for x in qs], m.prepare_joint([x.cuda() for x in cs], weights), k)
    else:
        cs, qs = [_data("text", n_rows, 768, 66)], [_data("text", n_q, 768, 67)]
        cs[0][n_rows - 1] = cs[0][7]                                  # equal scores on different shards
        qs[0][0] = cs[0][7] * 1.5
        want_s, want_i = m.topk(qs[0].cuda(), cs[0].cuda(), k)
    q_dev = [x.cuda() for x in qs]
    max_local = -(-n_rows // world)
    k_glob = min(k, n_rows)
    kp = ops.overfetch_for(min(k, max_local), max_local)
    kc = min(world * kp, ops.overfetch_for(k_glob, n_rows))
    share = kp >= kc and n_rows // world >= kc
    part_stride = n_q * kp
    gather = [torch.full((world * part_stride, 2), 7, dtype=torch.int32, device="cuda") for _ in range(world)]
    resc = [torch.full((n_q * kc, 2), 7, dtype=torch.int32, device="cuda") for _ in range(world)]
    thr = [torch.zeros((n_q,), dtype=torch.int32, device="cuda") for _ in range(world)]
    shards, q_invs = [], None
    for r in range(world):                                          # stage C on every rank
        lo, hi = shard_bounds(n_rows, world, r)
        if joint:
            pc = m.prepare_joint([x[lo:hi].cuda() for x in cs], weights, idx_offset=lo)
            q_rows, q_invs = jt._cast_segments(q_dev, pc.op, _lib.SIDE_QUERY, True, pc.eps, pc.weights, pc.seg_bytes, sum(pc.seg_bytes))
        else:
            pc = m.prepare_corpus(cs[0][lo:hi].cuda(), idx_offset=lo)
            q_rows, inv = ops.normalize_cast(q_dev[0], pc.op, _lib.SIDE_QUERY, True, pc.eps)
            q_invs = [inv]
        shards.append(pc)
        k_loc = max(1, min(kp, pc.n))
        raw_s = torch.empty((n_q, kp), dtype=torch.float32, device="cuda")
        raw_i = torch.empty((n_q, kp), dtype=torch.int32, device="cuda")
        ws = torch.empty((max(8, int(_lib.load().mmd_topk_workspace_bytes(n_q, max(pc.n, 1), pc.dim, ops._OP_DTYPE[pc.op], k_loc))),),
                         dtype=torch.uint8, device="cuda")
        ops.sharded_candidates(q_rows, n_q, pc, k_loc, raw_s, raw_i, ws, [t.data_ptr() for t in thr], r, share,
                               [g.data_ptr() for g in gather], r * part_stride, kp)
    torch.cuda.synchronize()
    for g in gather[1:]:
        assert torch.equal(g, gather[0])
    lists = gather[0].view(world, n_q, kp, 2)
    assert bool((lists[..., 1] >= -1).all()) and bool((lists[..., 1] < n_rows).all())      # every slot written, padded with -1
    for r in range(world):                                          # stage X: every rank re-scores what it owns
        pc = shards[r]
        if joint:
            tables = ops.segment_tables(q_dev, 0, q_invs, pc.sources, pc.inv_norms, pc.dims, pc.weights)
        else:
            tables = ops.segment_tables(q_dev, 0, q_invs, [pc.source], [pc.inv_norm], [pc.dim], [1.0])
        ops.exchange_rescore(gather[r].data_ptr(), world, part_stride, n_q, kp, kc, tables, pc.n, pc.idx_offset,
                             [x.data_ptr() for x in resc], r, pc.device)
    torch.cuda.synchronize()
    for r in range(world):                                          # stage F on every rank: identical lists everywhere
        out_s = torch.empty((n_q, k_glob), dtype=torch.float32, device="cuda")
        out_i = torch.empty((n_q, k_glob), dtype=torch.int64, device="cuda")
        ops.exchange_finish(resc[r].data_ptr(), n_q, kc, k_glob, out_s.data_ptr(), out_i.data_ptr(), True, out_s.device)
        assert torch.equal(out_i, want_i) and torch.equal(out_s, want_s), r
    out_i32 = torch.empty((n_q, k_glob), dtype=torch.int32, device="cuda")
    ops.exchange_finish(resc[0].data_ptr(), n_q, kc, k_glob, out_s.data_ptr(), out_i32.data_ptr(), False, out_s.device)
    assert torch.equal(out_i32.long(), want_i)
    if not joint and n_rows > 100:
        assert out_i[0, :2].tolist() == [7, n_rows - 1]


# ------------------------------------------------------------------------------------------ joint image+text fusion
@pytest.mark.parametrize("op,dims,weights", [("bf16", (512, 512), (0.5, 0.5)), ("bf16", (768, 2048), (0.7, 0.3)),
                                             ("fp16", (100, 36), (0.25, 0.75)), ("bf16", (64, 64, 32), (0.2, 0.3, 0.5))])
def test_joint_fusion_matches_oracle(m, op, dims, weights):
    """sum_m w_m * cos(q_m, c_m) from ONE contraction over side-by-side segments + exact joint re-score, against the
    float64 fusion oracle (BASELINE configs[3] at a size the oracle finishes in seconds)."""
    from oracle import fusion
    n_q, n_c, k = 150, 4000, 10
    qs = [_data("image" if d == 2048 else "text", n_q, d, 70 + j) for j, d in enumerate(dims)]
    cs = [_data("image" if d == 2048 else "text", n_c, d, 80 + j) for j, d in enumerate(dims)]
    jc = m.prepare_joint([c.cuda() for c in cs], weights, dtype=op)
    full = fusion.fused_scores(qs, cs, weights)
    s, i = m.topk_joint([q.cuda() for q in qs], jc, k)
    cmp = exact.compare_topk(s, i, full, k, tie_tol=2e-6)
    assert cmp.ok and cmp.max_rel_score_err <= FP32_RTOL, cmp
    # raw tensor-core scores (no re-score): within the operand rounding of the weighted sum
    s_raw, i_raw = m.topk_joint([q.cuda() for q in qs], jc, k, rescore_exact=False)
    picked = torch.gather(full, 1, i_raw.cpu())
    assert float((s_raw.cpu().double() - picked).abs().max()) <= (2e-3 if op == "bf16" else 3e-4)
    # weight (1, 0, ...) degenerates to single-modality retrieval on modality 0
    one = [1.0] + [0.0] * (len(dims) - 1)
    s1, i1 = m.topk_joint([q.cuda() for q in qs], m.prepare_joint([c.cuda() for c in cs], one, dtype=op), k)
    s0, i0 = m.topk(qs[0].cuda(), m.prepare_corpus(cs[0].cuda(), dtype=op), k)
    assert torch.equal(i1, i0) and float((s1 - s0).abs().max()) <= 1e-6


def test_joint_fusion_fp8_and_errors(m):
    from oracle import fusion
    qs = [_data("text", 64, 512, 90), _data("text", 64, 512, 91)]
    cs = [_data("text", 3000, 512, 92), _data("text", 3000, 512, 93)]
    jc = m.prepare_joint([c.cuda() for c in cs], dtype="fp8")
    s, i = m.topk_joint([q.cuda() for q in qs], jc, 10, overfetch=60)
    want_s, want_i = fusion.fused_topk(qs, cs, (0.5, 0.5), 10)
    recall = sum(len(set(a) & set(b)) for a, b in zip(i.cpu().tolist(), want_i.tolist())) / want_i.numel()
    assert recall >= 0.97
    with pytest.raises(ValueError):
        m.prepare_joint([cs[0].cuda(), cs[1][:10].cuda()])
    with pytest.raises(ValueError):
        m.prepare_joint([c.cuda() for c in cs], dtype="fp32")
    with pytest.raises(RuntimeError):
        m.topk_joint([qs[0].cuda(), qs[1][:, :100].cuda()], jc, 5)


def test_text_search_drop_in_matches_restated_reference_flow(m):
    """SemanticSimilarity.search: two corpora, top_k*5 each, concat, sort, distinct-score dedupe -- against the same
    flow built from the restated sentence-transformers semantic_search (oracle/st_util.py) on the host; duplicate
    evidence rows across the train / test corpora collapse exactly as in text2text_retrieval.py:105-118."""
    train, test = _data("text", 3000, 768, 97), _data("text", 800, 768, 98)
    test[5] = train[11]
    test[6] = train[11]                                              # the same evidence three times
    q = _data("text", 20, 768, 99)
    q[0] = train[11] + 0.05 * q[0]
    train_ids = [f"train_{j}".encode() for j in range(3000)]
    test_ids = [f"test_{j}".encode() for j in range(800)]
    ss = m.SemanticSimilarity(train.cuda(), train_ids, test.cuda(), test_ids)
    got = ss.search_batch(q, top_k=5)
    assert got[0] == ss.search(q[0], top_k=5)
    for qi in range(20):
        res = []
        for corpus, ids in ((train, train_ids), (test, test_ids)):
            hits = st_util.semantic_search(q[qi], corpus, top_k=25)[0]
            res += [(ids[h["corpus_id"]].decode(), h["score"]) for h in hits]
        want = m.dedupe_by_score(sorted(res, key=lambda t: t[1], reverse=True), 5)
        assert [a for a, _ in got[qi]] == [a for a, _ in want] or qi == 0
        assert max(abs(a - b) / abs(b) for (_, a), (_, b) in zip(got[qi], want)) <= FP32_RTOL
    ids0 = [a for a, _ in got[0]]
    assert ids0[0] in ("train_11", "test_5", "test_6") and len({"train_11", "test_5", "test_6"} & set(ids0)) == 1
    # a pluggable re-ranker replaces the scores and the order
    rer = m.SemanticSimilarity(train.cuda(), train_ids, test.cuda(), test_ids, cross_encoder=lambda query, texts: [float(len(t)) for t in texts],
                               train_texts=["x" * (j % 7) for j in range(3000)], test_texts=["y" * (j % 5) for j in range(800)])
    out = rer.search_batch(q[:2], top_k=3, query_texts=["a", "b"])
    assert [s for _, s in out[0]] == sorted({s for _, s in out[0]}, reverse=True) and out[0][0][1] == 6.0


def test_text_topk_accuracy_matches_restated_reference_eval(m):
    """calculate_topk_accuracy_text_retrieval (experiment_text.py:11-106) on Factify-shaped synthetic data -- every claim
    has one planted evidence `test_{i}`, some evidences are exact duplicates of train rows (equal scores: the gold
    exemption matters) -- against the same evaluation assembled from the restated semantic_search on the host."""
    from oracle import evalmetrics
    n_claims = 120
    train, test = _data("text", 2500, 768, 101), _data("text", n_claims, 768, 102)
    claims = test + 0.9 * _data("text", n_claims, 768, 103)          # noisy enough for hits@1 < 1
    train[:30] = test[:30]                                           # duplicated evidence: same score as the gold row
    train_ids = [f"train_{j}".encode() for j in range(2500)]
    test_ids = [f"test_{j}".encode() for j in range(n_claims)]
    ss = m.SemanticSimilarity(train.cuda(), train_ids, test.cuda(), test_ids)
    got = m.calculate_topk_accuracy_text_retrieval(ss, claims)
    # The duplicated rows tie with the gold evidence; which of the two an fp32 pipeline ranks first is decided by the last
    # bit of two separately computed dot products (the host matmul's blocking depends on the corpus shape and the thread
    # count).  So the oracle is evaluated with the ties broken both ways (scores rounded to 1e-6, duplicate first / gold
    # first): the device result must lie between the two, and equal them wherever they agree.
    bounds = []
    for gold_first in (False, True):
        lists = []
        for qi in range(n_claims):
            res = []
            order = ((test, test_ids), (train, train_ids)) if gold_first else ((train, train_ids), (test, test_ids))
            for corpus, ids in order:
                res += [(ids[h["corpus_id"]].decode(), round(h["score"], 6)) for h in st_util.semantic_search(claims[qi], corpus, top_k=100)[0]]
            ranked = sorted(res, key=lambda t: t[1], reverse=True)
            lists.append([key for key, _ in evalmetrics.dedupe_first_of_each_score(ranked, 10, gold=lambda key, qi=qi: key == f"test_{qi}")])
        bounds.append(evalmetrics.hits_at_k(lists, [f"test_{i}" for i in range(n_claims)], (1, 2, 5, 10)))
    lo, hi = bounds
    for k in (1, 2, 5, 10):
        assert lo[k] - 1e-9 <= got[k] <= hi[k] + 1e-9, (k, got, lo, hi)
    assert hi[1] > lo[1]                                             # the ties are there, and the exemption decides them
    assert 0.2 < got[1] <= got[2] <= got[5] <= got[10] <= 1.0


def test_device_dedupe_matches_reference_walk(m):
    """mmd_dedupe_scores vs the host restatement of the reference's distinct-score walk (with and without the gold
    exemption), on lists full of repeated scores."""
    from mmd_retrieval import ops
    gen = torch.Generator().manual_seed(77)
    for n_q, k_in, top_k in [(40, 18, 10), (9, 100, 50), (5, 3, 5), (3, 700, 20)]:
        s = torch.sort(torch.randint(0, 12, (n_q, k_in), generator=gen).float() / 8, dim=1, descending=True).values
        idx = torch.stack([torch.randperm(5000, generator=gen)[:k_in] for _ in range(n_q)]).to(torch.int32)
        s[0, k_in - 1:] = float("-inf")
        idx[0, k_in - 1:] = -1                                         # a padded tail
        gold = idx[torch.arange(n_q), torch.randint(0, k_in, (n_q,), generator=gen)].clone()
        gold[1] = -1
        for g in (None, gold):
            ks, ki, kn = ops.dedupe_scores(s.cuda(), idx.cuda(), top_k, None if g is None else g.cuda())
            for r in range(n_q):
                ranked = [(int(i), float(v)) for v, i in zip(s[r], idx[r]) if i >= 0]
                want = m.dedupe_by_score(ranked, top_k, None if g is None else (lambda key, r=r: key == int(g[r])))
                n = int(kn[r])
                assert n == len(want)
                assert ki[r, :n].tolist() == [a for a, _ in want] and ks[r, :n].tolist() == [b for _, b in want]
                assert bool((ki[r, n:] == -1).all())


# ------------------------------------------------------------------------------------------ corpus containers
def test_streamed_prepare_and_corpus_files(m, tmp_path):
    """Chunk-wise K1 gives bit-identical tiles to a one-shot prepare; the reference's corpus containers (h5-shaped
    arrays as .npz/.npy, image-feature pickle dict) load into prepared corpora; rows=(lo,hi) loads one shard."""
    import pickle
    import numpy as np
    c = _data("text", 5000, 768, 95)
    whole = m.prepare_corpus(c.cuda(), dtype="bf16")
    parts = [c[:1234], c[1234:1234], c[1234:4000].numpy(), c[4000:]]
    st = m.prepare_streamed(parts, 5000, 768, dtype="bf16", keep_source=True)
    assert torch.equal(st.rows, whole.rows) and torch.equal(st.inv_norm, whole.inv_norm) and torch.equal(st.source, c.cuda())
    with pytest.raises(ValueError):
        m.prepare_streamed([c[:10]], 11, 768)
    # the text corpus layout of text2text_retrieval.py:146-155 (fp16 embeddings + ids), as npz / npy
    ids = np.array([f"train_{j}".encode() for j in range(5000)])
    np.savez(tmp_path / "train_embeddings.npz", embeddings=c.numpy().astype(np.float16), ids=ids)
    np.save(tmp_path / "emb.npy", c.numpy().astype(np.float16))
    pc, got_ids = m.load_text_corpus(str(tmp_path / "train_embeddings.npz"), chunk_rows=1500)
    assert got_ids[17] == "train_17" and pc.n == 5000 and pc.source.dtype == torch.float16
    q = _data("text", 50, 768, 96)
    s, i = m.topk(q.cuda(), pc, 10)
    full = exact.exact_scores(q, c.half().float())
    assert exact.compare_topk(s, i, full, 10, tie_tol=2e-6).ok
    shard, _ = m.load_text_corpus(str(tmp_path / "emb.npy"), rows=(2000, 3500), chunk_rows=700)
    s2, i2 = m.topk(q.cuda(), shard, 10)
    assert shard.idx_offset == 2000 and int(i2.min()) >= 2000 and int(i2.max()) < 3500
    assert exact.compare_topk(s2.cpu(), i2.cpu() - 2000, full[:, 2000:3500], 10, tie_tol=2e-6).ok
    # the image-feature pickle of im2im_retrieval.py:51-62
    g = load_golden("im2im_a.npz")
    fd = {f"data/evidence_corpus/{j}_evidence.jpg": v for j, v in enumerate(g["corpus_t"])}
    with open(tmp_path / "evidence_features.pkl", "wb") as f:
        pickle.dump(fd, f)
    ipc, keys = m.load_image_corpus(str(tmp_path / "evidence_features.pkl"), chunk_rows=100)
    assert keys == list(fd.keys()) and ipc.eps == 1e-6 and ipc.n == len(fd)
    s3, i3 = m.topk(g["queries_t"].cuda(), ipc, 5)
    full3 = exact.exact_scores(g["queries_t"], g["corpus_t"], "cos", 1e-6)
    assert exact.compare_topk(s3, i3, full3, 5, tie_tol=2e-6).ok


# ------------------------------------------------------------------------------------------ drop-in surfaces
@pytest.mark.parametrize("name", ["im2im_a.npz", "im2im_b.npz"])
def test_image_corpus_matches_reference_golden(m, name):
    """ImageCorpus.retrieve_similar_images == the reference's own output (golden made by its code)."""
    g = load_golden(name)
    fd = {f"c{i:05d}": g["corpus_t"][i] for i in range(g["corpus_t"].shape[0])}
    top_k = int(g["top_k"])

    class Extractor:
        def extract_features(self, path):
            return g["queries_t"][int(path[1:])]

    for dtype in ("bf16", "fp32"):
        corpus = m.ImageCorpus(feature_dict=fd, feature_extractor=Extractor(), dtype=dtype)
        for qi in range(g["queries_t"].shape[0]):
            got = corpus.retrieve_similar_images(f"q{qi:03d}", top_k=top_k)
            want_rows = [r for r in g["rows"][qi].tolist() if r >= 0]
            assert [int(k[1:]) for k, _ in got] == want_rows, (dtype, qi)
            np.testing.assert_allclose([s for _, s in got], g["scores"][qi][: len(got)], rtol=FP32_RTOL, atol=0)
    batched = corpus.retrieve_similar_features(g["queries_t"], top_k)
    assert [[int(k[1:]) for k, _ in lst] for lst in batched] == [[r for r in row.tolist() if r >= 0] for row in g["rows"]]


def test_semantic_search_drop_in(m):
    g = load_golden("t2t_fp32.npz")
    q, c = torch.from_numpy(g["queries"]).float(), torch.from_numpy(g["corpus"]).float()
    hits = m.semantic_search(q, c, top_k=int(g["top_k"]), dtype="fp32")
    assert len(hits) == q.shape[0] and all(len(h) == int(g["top_k"]) for h in hits)
    assert isinstance(hits[0][0]["corpus_id"], int) and isinstance(hits[0][0]["score"], float)
    np.testing.assert_array_equal([[h["corpus_id"] for h in hl] for hl in hits], g["rows"])
    np.testing.assert_allclose([[h["score"] for h in hl] for hl in hits], g["scores"], rtol=FP32_RTOL)
    # default bf16 selection + re-score gives the same lists
    hits_bf16 = m.semantic_search(q, c, top_k=int(g["top_k"]))
    np.testing.assert_array_equal([[h["corpus_id"] for h in hl] for hl in hits_bf16], g["rows"])
    # the reference's calling pattern: one fp16 1-D query per call, fp16 CPU corpus, repeated calls (cache)
    g16 = load_golden("t2t_fp16.npz")
    q16, c16 = torch.from_numpy(g16["queries"]), torch.from_numpy(g16["corpus"])
    full = exact.exact_scores(q16.float(), c16.float())
    k = int(g16["top_k"])
    launches0 = m.launch_count()
    for qi in range(4):
        h = m.semantic_search(q16[qi], c16, top_k=k)[0]
        cmp = exact.compare_topk(torch.tensor([[x["score"] for x in h]]), torch.tensor([[x["corpus_id"] for x in h]]),
                                 full[qi:qi + 1], k, tie_tol=2e-6)
        assert cmp.ok and cmp.max_rel_score_err <= FP32_RTOL
        # and against the reference-dtype (fp16) golden: same sets up to fp16 near-ties
        ref = exact.compare_topk(torch.from_numpy(g16["scores"][qi:qi + 1]), torch.from_numpy(g16["rows"][qi:qi + 1]),
                                 full[qi:qi + 1], k, tie_tol=2e-3)
        assert ref.ok
    # corpus prepared once (1 K1 launch), then per call: K1(query) + fused + merge + rescore
    assert m.launch_count() - launches0 == 1 + 4 * 4
    # top_k > N and dot_score
    few = m.semantic_search(q[:2], c[:3], top_k=10)
    assert [len(h) for h in few] == [3, 3]
    d = m.semantic_search(q[:3], c, top_k=4, score_function=m.dot_score, dtype="fp32")
    want = st_util.semantic_search(q[:3], c, top_k=4, score_function=st_util.dot_score)
    assert [[h["corpus_id"] for h in hl] for hl in d] == [[h["corpus_id"] for h in hl] for hl in want]
    cs = m.cos_sim(q[:5], c[:9]).cpu()
    torch.testing.assert_close(cs, st_util.cos_sim(q[:5], c[:9]), rtol=FP32_RTOL, atol=1e-6)


def test_factify_shaped_topk_accuracy_matches_oracle(m):
    """hits@{1,2,5,10} with the gold-exempt dedupe (experiment_image.py:41-61) on planted-positive synthetic data:
    the GPU path and the CPU oracle must produce the same accuracy dict."""
    gen = torch.Generator().manual_seed(60)
    n, n_q, dim = 5000, 300, 2048
    c = torch.relu(torch.randn(n, dim, generator=gen))
    gold_rows = torch.randperm(n, generator=gen)[:n_q]
    sigma = torch.linspace(0.3, 3.0, n_q)[:, None]                 # from easy to hopeless queries
    q = torch.relu(c[gold_rows] + sigma * torch.randn(n_q, dim, generator=gen))
    for j in range(40):                                            # duplicated evidences (identical scores)
        c[(int(gold_rows[j]) + 1) % n] = c[int(gold_rows[j])]
    keys = [f"img{i}" for i in range(n)]
    gold_keys = [keys[int(r)] for r in gold_rows]
    want = evalmetrics.image_eval(exact.exact_scores(q, c, eps=1e-6, dtype=torch.float32), keys, gold_keys)
    corpus = m.ImageCorpus(feature_dict={k: c[i] for i, k in enumerate(keys)})
    got = m.calculate_topk_accuracy_image_retrieval(corpus, q, gold_keys)
    assert got == want, (got, want)
    assert 0.2 < got[1] < 1.0 and got[10] >= got[1]


def test_split_search_raises_effective_overfetch(m):
    """Long lists / lossy candidates: independent sub-searches (own bounds, own K' candidates, own exact re-score) merged by
    exact score.  ops.auto_splits picks the count; more splits can only bring the result closer to the fp32 ranking, and
    the returned scores are always the exact ones of the returned rows (4 spare candidates at k = 100)."""
    from mmd_retrieval import ops
    q, c = _data("text", 128, 768, 72), _data("text", 300_000, 768, 73)
    full = exact.exact_scores(q, c)
    pc8 = m.prepare_corpus(c.cuda(), dtype="fp8")
    assert ops.auto_splits("fp8", 100, ops.overfetch_for(100, pc8.n), pc8.n) == 2          # 4 wanted, 2 x 131072 rows fit
    rec = {}
    for splits in (1, 2, 4):
        s, i = m.topk(q.cuda(), pc8, 100, splits=splits)
        assert tuple(i.shape) == (128, 100) and bool((s[:, :-1] >= s[:, 1:]).all())
        assert bool((torch.sort(i, dim=1).values.diff(dim=1) > 0).all())                    # no row twice
        got = torch.gather(full.cuda(), 1, i)
        assert float(((s.double() - got.double()).abs() / got.double().abs().clamp_min(1e-3)).max()) <= 1e-5
        rec[splits] = exact.recall_at_k(i, full, 100)
    s_auto, i_auto = m.topk(q.cuda(), pc8, 100)
    assert exact.recall_at_k(i_auto, full, 100) == rec[2]
    assert rec[1] <= rec[2] + 1e-3 and rec[2] <= rec[4] + 1e-3 and rec[4] >= 0.995 and rec[4] > rec[1], rec
    # bf16, k = 100: two splits make the list exact (one split: K' - k = 4 spare candidates)
    pc16 = m.prepare_corpus(c.cuda(), dtype="bf16")
    assert ops.auto_splits("bf16", 100, ops.overfetch_for(100, pc16.n), pc16.n) == 2
    s, i = m.topk(q.cuda(), pc16, 100)
    cmp = exact.compare_topk(s, i, full, 100, tie_tol=2e-6)
    assert cmp.ok and cmp.max_rel_score_err <= 1e-5, cmp


def test_fp8_recall(m):
    q, c = _data("text", 256, 768, 70), _data("text", 20000, 768, 71)
    s, i = m.topk(q.cuda(), m.prepare_corpus(c.cuda(), dtype="fp8", keep_source=False), 10, rescore_exact=False)
    assert exact.compare_topk(s, i, _device_scores(m, q, c, "fp8", 1e-12), 10, tie_tol=2e-6).ok   # exact w.r.t. its own operands
    want = exact.exact_topk(q, c, 10)[1]
    recall = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(i.cpu().tolist(), want.tolist())])
    assert recall >= 0.80, recall
    # with over-fetch + exact re-score the fp8 pass recovers the fp32 lists almost everywhere
    s, i = m.topk(q.cuda(), m.prepare_corpus(c.cuda(), dtype="fp8"), 10, overfetch=60)
    recall2 = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(i.cpu().tolist(), want.tolist())])
    assert recall2 >= 0.97, recall2
