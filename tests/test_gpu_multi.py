"""GPU, >= 2 devices on one box: the row-sharded path (one process per GPU, peer-memory exchange with flags, or NCCL)
against the CPU oracle.  Skipped on a single-GPU box (tests/test_gpu_parity.py covers the shard-invariance of the kernels
and the exchange stages X / F there, tests/test_sharded_gloo.py the collective plumbing).

    python -m pytest tests/test_gpu_multi.py -m gpu            # min(device_count, 4) ranks
    MMD_TEST_WORLD=8 python -m pytest tests/test_gpu_multi.py -m gpu
"""
import os
import socket

import pytest
import torch

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(rank, world, port):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    return dist


def _worker_text(rank, world, port, out):
    dist = _setup(rank, world, port)
    import mmd_retrieval as m
    from oracle import exact
    try:
        gen = torch.Generator().manual_seed(99)
        corpus = torch.randn(30011, 768, generator=gen)
        corpus[30010] = corpus[5]                              # duplicate rows on different ranks
        queries = torch.randn(300, 768, generator=gen)
        queries[0] = corpus[5] * 2
        sc = m.ShardedCorpus.from_full(corpus.cuda())
        s, i = sc.topk(queries.cuda(), 10)
        assert sc.exchange == os.environ.get("MMD_EXPECT_EXCHANGE", sc.exchange)
        full = exact.exact_scores(queries, corpus)
        cmp = exact.compare_topk(s, i, full, 10, tie_tol=2e-6)
        assert cmp.ok and cmp.max_rel_score_err <= 1e-5, cmp
        assert i[0, :2].tolist() == [5, 30010]
        # host queries in, the public call moves them; repeated calls walk around the ring of exchange buffers
        for _ in range(7):
            s2, i2 = sc.topk(queries, 10)
            assert torch.equal(i2, i) and torch.equal(s2, s)
        # batch sizes L, L, S, L: a smaller batch must not leave stale pruning bounds behind
        for n_q in (300, 300, 7, 300, 1, 300, 257, 300):
            sv, iv = sc.topk(queries[:n_q].cuda(), 10)
            assert torch.equal(iv, i[:n_q]) and torch.equal(sv, s[:n_q]), n_q
        # ... the same with DIFFERENT queries of different batch sizes in a row (stale bounds of other queries would prune)
        gq = torch.Generator().manual_seed(7)
        for n_q in (300, 41, 300, 5):
            qq = torch.randn(n_q, 768, generator=gq)
            sv, iv = sc.topk(qq.cuda(), 10)
            assert exact.compare_topk(sv, iv, exact.exact_scores(qq, corpus), 10, tie_tol=2e-6).ok, n_q
        # every exchange implementation, stage order, threshold mode and sub-batch split gives the same lists
        for kw in ({"exchange": "nccl"}, {"rescore": "local"}, {"exchange": "nccl", "rescore": "local"}, {"share_thresholds": False},
                   {"share_thresholds": "always"}, {"sub_batches": 1}, {"sub_batches": 2}, {"sub_batches": 5}):
            other = m.ShardedCorpus.from_full(corpus.cuda(), **kw)
            s3, i3 = other.topk(queries.cuda(), 10)
            if kw.get("share_thresholds") == "always":           # overlapped batches WITH cross-GPU bounds (the stream's default is without)
                for sv, iv in other.topk_stream(iter([queries.cuda(), queries[:77].cuda(), queries.cuda()]), 10):
                    n = sv.shape[0]
                    assert torch.equal(iv, i[:n]) and torch.equal(sv, s[:n]), n
            if "exchange" in kw or "rescore" in kw:
                assert other.exchange == "nccl", (kw, other.exchange)
            cmp3 = exact.compare_topk(s3, i3, full, 10, tie_tol=2e-6)
            assert cmp3.ok and cmp3.max_rel_score_err <= 1e-5, (kw, cmp3)
            assert torch.equal(i3, i) and torch.equal(s3, s), kw
        # host-resident queries: every rank uploads a 1/world slice and the slices are all-gathered (odd sizes too)
        for n_q in (300, 299, 3, 1):
            up = sc.upload_queries(queries[:n_q])[0]
            assert up.is_cuda and torch.equal(up.cpu(), queries[:n_q])
        # pipelined stream of batches (device and host results), sizes varying
        sizes = [300, 128, 300, 1, 77, 300]
        batches = [queries[:n] for n in sizes]
        for to_host in (False, True):
            got = list(sc.topk_stream(iter(batches), 10, to_host=to_host))
            assert len(got) == len(sizes)
            for n, (sv, iv) in zip(sizes, got):
                assert sv.is_cuda != to_host
                assert torch.equal(iv.cpu(), i[:n].cpu()) and torch.equal(sv.cpu(), s[:n].cpu()), (to_host, n)
        # the whole step replayed from a CUDA graph, new queries copied in each time; eager calls in between
        graphed = sc.capture(queries.cuda(), 10)
        for rep in range(5):
            qq = queries if rep % 2 == 0 else queries.flip(0)
            sg, ig = graphed(qq.cuda())
            want_s, want_i = (s, i) if rep % 2 == 0 else (s.flip(0), i.flip(0))
            assert torch.equal(ig, want_i) and torch.equal(sg, want_s), rep
            if rep == 2:
                s2, i2 = sc.topk(queries.cuda(), 10)
                assert torch.equal(i2, i) and torch.equal(s2, s)
        # top-100 (two-register-wide lists in the exchange stages)
        s100, i100 = sc.topk(queries.cuda(), 100)
        cmp100 = exact.compare_topk(s100, i100, full, 100, tie_tol=2e-6)
        assert cmp100.max_rel_score_err <= 1e-5 and cmp100.violations <= 3, cmp100   # K' - K = 4 at K = 100: bf16 near-ties at rank 100
        assert exact.recall_at_k(i100, full, 100) >= 0.999
        out[rank] = sc.exchange
    finally:
        dist.destroy_process_group()


def _worker_small(rank, world, port, out):
    """Small and uneven shards: fewer rows than ranks * K', empty shards, short lists padded with (-inf, -1); the shared
    pruning thresholds must never drop a row the global list needs (timing dependent => repeated, randomised)."""
    dist = _setup(rank, world, port)
    import mmd_retrieval as m
    from mmd_retrieval import ops
    from oracle import exact
    try:
        gen = torch.Generator().manual_seed(5)
        corpus = torch.randn(4096, 256, generator=gen)
        kp = ops.overfetch_for(10, 10 ** 6)
        sizes = sorted({1, 2, 5, 41, world * kp - 1, world * kp, world * kp + 1, 199, 1000, 4096})
        for n_rows in sizes:
            sc = m.ShardedCorpus.from_full(corpus[:n_rows].cuda())
            for it in range(20):
                qq = torch.randn(64 + it, 256, generator=gen)
                sv, iv = sc.topk(qq.cuda(), 10)
                want = min(10, n_rows)
                assert tuple(iv.shape) == (64 + it, want), (n_rows, iv.shape)
                cmp = exact.compare_topk(sv, iv, exact.exact_scores(qq, corpus[:n_rows]), want, tie_tol=2e-6)
                # (256-d random pairs score ~0.06 and sometimes ~0.001: fp32 rounding is then > 1e-5 RELATIVE; bound it absolutely)
                assert cmp.ok and (cmp.max_rel_score_err <= 1e-5 or cmp.max_score_err <= 2e-7), (n_rows, it, cmp)
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


def _worker_joint_fp8(rank, world, port, out):
    """BASELINE configs[3] (joint image+text, fused top-10) and configs[4] (fp8 shard, top-100) at test scale."""
    dist = _setup(rank, world, port)
    import mmd_retrieval as m
    from mmd_retrieval.sharded import shard_bounds
    from oracle import exact, fusion
    try:
        gen = torch.Generator().manual_seed(11)
        n_rows = 20003
        txt, img = torch.randn(n_rows, 512, generator=gen), torch.randn(n_rows, 512, generator=gen)
        qt, qi = torch.randn(200, 512, generator=gen), torch.randn(200, 512, generator=gen)
        lo, hi = shard_bounds(n_rows, world, rank)
        sc = m.ShardedCorpus.from_joint([txt[lo:hi].cuda(), img[lo:hi].cuda()], n_rows, lo, weights=(0.5, 0.5))
        for _ in range(3):
            s, i = sc.topk([qt.cuda(), qi.cuda()], 10)
        assert sc.exchange == "peer" or os.environ.get("MMD_EXPECT_EXCHANGE") != "peer"
        full = fusion.fused_scores([qt, qi], [txt, img], (0.5, 0.5))
        cmp = exact.compare_topk(s, i, full, 10, tie_tol=2e-6)
        assert cmp.ok and cmp.max_rel_score_err <= 1e-5, cmp
        # the NCCL variant of the joint path gives the same lists
        other = m.ShardedCorpus.from_joint([txt[lo:hi].cuda(), img[lo:hi].cuda()], n_rows, lo, weights=(0.5, 0.5), exchange="nccl")
        s2, i2 = other.topk([qt.cuda(), qi.cuda()], 10)
        assert torch.equal(i2, i) and torch.allclose(s2, s, rtol=1e-6, atol=1e-7)
        # host-resident joint queries through the stream API
        got = list(sc.topk_stream(iter([[qt, qi], [qt[:50], qi[:50]]]), 10, to_host=True))
        assert torch.equal(got[0][1], i.cpu()) and torch.equal(got[1][1], i[:50].cpu())

        # fp8 shard with fp16 source, top-100.  e4m3 candidates are lossy, so the sharded default is the exact re-score PER
        # SHARD (every shard's K' candidates survive: a world-fold over-fetch) and the exact lists are merged: the result is
        # at least as close to the fp32 ranking as the unsharded search of the same rows, and the scores are exact.
        n_rows = 60000
        corpus = torch.randn(n_rows, 768, generator=gen)
        queries = torch.randn(512, 768, generator=gen)
        lo, hi = shard_bounds(n_rows, world, rank)
        shard = m.prepare_streamed(iter([corpus[lo:hi].cuda()]), hi - lo, 768, dtype="fp8", keep_source=torch.float16, idx_offset=lo)
        sc8 = m.ShardedCorpus.from_prepared(shard, n_rows)
        assert sc8.rescore == "local"
        s8, i8 = sc8.topk(queries.cuda(), 100)
        one = m.prepare_streamed(iter([corpus.cuda()]), n_rows, 768, dtype="fp8", keep_source=torch.float16)
        s1, i1 = m.topk(queries.cuda(), one, 100)
        ref = exact.exact_scores(queries, corpus.half().float())
        rec8, rec1 = exact.recall_at_k(i8, ref, 100), exact.recall_at_k(i1, ref, 100)
        assert rec1 >= 0.93 and rec8 >= 0.985 and rec8 >= rec1, (rec8, rec1)   # 4 rows of over-fetch at K = 100 vs world x 104 candidates
        assert bool((s8[:, :-1] >= s8[:, 1:]).all())
        got = torch.gather(ref.cuda(), 1, i8)                                      # returned scores are the exact ones of their rows
        assert float(((s8.double() - got.double()).abs() / got.double().abs().clamp_min(1e-3)).max()) <= 2e-4
        # the global-re-score exchange (one K'-list after the candidate merge) still equals the unsharded search bit for bit
        scg = m.ShardedCorpus.from_prepared(shard, n_rows, rescore="global")
        sg, ig = scg.topk(queries.cuda(), 100)
        assert torch.equal(ig, i1) and torch.equal(sg, s1)
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


def _world():
    n = torch.cuda.device_count()
    want = int(os.environ.get("MMD_TEST_WORLD", "0"))
    return min(n, want) if want > 0 else min(n, 4)


def _spawn(fn):
    import torch.multiprocessing as mp
    world = _world()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(fn, args=(world, _free_port(), out), nprocs=world, join=True)
    return world, dict(out)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_topk_multi_gpu():
    world, out = _spawn(_worker_text)
    assert all(out.get(r) in ("peer", "nccl") for r in range(world)), out
    print("exchange used:", out)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_small_and_uneven_shards_stress():
    world, out = _spawn(_worker_small)
    assert all(out.get(r) == "ok" for r in range(world)), out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_joint_and_fp8_top100():
    world, out = _spawn(_worker_joint_fp8)
    assert all(out.get(r) == "ok" for r in range(world)), out
