"""GPU, >= 2 devices on one box: the row-sharded path with NCCL (one process per GPU) against the CPU oracle.
Skipped on a single-GPU box (tests/test_gpu_parity.py::test_sharded_equals_unsharded covers the shard-invariance
of the kernels there, tests/test_sharded_gloo.py the collective plumbing)."""
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


def _worker(rank, world, port, out):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import mmd_retrieval as m
    from oracle import exact
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        gen = torch.Generator().manual_seed(99)
        corpus = torch.randn(30011, 768, generator=gen)
        corpus[30010] = corpus[5]                              # duplicate rows on different ranks
        queries = torch.randn(300, 768, generator=gen)
        queries[0] = corpus[5] * 2
        sc = m.ShardedCorpus.from_full(corpus.cuda())
        s, i = sc.topk(queries.cuda(), 10)
        full = exact.exact_scores(queries, corpus)
        cmp = exact.compare_topk(s, i, full, 10, tie_tol=2e-6)
        assert cmp.ok and cmp.max_rel_score_err <= 1e-5, cmp
        assert i[0, :2].tolist() == [5, 30010]
        # host queries in, the public call moves them; repeated calls reuse the double-buffered exchange
        for _ in range(3):
            s2, i2 = sc.topk(queries, 10)
            assert torch.equal(i2, i) and torch.equal(s2, s)
        # both exchange implementations and both stage orders give the same lists
        for kw in ({"exchange": "nccl"}, {"rescore": "local"}, {"exchange": "nccl", "rescore": "local"}, {"share_thresholds": False}):
            other = m.ShardedCorpus.from_full(corpus.cuda(), **kw)
            s3, i3 = other.topk(queries.cuda(), 10)
            assert other.exchange == kw.get("exchange", sc.exchange)
            cmp3 = exact.compare_topk(s3, i3, full, 10, tie_tol=2e-6)
            assert cmp3.ok and cmp3.max_rel_score_err <= 1e-5, (kw, cmp3)
            assert torch.equal(i3, i) and torch.equal(s3, s), kw
        # host-resident queries: every rank uploads a 1/world slice and the slices are all-gathered (odd sizes too)
        for n_q in (300, 299, 3, 1):
            up = sc.upload_queries(queries[:n_q])
            assert up.is_cuda and torch.equal(up.cpu(), queries[:n_q])
        # the whole step replayed from CUDA graphs (one per exchange-buffer parity), new queries copied in each time
        graphed = sc.capture(queries.cuda(), 10)
        for rep in range(4):
            qq = queries if rep % 2 == 0 else queries.flip(0)
            sg, ig = graphed(qq.cuda())
            want_s, want_i = (s, i) if rep % 2 == 0 else (s.flip(0), i.flip(0))
            assert torch.equal(ig, want_i) and torch.equal(sg, want_s), rep
        # fewer corpus rows than ranks * k: short and empty local lists are padded with (-inf, -1)
        tiny = m.ShardedCorpus.from_full(corpus[:5].cuda())
        for _ in range(5):          # (repeated: a shard's bound must not prune rows the global list needs -- timing dependent)
            s4, i4 = tiny.topk(queries.cuda(), 10)
            assert tuple(i4.shape) == (300, 5) and torch.equal(i4.cpu(), exact.exact_topk(queries, corpus[:5], 5)[1])
        small = m.ShardedCorpus.from_full(corpus[:41].cuda())
        for _ in range(3):
            s5, i5 = small.topk(queries.cuda(), 10)
            assert exact.compare_topk(s5, i5, exact.exact_scores(queries, corpus[:41]), 10, tie_tol=2e-6).ok
        out[rank] = sc.exchange
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_topk_nccl():
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out.get(r) in ("peer", "nccl") for r in range(world)), dict(out)
    print("exchange used:", dict(out))
