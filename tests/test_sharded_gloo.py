"""CPU, world_size 2, gloo: the row-sharding / global-offset / all-gather / merge plumbing of
mmd_retrieval.sharded.ShardedCorpus.  The per-shard top-K and the merge are injected from the CPU oracle
(test-only); on the GPU box the same class runs with the CUDA ops (tests/test_gpu_multi.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_local_topk(queries, shard, k):
    from oracle import exact
    rows, start = shard
    if rows.shape[0] == 0:
        return torch.empty(queries.shape[0], 0), torch.empty(queries.shape[0], 0, dtype=torch.int32)
    v, i = exact.exact_topk(queries, rows, k)
    return v.float(), (i + start).to(torch.int32)


def _oracle_merge(scores, idx, k):
    parts, n_q, k_in = scores.shape
    s = scores.permute(1, 0, 2).reshape(n_q, parts * k_in).double()
    i = idx.permute(1, 0, 2).reshape(n_q, parts * k_in).long()
    # order by (score desc, row asc); empty slots (row -1) last
    key = torch.where(i < 0, torch.full_like(s, float("-inf")), s)
    order = torch.argsort(i, dim=1, stable=True)
    key, i = torch.gather(key, 1, order), torch.gather(i, 1, order)
    order = torch.argsort(key, dim=1, descending=True, stable=True)
    return torch.gather(key, 1, order)[:, :k].float(), torch.gather(i, 1, order)[:, :k].to(torch.int32)


def _worker(rank, world, port, n_rows, k, out):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from mmd_retrieval.sharded import ShardedCorpus
    from oracle import exact
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(1234)
        corpus = torch.randn(n_rows, 32, generator=gen)
        if n_rows > 20:
            corpus[n_rows - 1] = corpus[0]             # a duplicate pair straddling the two shards (tie rule)
        queries = torch.randn(9, 32, generator=gen)
        sc = ShardedCorpus.from_full(corpus, local_topk=_oracle_local_topk, merge=_oracle_merge,
                                     prepare=lambda rows, start: (rows, start))
        s, i = sc.topk(queries, k)
        want_s, want_i = exact.exact_topk(queries, corpus, k)
        assert i.dtype == torch.int64 and tuple(i.shape) == (9, min(k, n_rows))
        assert torch.equal(i, want_i), (rank, i, want_i)
        assert torch.allclose(s.double(), want_s, atol=1e-6)
        gathered = [torch.zeros_like(i) for _ in range(world)]
        dist.all_gather(gathered, i)
        assert all(torch.equal(g, i) for g in gathered)   # every rank ends with the same global list
        out[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,k", [(101, 10), (3, 10), (1, 4)])
def test_sharded_topk_world2_gloo(n_rows, k):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_rows, k, out), nprocs=world, join=True)
    assert all(out.get(r) for r in range(world))
