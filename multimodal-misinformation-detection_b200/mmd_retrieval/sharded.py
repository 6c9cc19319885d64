"""Row-sharded corpus across the GPUs of one box.

One process per GPU (torchrun).  The corpus rows are split contiguously over the ranks; queries are
replicated; every rank runs the fused top-K on its shard and emits (score f32, GLOBAL row i32) lists; one
exchange of Q*K*8 bytes per rank and a `world`-way merge on the device (K4) give every rank the global top-K.
Scoring never crosses GPUs; the exchange is the path's only communication step.  Two implementations:

  exchange="peer"  the re-score kernel (K5) stores each rank's list directly into EVERY rank's gather buffer
                   (peer-mapped symmetric memory, the stores travel over NVLink 5 / NVSwitch while the kernel
                   is still scoring other queries); one signal-pad barrier follows.  No collective launch.
  exchange="nccl"  K5 fills a local send buffer, one `all_gather_into_tensor` moves it.

"auto" (default) uses "peer" when symmetric memory can be set up on every rank, else "nccl".  The reference has
no counterpart (single process, single device).

`local_topk` / `merge` are injectable so the partition / offset / gather plumbing can be exercised with
gloo on CPU (tests/test_sharded_gloo.py injects the CPU oracle there -- the product path below uses the
CUDA ops and nothing else).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split: the first n_rows % world ranks hold one extra row."""
    base, extra = divmod(n_rows, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _cuda_local_topk(queries, shard, k):
    from . import ops
    return ops.topk(queries, shard, k, index_dtype=torch.int32)


def _cuda_merge(scores, idx, k):
    from . import ops
    return ops.merge_topk(scores, idx, k)


class _PeerExchange:
    """Double-buffered gather buffers [2][world][cap pairs] in symmetric (peer-mapped) memory.

    Step i writes into buffer i % 2 of every rank and then passes ONE barrier.  Reuse is safe without a second
    barrier: a rank can only pass the barrier of step i after every peer has enqueued-and-finished its own stores of
    step i, which in stream order come after that peer's merge of step i-1 -- the last reader of buffer (i+1) % 2."""

    def __init__(self, group, world: int, cap_pairs: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.cap = world, cap_pairs
        self.buf = symm_mem.empty((2, world, cap_pairs, 2), dtype=torch.int32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.step = 0

    def slot(self):
        """(peer base pointers of this step's buffer, local view [world, cap, 2] of it)."""
        b = self.step % 2
        off_bytes = b * self.world * self.cap * 2 * 4
        return [p + off_bytes for p in self.ptrs], self.buf[b]

    def commit(self):
        self.hdl.barrier(channel=0)
        self.step += 1


class ShardedCorpus:
    """This rank's shard of a row-sharded corpus plus the collective that merges local top-K lists."""

    def __init__(self, local_rows, n_total: int, start: int, group=None, dtype: str = "bf16", metric: str = "cos",
                 eps: float = 1e-12, keep_source: bool = True, exchange: str = "auto",
                 local_topk: Optional[Callable] = None, merge: Optional[Callable] = None, prepare: Optional[Callable] = None,
                 _shard=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = int(n_total)
        self.start = int(start)
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
        self._exchange_req = exchange
        self.exchange = "nccl"            # what is actually in use; "peer" once symmetric memory is up on every rank
        self._peer: Optional[_PeerExchange] = None
        self._peer_failed = False
        self._injected = local_topk is not None or merge is not None or prepare is not None
        self._local_topk = local_topk or _cuda_local_topk
        self._merge = merge or _cuda_merge
        if _shard is not None:
            self.shard = _shard
            self.n_local = _shard.n
        elif prepare is None:
            from . import ops
            self.shard = ops.prepare_corpus(local_rows, dtype=dtype, metric=metric, eps=eps, keep_source=keep_source,
                                            idx_offset=self.start)
            self.n_local = self.shard.n
        else:
            self.shard = prepare(local_rows, self.start)
            self.n_local = int(local_rows.shape[0]) if hasattr(local_rows, "shape") else int(local_rows[0].shape[0])

    @classmethod
    def from_full(cls, corpus: torch.Tensor, group=None, **kw) -> "ShardedCorpus":
        """Every rank passes the same full corpus (or a view of it); each keeps only its own rows."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        lo, hi = shard_bounds(corpus.shape[0], world, rank)
        return cls(corpus[lo:hi], corpus.shape[0], lo, group=group, **kw)

    @classmethod
    def from_prepared(cls, shard, n_total: int, group=None, exchange: str = "auto") -> "ShardedCorpus":
        """Wrap this rank's already prepared shard (e.g. corpus_io.prepare_streamed(..., idx_offset=start))."""
        return cls(None, n_total, shard.idx_offset, group=group, exchange=exchange, _shard=shard)

    @classmethod
    def from_joint(cls, local_corpora, n_total: int, start: int, weights=None, group=None, dtype: str = "bf16",
                   metric: str = "cos", eps: float = 1e-12) -> "ShardedCorpus":
        """Row-sharded JOINT (multi-modality) corpus: local_corpora = this rank's rows of every modality; queries are
        passed to topk() as a list with one matrix per modality."""
        from . import joint, ops

        def prepare(rows, first):
            return joint.prepare_joint(rows, weights, dtype=dtype, metric=metric, eps=eps, idx_offset=first)

        def local_topk(queries, shard, k):
            return joint.topk_joint(queries, shard, k, index_dtype=torch.int32)

        sc = cls(local_corpora, n_total, start, group=group, prepare=prepare, local_topk=local_topk,
                 merge=lambda s, i, k: ops.merge_topk(s, i, k))
        sc.n_local = sc.shard.n
        return sc

    def topk(self, queries, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Global top-k over all shards: (scores f32 [Q,k'], global rows i64 [Q,k']), k' = min(k, n_total)."""
        if self._injected:
            return self._topk_generic(queries, k)
        from . import ops
        k_glob = min(k, self.n_total)
        shard = self.shard
        if self.world == 1:
            return ops.topk(queries, shard, k)
        if shard.source is None:
            return self._topk_generic(queries, k)
        # local stage: K1 + fused tensor-core top-K' + strip merge, then the exact re-score writes this rank's
        # k best as packed {score bits, global row} pairs straight into the all-gather's send buffer
        dev = shard.device
        q = ops._as_rows(queries, dev)
        n_queries = q.shape[0]
        peer = self._peer_for(n_queries * k_glob, dev)
        if peer is not None:
            # K5 stores this rank's list into slot `rank` of EVERY rank's gather buffer (NVLink peer stores)
            ptrs, gathered_flat = peer.slot()
            dst, offset = ptrs, self.rank * peer.cap
            gathered = gathered_flat[:, :n_queries * k_glob].view(self.world, n_queries, k_glob, 2)
        else:
            send = torch.empty((n_queries, k_glob, 2), dtype=torch.int32, device=dev)
            dst, offset = [send.data_ptr()], 0
        # (an empty shard goes through the same calls: every candidate is (-inf, -1))
        q, q_inv, _, cand = ops.topk_candidates(q, shard, k)
        ops.rescore_pairs(q, q_inv, shard, cand, k_glob, dst, dst_offset_pairs=offset)
        if peer is not None:
            peer.commit()                      # one signal-pad barrier: every peer's stores have landed
            if gathered.is_contiguous():
                s, i = ops.merge_pairs(gathered, k_glob)
            else:
                s, i = ops.merge_pairs(gathered.contiguous(), k_glob)
            return s, i.to(torch.int64)
        # NCCL: Q * k * 8 bytes per rank over NVLink
        gathered = torch.empty((self.world, n_queries, k_glob, 2), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(gathered.view(self.world * n_queries, k_glob, 2), send, group=self.group)
        s, i = ops.merge_pairs(gathered, k_glob)
        return s, i.to(torch.int64)

    def _peer_for(self, n_pairs: int, dev: torch.device) -> Optional[_PeerExchange]:
        """Symmetric gather buffers big enough for n_pairs per rank, or None (-> NCCL).  Collective: every rank
        calls it with the same n_pairs and all of them agree on the outcome."""
        if self._exchange_req == "nccl" or self._peer_failed:
            return None
        if self._peer is not None and self._peer.cap >= n_pairs:
            return self._peer
        ok = 1
        peer = None
        try:
            peer = _PeerExchange(self.group, self.world, n_pairs, dev)
        except Exception as e:  # noqa: BLE001
            ok = 0
            self._peer_error = repr(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            self._peer_failed = True
            self._peer = None
            if self._exchange_req == "peer":
                raise RuntimeError(f"exchange='peer' requested but symmetric memory is unavailable: {getattr(self, '_peer_error', 'a peer failed')}")
            return None
        self._peer = peer
        self.exchange = "peer"
        return peer

    def _topk_generic(self, queries, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Same plumbing with separate score / row tensors and injectable stages (CPU tests; corpora without source)."""
        k_glob = min(k, self.n_total)
        s_loc, i_loc = self._local_topk(queries, self.shard, k)          # [Q, min(k, n_local)], global rows
        n_queries = s_loc.shape[0]
        # pad to a common width k_glob so that every rank gathers equal-sized blocks
        if s_loc.shape[1] < k_glob:
            pad = k_glob - s_loc.shape[1]
            s_loc = torch.cat([s_loc, s_loc.new_full((n_queries, pad), float("-inf"))], dim=1)
            i_loc = torch.cat([i_loc.to(torch.int32), i_loc.new_full((n_queries, pad), -1).to(torch.int32)], dim=1)
        s_loc = s_loc[:, :k_glob].contiguous().float()
        i_loc = i_loc[:, :k_glob].contiguous().to(torch.int32)
        if self.world == 1:
            return s_loc, i_loc.to(torch.int64)
        # one collective: (score bits, row) packed as int32 pairs -> Q * k * 8 bytes per rank
        packed = torch.stack([s_loc.view(torch.int32), i_loc], dim=-1).contiguous()
        gathered = torch.empty((self.world * n_queries,) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(gathered, packed, group=self.group)      # rank-major concatenation along dim 0
        gathered = gathered.view((self.world, n_queries) + tuple(packed.shape[1:]))
        s_all = gathered[..., 0].contiguous().view(torch.float32)
        i_all = gathered[..., 1].contiguous()
        s, i = self._merge(s_all, i_all, k_glob)
        return s, i.to(torch.int64)
