"""Row-sharded corpus across the GPUs of one box.

One process per GPU (torchrun).  The corpus rows are split contiguously over the ranks; queries are
replicated; every rank runs the fused top-K on its shard and emits (score f32, GLOBAL row i32) lists; one
exchange of Q*K*8 bytes per rank and a `world`-way merge on the device (K4) give every rank the global top-K.
Scoring never crosses GPUs; the exchange is the path's only communication step.  Two implementations:

  exchange="peer"  the re-score kernel (K5) stores each rank's list directly into EVERY rank's gather buffer
                   (peer-mapped symmetric memory, the stores travel over NVLink 5 / NVSwitch while the kernel
                   is still scoring other queries); one signal-pad barrier follows.  No collective launch.
  exchange="nccl"  K5 fills a local send buffer, one `all_gather_into_tensor` moves it.

"auto" (default) uses "peer" when symmetric memory can be set up on every rank, else "nccl".

Order of the stages (`rescore="global"`, the default): the tensor-core pass over-fetches K' candidates per shard; the
K' raw lists are exchanged and merged FIRST, and only the global K' candidates are re-scored exactly -- each rank
re-scores the candidates that live in its own shard (K'/world per query on average instead of K') and a second, smaller
exchange + merge gives the final list.  `rescore="local"` re-scores every shard's K' candidates before a single exchange.
`ShardedCorpus.capture()` records the whole step (all kernels and the exchange) in a CUDA graph for replay.
The reference has no counterpart (single process, single device).

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

    Step i uses buffer i % 2 of every rank; every exchange inside a step is "store to all peers, then ONE signal-pad
    barrier".  Reuse without a barrier in front of the stores is safe: a rank passes the first barrier of step i only
    after every peer has finished its own stores of step i, which in stream order come after that peer's last read of
    step i-1 -- and the buffer written in step i+1 was last read in step i-1."""

    def __init__(self, group, world: int, cap_pairs: int, q_cap: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        self.world, self.cap, self.q_cap = world, cap_pairs, q_cap
        self.buf = symm_mem.empty((2, world, cap_pairs, 2), dtype=torch.int32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, grp)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        # per-query pruning thresholds shared by all ranks (one array per step parity), see mmd_topk_scores_shared
        self.thr = symm_mem.empty((2, q_cap), dtype=torch.int32, device=device)
        self.thr_hdl = symm_mem.rendezvous(self.thr, grp)
        self.thr_ptrs = [int(p) for p in self.thr_hdl.buffer_ptrs]
        self.thr.zero_()
        self.step = 0
        self.hdl.barrier(channel=0)            # every rank's threshold arrays are zero before anyone publishes

    def thr_slot(self, parity: int):
        """(this rank's threshold array of the given parity, every rank's) as device pointers."""
        off = parity * self.q_cap * 4
        return self.thr.data_ptr() + off, [p + off for p in self.thr_ptrs]

    def slot(self, parity: Optional[int] = None):
        """(peer base pointers of this step's buffer, local view [world, cap, 2] of it)."""
        b = self.step % 2 if parity is None else parity
        off_bytes = b * self.world * self.cap * 2 * 4
        return [p + off_bytes for p in self.ptrs], self.buf[b]

    def barrier(self):
        self.hdl.barrier(channel=0)

    def next_step(self):
        self.step += 1


class ShardedCorpus:
    """This rank's shard of a row-sharded corpus plus the collective that merges local top-K lists."""

    def __init__(self, local_rows, n_total: int, start: int, group=None, dtype: str = "bf16", metric: str = "cos",
                 eps: float = 1e-12, keep_source: bool = True, exchange: str = "auto", rescore: str = "global", share_thresholds: bool = True,
                 local_topk: Optional[Callable] = None, merge: Optional[Callable] = None, prepare: Optional[Callable] = None,
                 _shard=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = int(n_total)
        self.start = int(start)
        self._max_local = -(-self.n_total // self.world)           # rows of the largest shard (balanced contiguous split)
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
        if rescore not in ("global", "local"):
            raise ValueError("rescore must be 'global' or 'local'")
        self.rescore = rescore
        self.share_thresholds = share_thresholds
        self.phases = 1                   # single-GPU searches: launches per sweep of the shard (ops.topk_prepared_phased)
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
    def from_prepared(cls, shard, n_total: int, group=None, exchange: str = "auto", rescore: str = "global") -> "ShardedCorpus":
        """Wrap this rank's already prepared shard (e.g. corpus_io.prepare_streamed(..., idx_offset=start))."""
        return cls(None, n_total, shard.idx_offset, group=group, exchange=exchange, rescore=rescore, _shard=shard)

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
            return ops.topk(queries, shard, k, phases=self.phases)
        if shard.source is None:
            return self._topk_generic(queries, k)
        dev = shard.device
        q = self.upload_queries(queries)
        n_queries = q.shape[0]
        # every rank must exchange lists of one common width: the over-fetch of the LARGEST shard
        kp_glob = ops.overfetch_for(min(k, self._max_local), self._max_local)
        two_phase = self.rescore == "global"
        cap = n_queries * (kp_glob + k_glob) if two_phase else n_queries * k_glob
        peer = self._peer_for(cap, n_queries, dev)
        return self._search(q, k, k_glob, kp_glob, peer, two_phase, peer.step % 2 if peer is not None else 0, advance=True)

    def upload_queries(self, queries, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host-resident query batch (the SAME on every rank) -> device.  With more than one rank each rank uploads only its
        1/world slice over PCIe and the slices are all-gathered over NVLink, instead of every rank pulling the whole batch
        through the host's memory system: 8 x 50 MB per step at C3 otherwise.  Device tensors pass through."""
        from . import ops
        dev = self.shard.device
        if (isinstance(queries, torch.Tensor) and queries.is_cuda) or self.world == 1 or self._injected:
            q = ops._as_rows(queries, dev)
            if out is not None:
                out.copy_(q, non_blocking=True)
                return out
            return q
        q_host = ops._as_rows(queries)
        n_queries, dim = q_host.shape
        per = -(-n_queries // self.world)
        lo = min(n_queries, self.rank * per)
        hi = min(n_queries, lo + per)
        mine = torch.zeros((per, dim), dtype=q_host.dtype, device=dev) if hi - lo < per else \
            torch.empty((per, dim), dtype=q_host.dtype, device=dev)
        if hi > lo:
            mine[:hi - lo].copy_(q_host[lo:hi], non_blocking=True)
        full = torch.empty((self.world * per, dim), dtype=q_host.dtype, device=dev)
        dist.all_gather_into_tensor(full, mine, group=self.group)
        q = full[:n_queries]
        if out is not None:
            out.copy_(q, non_blocking=True)
            return out
        return q

    def _search(self, q, k, k_glob, kp_glob, peer, two_phase, parity, advance):
        """One search step on the current stream (eager or under CUDA-graph capture)."""
        from . import ops
        shard, dev, world, rank = self.shard, self.shard.device, self.world, self.rank
        n_queries = q.shape[0]
        # (an empty shard goes through the same calls: every candidate is (-inf, -1))
        # peer memory + global stage order: the pruning thresholds are shared across the GPUs as well
        # (a shard's K'-th best bounds the global K'-th best only if the shard lists are as long as the global candidate
        #  list: not the case for corpora so small that a shard holds fewer rows than that list)
        kc = min(world * kp_glob, ops.overfetch_for(k_glob, self.n_total))
        share_thr = peer is not None and two_phase and self.share_thresholds and kp_glob >= kc
        if peer is not None:
            ptrs, local = peer.slot(parity)                        # local: [world, cap, 2]
        # ... and when this shard can fill the common list width, the strip merge stores the raw candidate list straight
        # into every rank's gather buffer (no scatter kernel)
        fused_scatter = share_thr and shard.n >= kp_glob
        qd, q_inv, raw_s, cand = ops.topk_candidates(q, shard, k, overfetch=kp_glob,
                                                     shared_thr=peer.thr_slot(parity) if share_thr else None,
                                                     pair_dst=(ptrs, rank * peer.cap) if fused_scatter else None)

        def exchange(fill, width, region_off):
            """fill(dst_ptrs, pair_offset) stores this rank's [Q, width] list; returns the gathered [world] lists as
            (tensor, layout) ready for merge_pairs."""
            if peer is not None:
                fill(ptrs, rank * peer.cap + region_off)
                peer.barrier()                                     # every peer's stores have landed
                region = local[:, region_off:, :]                  # parts are peer.cap pairs apart
                if region_off == 0:
                    return ("strided", local)
                return ("strided_off", region)
            send = torch.empty((n_queries, width, 2), dtype=torch.int32, device=dev)
            fill([send.data_ptr()], 0)
            gathered = torch.empty((world, n_queries, width, 2), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(gathered.view(world * n_queries, width, 2), send, group=self.group)
            return ("dense", gathered)

        def merged(kind_t, width, k_out):
            kind, t = kind_t
            # every list in flight was produced sorted by this library (strip merge / re-score): no per-list sort
            if kind == "dense":
                return ops.merge_pairs(t, k_out, parts_sorted=True)
            if kind == "strided":
                return ops.merge_pairs(t, k_out, n_queries=n_queries, k_in=width, parts_sorted=True)
            return ops.merge_pairs_at(t, peer.cap, k_out, n_queries, width, parts_sorted=True)

        if two_phase:
            kp = cand.shape[1]
            if kp < kp_glob:                                       # a small shard: pad its raw list to the common width
                raw_s = torch.cat([raw_s, raw_s.new_full((n_queries, kp_glob - kp), float("-inf"))], dim=1).contiguous()
                cand = torch.cat([cand, cand.new_full((n_queries, kp_glob - kp), -1)], dim=1).contiguous()
            g1 = exchange((lambda d, off: None) if fused_scatter else (lambda d, off: ops.scatter_pairs(raw_s, cand, d, off)),
                          kp_glob, 0)
            if share_thr:
                # between the step's two barriers: every rank is past this step's contraction, nobody can have begun the
                # next one -- the other parity's thresholds (used by the next step) are cleared here
                peer.thr[1 - parity, :n_queries].zero_()
            _, cand_glob = merged(g1, kp_glob, kc)                 # the global K' candidates, identical on every rank
            g2 = exchange(lambda d, off: ops.rescore_pairs(qd, q_inv, shard, cand_glob, k_glob, d, dst_offset_pairs=off),
                          k_glob, n_queries * kp_glob)
            s, i = merged(g2, k_glob, k_glob)
        else:
            g = exchange(lambda d, off: ops.rescore_pairs(qd, q_inv, shard, cand, k_glob, d, dst_offset_pairs=off), k_glob, 0)
            s, i = merged(g, k_glob, k_glob)
        if peer is not None and advance:
            peer.next_step()
        return s, i.to(torch.int64)

    def capture(self, queries, k: int) -> "GraphedSearch":
        """Record the whole search step for `queries`' shape in CUDA graphs (one per exchange-buffer parity) and
        return a callable that replays them: `s, i = graphed(new_queries)`; results live in static output buffers
        until the next replay of the same parity."""
        return GraphedSearch(self, queries, k)

    def _peer_for(self, n_pairs: int, n_queries: int, dev: torch.device) -> Optional[_PeerExchange]:
        """Symmetric gather buffers big enough for n_pairs per rank, or None (-> NCCL).  Collective: every rank
        calls it with the same n_pairs and all of them agree on the outcome."""
        if self._exchange_req == "nccl" or self._peer_failed:
            return None
        if self._peer is not None and self._peer.cap >= n_pairs and self._peer.q_cap >= n_queries:
            return self._peer
        ok = 1
        peer = None
        try:
            peer = _PeerExchange(self.group, self.world, n_pairs, n_queries, dev)
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


class GraphedSearch:
    """A ShardedCorpus search step frozen into CUDA graphs (all kernels + the exchange), replayed per query batch."""

    def __init__(self, sc: ShardedCorpus, queries, k: int):
        from . import ops
        if sc._injected or sc.shard.source is None:
            raise RuntimeError("capture() needs the CUDA path with source embeddings kept (exact re-score)")
        self.sc, self.k = sc, k
        dev = sc.shard.device
        q = ops._as_rows(queries, dev)
        self.q_static = q.clone()
        n_queries = q.shape[0]
        k_glob = min(k, sc.n_total)
        kp_glob = ops.overfetch_for(min(k, sc._max_local), sc._max_local)
        two_phase = sc.rescore == "global" and sc.world > 1
        self.peer = None
        if sc.world > 1:
            cap = n_queries * (kp_glob + k_glob) if two_phase else n_queries * k_glob
            sc.topk(self.q_static, k)                              # warm-up: lazy initialisation (attributes, symmetric memory)
            if sc.exchange == "peer":
                self.peer = _PeerExchange(sc.group, sc.world, cap, n_queries, dev)      # this object's own double buffer
        else:
            ops.topk(self.q_static, sc.shard, k)
        torch.cuda.synchronize(dev)
        self.graphs, self.outs = [], []
        self.launches_per_step = 0
        n_graphs = 2 if self.peer is not None else 1
        stream = torch.cuda.Stream(device=dev)
        for parity in range(n_graphs):
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g, stream=stream):
                if sc.world > 1:
                    out = sc._search(self.q_static, k, k_glob, kp_glob, self.peer, two_phase, parity, advance=False)
                else:
                    out = ops.topk(self.q_static, sc.shard, k)
            self.launches_per_step = ops.launch_count() - n0      # this library's kernels inside one replay
            self.graphs.append(g)
            self.outs.append(out)
        self.calls = 0

    def __call__(self, queries=None):
        if queries is not None:
            self.sc.upload_queries(queries, out=self.q_static)
        j = self.calls % len(self.graphs)
        self.calls += 1
        self.graphs[j].replay()
        return self.outs[j]
