"""Row-sharded corpus across the GPUs of one box.

One process per GPU (torchrun).  The corpus rows are split contiguously over the ranks; queries are replicated; every
rank runs the fused tensor-core top-K' on its shard; the ranks' candidate lists are exchanged, merged, re-scored exactly
and merged again.  Scoring never crosses GPUs.  The reference has no counterpart (single process, single device): the step
stands where sentence_transformers.util.semantic_search folds its corpus chunks together with a per-query heap (call
sites src/evidence/text2text_retrieval.py:56-64, src/evidence/experiment_text.py:25-33).

exchange="peer" (default when symmetric memory comes up on every rank): ONE search step is three kernels-with-flags per
rank and NO collective or barrier launch (csrc/exchange.cu, csrc/peer_sync.cuh):

  stage C   K1 on the queries -> fused tcgen05 score + top-K' over the shard (pruning thresholds shared by all GPUs through
            system-scope atomics) -> strip merge, which stores this rank's raw candidate list into EVERY rank's gather buffer
            over NVLink and raises flag set 1;
  stage X   wait for flag set 1 -> merge the `world` candidate lists (same result on every rank) -> exact fp32 re-score of
            the global candidates that live in THIS shard (K'/world per query on average) -> store each at its position in
            every rank's re-score buffer -> raise flag set 2;
  stage F   wait for flag set 2 -> sort the K' exact keys per query -> final top-k (int64 rows, no conversion kernel).

The step is PIPELINED: stage C runs on one stream, stages X and F on another, and a query batch is cut into sub-batches,
so the exchange/re-score tail of sub-batch i hides under the contraction of sub-batch i+1 (the tail -- ~0.3 ms of small
kernels at 8 GPUs -- was what bounded the 8-GPU efficiency of round 1).  All cross-GPU buffers form a ring of three slots;
why three, and why no barrier is needed:

  * stage C of step s starts only after THIS rank's stage F of step s-2 has finished.  F(s-2) waited for flag set 2 of
    step s-2 from every rank; a rank raises it at the end of its X(s-2), which runs (in-order stream) after its F(s-3).
    So when any rank begins to store step s's candidates into slot s % 3, every rank has finished reading slot (s-3) % 3.
  * the shared threshold array of a slot is reset at the start of stage C (a memset in front of the contraction).  Peers
    publish into it only during their own stage C of the same step (same queries => valid bounds); bounds that land before
    the reset are lost, which only weakens the pruning; bounds of step s-3 cannot arrive late because this rank passed
    flag set 1 of step s-3 long ago.  Every step resets exactly the entries it is about to read, so a smaller batch
    after a larger one never sees the larger one's bounds.

`topk()` waits for the last sub-batch before returning; `topk_stream()` also overlaps successive batches, including the
host->device upload of the next batch and the device->host read of the previous one.

exchange="nccl" (or rescore="local"): the plain variant -- per-rank lists, one `all_gather_into_tensor`, K4 merge.
`local_topk` / `merge` are injectable so the partition / offset / gather plumbing can be exercised with gloo on CPU
(tests/test_sharded_gloo.py injects the CPU oracle there -- the product path uses the CUDA ops and nothing else).
"""
from __future__ import annotations

import ctypes as C
from collections import deque
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

RING = 3          # slots of cross-GPU buffers (see module docstring)


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split: the first n_rows % world ranks hold one extra row."""
    base, extra = divmod(n_rows, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _cuda_local_topk(queries, shard, k, world: int = 1):
    """This rank's exact top-k of its shard.  The ranks' shards already act as independent sub-searches (ops.auto_splits), so
    a shard is split further only as far as the world size leaves the effective over-fetch short."""
    from . import joint, ops
    if isinstance(shard, joint.JointCorpus):
        return joint.topk_joint(queries, shard, k, index_dtype=torch.int32)
    k_eff = max(1, min(k, shard.n))
    want = ops.auto_splits(shard.op, k_eff, ops.overfetch_for(k_eff, shard.n), shard.n * world)
    return ops.topk(queries, shard, k, index_dtype=torch.int32, splits=-(-want // max(1, world)))


def _cuda_merge(scores, idx, k):
    from . import ops
    return ops.merge_topk(scores, idx, k)


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class _PeerRing:
    """Symmetric (peer-mapped) memory of the sharded step: RING slots of {gather buffer [world][q_cap x kp] pairs, re-score
    buffer [q_cap x kc] pairs, threshold array [q_cap]} plus two flag arrays [world], carved out of ONE allocation; local
    sync words and scratch (prepared queries, raw lists, workspace) per slot."""

    def __init__(self, group, world: int, rank: int, q_cap: int, kp: int, kc: int, row_bytes: int, n_seg: int,
                 device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        self.world, self.rank, self.q_cap, self.kp, self.kc = world, rank, q_cap, kp, kc
        self.row_bytes, self.n_seg, self.ws_bytes = row_bytes, n_seg, 0
        self.device = device
        self.gather_bytes = _align(world * q_cap * kp * 8)
        self.resc_bytes = _align(q_cap * kc * 8)
        self.thr_bytes = _align(q_cap * 4)
        self.off_gather = 0
        self.off_resc = self.off_gather + RING * self.gather_bytes
        self.off_thr = self.off_resc + RING * self.resc_bytes
        self.off_flags = self.off_thr + RING * self.thr_bytes
        total = self.off_flags + 2 * 256
        self.buf = symm_mem.empty((total,), dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, grp)
        self.base = [int(p) for p in self.hdl.buffer_ptrs]
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=group)                                  # every rank's flags are zero before anyone raises one
        # local: sync words {launches, blocks done} of stages C, X, F (64 B apart) and per-slot scratch
        self.sync = torch.zeros((48,), dtype=torch.int32, device=device)
        self.q_rows = [torch.empty((q_cap, row_bytes), dtype=torch.uint8, device=device) for _ in range(RING)]
        self.q_inv = [[torch.empty((q_cap,), dtype=torch.float32, device=device) for _ in range(n_seg)] for _ in range(RING)]
        self.raw_s = [torch.empty((q_cap, kp), dtype=torch.float32, device=device) for _ in range(RING)]
        self.raw_i = [torch.empty((q_cap, kp), dtype=torch.int32, device=device) for _ in range(RING)]
        self.ws = [None] * RING                                    # fused-kernel workspace per slot, grown on demand (local)
        self.ev_c = [torch.cuda.Event() for _ in range(RING)]
        self.ev_f = [torch.cuda.Event() for _ in range(RING)]
        self.step = 0
        torch.cuda.synchronize(device)

    def fits(self, q_sub: int, kp: int, kc: int, row_bytes: int, n_seg: int) -> bool:
        return q_sub <= self.q_cap and kp == self.kp and kc == self.kc and row_bytes == self.row_bytes and n_seg == self.n_seg

    def ensure_workspace(self, nbytes: int) -> None:
        """Local scratch, no collective.  Growing it waits for the device first: nothing may still be using the old one."""
        if nbytes > self.ws_bytes:
            torch.cuda.synchronize(self.device)
            self.ws = [torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=self.device) for _ in range(RING)]
            self.ws_bytes = nbytes

    # device pointers (ints)
    def gather_all(self, slot):  # every rank's gather buffer of this slot
        return [b + self.off_gather + slot * self.gather_bytes for b in self.base]

    def resc_all(self, slot):
        return [b + self.off_resc + slot * self.resc_bytes for b in self.base]

    def thr_all(self, slot):
        return [b + self.off_thr + slot * self.thr_bytes for b in self.base]

    def flags_local(self, which):
        return self.base[self.rank] + self.off_flags + which * 256

    def flags_arrive(self, which):  # this rank's word in every rank's flag array
        return [b + self.off_flags + which * 256 + self.rank * 4 for b in self.base]

    def sync_ptr(self, stage):
        return self.sync.data_ptr() + stage * 64


class ShardedCorpus:
    """This rank's shard of a row-sharded corpus plus the exchange that merges the ranks' lists."""

    def __init__(self, local_rows, n_total: int, start: int, group=None, dtype: str = "bf16", metric: str = "cos",
                 eps: float = 1e-12, keep_source: bool = True, exchange: str = "auto", rescore: Optional[str] = None, share_thresholds: bool = True,
                 sub_batches: int = 0, local_topk: Optional[Callable] = None, merge: Optional[Callable] = None,
                 prepare: Optional[Callable] = None, _shard=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = int(n_total)
        self.start = int(start)
        self._max_local = -(-self.n_total // self.world)           # rows of the largest shard (balanced contiguous split)
        self._min_local = self.n_total // self.world
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
        if rescore not in (None, "global", "local"):
            raise ValueError("rescore must be 'global', 'local' or None (auto)")
        self._rescore_req = rescore
        self.rescore = rescore or "global"        # auto: settled below, once the shard's operand type is known
        self.share_thresholds = share_thresholds
        self.sub_batches = int(sub_batches)   # 0 = auto
        self.phases = 1                   # single-GPU searches: launches per sweep of the shard (ops.topk_prepared_phased)
        self._exchange_req = exchange
        self.exchange = "nccl"            # what is actually in use; "peer" once symmetric memory is up on every rank
        self._ring: Optional[_PeerRing] = None            # the ring of the call in progress
        self._rings = {}                                  # (kp, kc, row_bytes, n_seg) -> ring (one per list geometry, grown on demand)
        self._peer_failed = False
        self._streams = None
        self._injected = local_topk is not None or merge is not None or prepare is not None
        self._local_topk = local_topk or (lambda q, shard, k: _cuda_local_topk(q, shard, k, self.world))
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
        if self._rescore_req is None and getattr(self.shard, "op", None) == "fp8" and not self._injected:
            # e4m3 candidates are lossy: merging the ranks' candidate lists by their fp8 scores down to one K'-list before the
            # re-score would throw away what the sharding buys (every shard's K' candidates are a `world`-fold over-fetch).
            # Re-score per shard, exchange exact lists.  (bf16 / fp16 candidates: re-score after the global merge, K'/world rows
            # per rank -- the sub-millisecond C3 step at 8 GPUs is where that matters.)
            self.rescore = "local"

    # ------------------------------------------------------------------------------------------ constructors
    @classmethod
    def from_full(cls, corpus: torch.Tensor, group=None, **kw) -> "ShardedCorpus":
        """Every rank passes the same full corpus (or a view of it); each keeps only its own rows."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        lo, hi = shard_bounds(corpus.shape[0], world, rank)
        return cls(corpus[lo:hi], corpus.shape[0], lo, group=group, **kw)

    @classmethod
    def from_prepared(cls, shard, n_total: int, group=None, **kw) -> "ShardedCorpus":
        """Wrap this rank's already prepared shard (e.g. corpus_io.prepare_streamed(..., idx_offset=start)), a PreparedCorpus
        or a JointCorpus."""
        return cls(None, n_total, shard.idx_offset, group=group, _shard=shard, **kw)

    @classmethod
    def from_joint(cls, local_corpora, n_total: int, start: int, weights=None, group=None, dtype: str = "bf16",
                   metric: str = "cos", eps: float = 1e-12, **kw) -> "ShardedCorpus":
        """Row-sharded JOINT (multi-modality) corpus: local_corpora = this rank's rows of every modality; queries are
        passed to topk() as a list with one matrix per modality.  Same exchange as the single-modality corpus."""
        from . import joint
        jc = joint.prepare_joint(local_corpora, weights, dtype=dtype, metric=metric, eps=eps, idx_offset=start)
        return cls(None, n_total, start, group=group, _shard=jc, **kw)

    # ------------------------------------------------------------------------------------------ helpers
    @property
    def _is_joint(self) -> bool:
        from . import joint
        return isinstance(self.shard, joint.JointCorpus)

    def _has_source(self) -> bool:
        return self.shard.source is not None

    def _query_mats(self, queries, device=None) -> List[torch.Tensor]:
        """queries (one matrix, or one per modality for a joint corpus) -> list of 2-D row tensors."""
        from . import ops
        if self._is_joint:
            if not isinstance(queries, (list, tuple)) or len(queries) != len(self.shard.dims):
                raise ValueError(f"a joint corpus takes a list of {len(self.shard.dims)} query matrices (one per modality)")
            mats = [ops._as_rows(q, device) for q in queries]
            dims = self.shard.dims
        else:
            mats = [ops._as_rows(queries, device)]
            dims = [self.shard.dim]
        n = mats[0].shape[0]
        for m, d in zip(mats, dims):
            if m.shape[1] != d:
                raise RuntimeError(f"query dim {m.shape[1]} does not match corpus dim {d}")
            if m.shape[0] != n:
                raise ValueError("every modality must have the same number of queries")
        return mats

    def _pipe_streams(self):
        if self._streams is None:
            dev = self.shard.device
            self._streams = tuple(torch.cuda.Stream(device=dev) for _ in range(5))     # contraction, tail, h2d, d2h, query K1
        return self._streams

    def _widths(self, k: int) -> Tuple[int, int, int]:
        """(k_glob, kp, kc): final list length, candidate-list width every rank exchanges, global candidate-list length."""
        from . import ops
        k_glob = min(k, self.n_total)
        kp = ops.overfetch_for(min(k, self._max_local), self._max_local)
        kc = min(self.world * kp, ops.overfetch_for(k_glob, self.n_total))
        return k_glob, kp, kc

    # ------------------------------------------------------------------------------------------ public search
    def topk(self, queries, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Global top-k over all shards: (scores f32 [Q,k'], global rows i64 [Q,k']), k' = min(k, n_total)."""
        if k <= 0:
            raise ValueError("k must be positive")
        if self._injected:
            return self._topk_generic(queries, k)
        if self.world == 1:
            return _cuda_local_topk_i64(queries, self.shard, k, self.phases)
        if not self._has_source() or self.rescore == "local" or self._exchange_req == "nccl":
            return self._topk_gather(queries, k)
        dev = self.shard.device
        mats = self.upload_queries(queries)
        if not self._ring_for(mats[0].shape[0], k):
            return self._topk_gather(mats if self._is_joint else mats[0], k)
        s, i, done = self._run_pipelined(mats, k, None)
        if done is not None:
            torch.cuda.current_stream(dev).wait_event(done)
        return s, i

    def topk_stream(self, batches: Iterable, k: int, to_host: bool = False, depth: int = 2):
        """Pipelined search over an iterable of query batches; yields (scores, rows) per batch, in order.

        The host->device upload of batch i+1 (own stream), the search of batch i and -- with to_host=True -- the
        device->host read of batch i-1 into pinned buffers (own stream) overlap; with several GPUs the exchange tail of a
        batch also hides under the contraction of the next one.  Up to `depth` batches are in flight behind the one being
        yielded.  to_host=False yields device tensors that are safe to use on the caller's current stream."""
        if self._injected:
            for b in batches:
                yield self._topk_generic(b, k)
            return
        dev = self.shard.device
        c_stream, t_stream, h2d, d2h, _ = self._pipe_streams()
        cur = torch.cuda.current_stream(dev)
        pending = deque()
        plain = self.world > 1 and (not self._has_source() or self.rescore == "local" or self._exchange_req == "nccl")

        def finish(item):
            s, i, done, ev_host, hs, hi, keep = item
            if to_host:
                ev_host.synchronize()
                return hs, hi
            if done is not None:
                cur.wait_event(done)
            s.record_stream(cur)
            i.record_stream(cur)
            return s, i

        for b in batches:
            ev_b = torch.cuda.Event()                              # the batch (if device-resident) is ready on the caller's stream
            ev_b.record(cur)
            h2d.wait_event(ev_b)
            with torch.cuda.stream(h2d):
                mats = self.upload_queries(b)
                ev_up = torch.cuda.Event()
                ev_up.record(h2d)
            if self.world == 1 or plain or not self._ring_for(mats[0].shape[0], k):
                with torch.cuda.stream(c_stream):
                    c_stream.wait_event(ev_up)
                    qq = mats if self._is_joint else mats[0]
                    if self.world == 1:
                        s, i = _cuda_local_topk_i64(qq, self.shard, k, self.phases)
                    else:
                        s, i = self._topk_gather(qq, k)
                    done = torch.cuda.Event()
                    done.record(c_stream)
                for m in mats:
                    m.record_stream(c_stream)
            else:
                s, i, done = self._run_pipelined(mats, k, ev_up, overlapped=True)
                for m in mats:
                    m.record_stream(c_stream)
                    m.record_stream(t_stream)
            ev_host = hs = hi = None
            if to_host:
                with torch.cuda.stream(d2h):
                    if done is not None:
                        d2h.wait_event(done)
                    hs = torch.empty(s.shape, dtype=s.dtype, pin_memory=True)
                    hi = torch.empty(i.shape, dtype=i.dtype, pin_memory=True)
                    hs.copy_(s, non_blocking=True)
                    hi.copy_(i, non_blocking=True)
                    ev_host = torch.cuda.Event()
                    ev_host.record(d2h)
                s.record_stream(d2h)
                i.record_stream(d2h)
            pending.append((s, i, done, ev_host, hs, hi, mats))
            while len(pending) > depth:
                yield finish(pending.popleft())
        while pending:
            yield finish(pending.popleft())

    def upload_queries(self, queries, out: Optional[torch.Tensor] = None):
        """Host-resident query batch (the SAME on every rank) -> list of device matrices (one per modality) on the current
        stream.  With more than one rank each rank uploads only its 1/world slice over PCIe and the slices are all-gathered
        over NVLink, instead of every rank pulling the whole batch through the host's memory system: 8 x 50 MB per step at
        C3 otherwise.  Device tensors pass through.  `out` (single-modality only): static destination buffer."""
        dev = self.shard.device
        if self._is_joint and out is None and isinstance(queries, (list, tuple)):
            return [self._upload_one(q, None) for q in self._query_mats(queries)]
        mats = self._query_mats(queries)
        return [self._upload_one(mats[0], out)]

    def _upload_one(self, q, out):
        from . import ops
        dev = self.shard.device
        if q.is_cuda or self.world == 1 or self._injected:
            q = ops._as_rows(q, dev)
            if out is not None:
                out.copy_(q, non_blocking=True)
                return out
            return q
        n_queries, dim = q.shape
        per = -(-n_queries // self.world)
        lo = min(n_queries, self.rank * per)
        hi = min(n_queries, lo + per)
        mine = torch.zeros((per, dim), dtype=q.dtype, device=dev) if hi - lo < per else \
            torch.empty((per, dim), dtype=q.dtype, device=dev)
        if hi > lo:
            mine[:hi - lo].copy_(q[lo:hi], non_blocking=True)
        full = torch.empty((self.world * per, dim), dtype=q.dtype, device=dev)
        dist.all_gather_into_tensor(full, mine, group=self.group)
        res = full[:n_queries]
        if out is not None:
            out.copy_(res, non_blocking=True)
            return out
        return res

    # ------------------------------------------------------------------------------------------ peer-memory pipelined step
    def _sub_sizes(self, n_queries: int) -> List[int]:
        """Sub-batches of one call (whole 256-query tiles).  Default 1: every contraction launch pays its own cold start and
        tail quantisation (measured at 2 GPUs, C3: two half-batch launches cost 0.18 ms more kernel time than one), which is
        what a split would hide of the exchange tail; successive batches overlap through topk_stream() instead."""
        want = self.sub_batches if self.sub_batches > 0 else 1
        tiles = -(-n_queries // 256)
        want = max(1, min(want, tiles))
        per = -(-tiles // want) * 256
        sizes = []
        left = n_queries
        while left > 0:
            sizes.append(min(per, left))
            left -= sizes[-1]
        return sizes

    def _ring_for(self, n_queries: int, k: int) -> bool:
        """Make sure the symmetric ring fits this call (COLLECTIVE: every rank calls it with the same arguments and all
        agree on the outcome).  False -> fall back to the NCCL gather."""
        if self._peer_failed:
            return False
        from . import _lib, ops
        shard = self.shard
        k_glob, kp, kc = self._widths(k)
        q_sub = max(self._sub_sizes(n_queries)) if n_queries else 1
        row_bytes = shard.rows.shape[1]
        n_seg = len(shard.dims) if self._is_joint else 1
        key = (kp, kc, row_bytes, n_seg)
        old = self._rings.get(key)
        if old is not None and old.fits(q_sub, kp, kc, row_bytes, n_seg):
            self._ring = old
            return True
        dev = shard.device
        ok, ring = 1, None
        try:
            if old is not None:
                torch.cuda.synchronize(dev)                       # nothing of the old ring is in flight
                q_sub = max(q_sub, old.q_cap)
            ring = _PeerRing(self.group, self.world, self.rank, q_sub, kp, kc, row_bytes, n_seg, dev)
        except Exception as e:  # noqa: BLE001
            ok = 0
            self._peer_error = repr(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            self._peer_failed = True
            self._ring = None
            if self._exchange_req == "peer":
                raise RuntimeError(f"exchange='peer' requested but symmetric memory is unavailable: {getattr(self, '_peer_error', 'a peer failed')}")
            return False
        self._rings[key] = ring
        self._ring = ring
        self.exchange = "peer"
        return True

    def _segment_tables(self, mats: List[torch.Tensor], row0: int, q_invs):
        """ctypes tables of the re-score for the query rows starting at row0."""
        from . import ops
        shard = self.shard
        cos = shard.metric == "cos"
        if self._is_joint:
            srcs, c_invs, dims, weights = shard.sources, shard.inv_norms, shard.dims, shard.weights
        else:
            srcs, c_invs, dims, weights = [shard.source], [shard.inv_norm], [shard.dim], [1.0]
        return ops.segment_tables(mats, row0, [t if cos else None for t in q_invs], srcs, [t if cos else None for t in c_invs],
                                  dims, weights)

    def _prepare_queries(self, sub: List[torch.Tensor], ring: _PeerRing, slot: int):
        """K1 on a sub-batch into the slot's scratch -> (prepared rows, [inv_norm per modality])."""
        from . import _lib, joint, ops
        shard = self.shard
        n = sub[0].shape[0]
        if self._is_joint:
            rows, invs = joint._cast_segments(sub, shard.op, _lib.SIDE_QUERY, shard.metric == "cos", shard.eps, shard.weights,
                                              shard.seg_bytes, sum(shard.seg_bytes), out=(ring.q_rows[slot], ring.q_inv[slot]))
            return rows, invs
        rows, inv = ops.normalize_cast(sub[0], shard.op, _lib.SIDE_QUERY, shard.metric == "cos", shard.eps,
                                       out=(ring.q_rows[slot], ring.q_inv[slot][0]))
        return rows, [inv]

    def _run_pipelined(self, mats: List[torch.Tensor], k: int, ev_in, overlapped: bool = False):
        """Enqueue one query batch (device matrices, ready on the current stream and -- if given -- at `ev_in`) as pipelined
        sub-steps.  Returns (scores, rows, event that fires when the results are complete).  Nothing here blocks the host."""
        from . import _lib, ops
        lib = _lib.load()
        shard, dev, world, rank, ring = self.shard, self.shard.device, self.world, self.rank, self._ring
        c_stream, t_stream, _, _, k_stream = self._pipe_streams()
        n_queries = mats[0].shape[0]
        k_glob, kp, kc = self._widths(k)
        out_s = torch.empty((n_queries, k_glob), dtype=torch.float32, device=dev)
        out_i = torch.empty((n_queries, k_glob), dtype=torch.int64, device=dev)
        if n_queries == 0:
            return out_s, out_i, None
        capturing = torch.cuda.is_current_stream_capturing()
        op = ops._OP_DTYPE[shard.op]
        # a shard's K'-th best bounds the global K'-th best only if every shard's list is as long as the global candidate list
        share_thr = self.share_thresholds and kp >= kc and self._min_local >= kc
        # Overlapped batches (topk_stream): the ranks drift apart by up to a batch, a rank's publications then reach peers that
        # are in another stage, and what remains of the sharing is its cost -- one system-scope atomic per rank and bound over
        # NVLink.  Measured at 8 GPUs on C3, same box, interleaved (profiles/r2_n8_policy_probe.log): overlapped batches 3.08 ms
        # with shared thresholds, 2.59 without; one batch at a time 2.69 with, 2.86 without.  At 2 GPUs the stream measured
        # 9.20 ms with sharing and 9.68 without (different boxes).  So: shared for topk(); for the stream shared up to 4 ranks and
        # private beyond -- unless the caller asked for sharing explicitly (share_thresholds="always").
        if overlapped and world > 4 and self.share_thresholds != "always":
            share_thr = False
        k_loc = max(1, min(kp, shard.n))
        part_stride = ring.q_cap * kp
        sizes = self._sub_sizes(n_queries)
        if shard.n > 0:
            ring.ensure_workspace(max(int(lib.mmd_topk_workspace_bytes(n, shard.n, shard.dim, op, k_loc)) for n in set(sizes)))
        else:
            ring.ensure_workspace(8)
        # the internal streams start behind the caller's stream as of NOW: the queries are ready there, and the freshly
        # allocated outputs (caller's pool) may reuse memory whose last use was enqueued there
        ev_now = torch.cuda.Event()
        ev_now.record(torch.cuda.current_stream(dev))
        k_stream.wait_event(ev_now)
        if ev_in is not None:
            k_stream.wait_event(ev_in)
        row0 = 0
        # under CUDA-graph capture the events must belong to the capture; replays of a graph are serialised as a whole, so
        # the cross-step waits are only needed (and only legal) between sub-steps of the same capture
        ev_k = [torch.cuda.Event() for _ in sizes] if capturing else None
        ev_f = [torch.cuda.Event() for _ in sizes] if capturing else None
        for j, n in enumerate(sizes):
            step = ring.step
            slot = step % RING
            sub = [m[row0:row0 + n] for m in mats]
            e_k = ev_k[j] if capturing else ring.ev_c[slot]
            e_f = ev_f[j] if capturing else ring.ev_f[slot]
            # ---- query side of stage C on its own stream (hides under the previous contraction): reset of the slot's shared
            # thresholds + K1.  The slot was last used by step - 3: its stage F must be over (then this rank has seen flag
            # set 1 of that step, i.e. no peer can still publish a bound of those queries; and the scratch is free).
            with torch.cuda.stream(k_stream):
                if capturing:
                    if j >= RING:
                        k_stream.wait_event(ev_f[j - RING])
                elif step >= RING:
                    k_stream.wait_event(ring.ev_f[slot])
                if share_thr:
                    ops.zero_u32(ring.thr_all(slot)[rank], n, dev)
                q_rows, q_invs = self._prepare_queries(sub, ring, slot)
                e_k.record(k_stream)
            # ---- stage C: the contraction; its strip merge (+ stores to the peers + flag set 1) goes to the tail stream
            with torch.cuda.stream(c_stream):
                c_stream.wait_event(e_k)
                if capturing:
                    if j >= 2:
                        c_stream.wait_event(ev_f[j - 2])
                elif step >= 2:
                    c_stream.wait_event(ring.ev_f[(step - 2) % RING])          # slot free on every rank (see module docstring)
                ops.sharded_candidates(q_rows, n, shard, k_loc, ring.raw_s[slot], ring.raw_i[slot], ring.ws[slot], ring.thr_all(slot),
                                       rank, share_thr, ring.gather_all(slot), rank * part_stride, kp, ring.flags_arrive(0),
                                       ring.sync_ptr(0), reset_thr=False, merge_stream=t_stream)
            # ---- stages X and F
            with torch.cuda.stream(t_stream):
                ops.exchange_rescore(ring.gather_all(slot)[rank], world, part_stride, n, kp, kc, self._segment_tables(mats, row0, q_invs),
                                     shard.n, shard.idx_offset, ring.resc_all(slot), rank, dev, ring.flags_local(0), world,
                                     ring.flags_arrive(1), ring.sync_ptr(1))
                ops.exchange_finish(ring.resc_all(slot)[rank], n, kc, k_glob, out_s.data_ptr() + row0 * k_glob * 4,
                                    out_i.data_ptr() + row0 * k_glob * 8, True, dev, ring.flags_local(1), world, ring.sync_ptr(2))
                e_f.record(t_stream)
            ring.step += 1
            row0 += n
        done = torch.cuda.Event()
        with torch.cuda.stream(t_stream):
            done.record(t_stream)
        return out_s, out_i, done

    # ------------------------------------------------------------------------------------------ plain gather variants
    def _topk_gather(self, queries, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """NCCL variant on the current stream: local lists (re-scored per shard, or raw candidates first and the global
        candidates re-scored after a first gather), `all_gather_into_tensor`, K4 merge."""
        from . import joint, ops
        shard, dev, world = self.shard, self.shard.device, self.world
        if not self._has_source() or self._is_joint or self.rescore == "local":
            return self._topk_generic(queries, k)
        k_glob, kp, kc = self._widths(k)
        q = self.upload_queries(queries)[0]
        n_queries = q.shape[0]
        qd, q_inv, raw_s, cand = ops.topk_candidates(q, shard, k, overfetch=kp)
        if cand.shape[1] < kp:                                   # a small shard: pad its raw list to the common width
            pad = kp - cand.shape[1]
            raw_s = torch.cat([raw_s, raw_s.new_full((n_queries, pad), float("-inf"))], dim=1).contiguous()
            cand = torch.cat([cand, cand.new_full((n_queries, pad), -1)], dim=1).contiguous()

        def gather(fill, width):
            send = torch.empty((n_queries, width, 2), dtype=torch.int32, device=dev)
            fill(send)
            gathered = torch.empty((world, n_queries, width, 2), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(gathered.view(world * n_queries, width, 2), send, group=self.group)
            return gathered

        g1 = gather(lambda send: ops.scatter_pairs(raw_s, cand, [send.data_ptr()], 0), kp)
        _, cand_glob = ops.merge_pairs(g1, kc, parts_sorted=True)
        g2 = gather(lambda send: ops.rescore_pairs(qd, q_inv, shard, cand_glob, k_glob, [send.data_ptr()], 0), k_glob)
        s, i = ops.merge_pairs(g2, k_glob, parts_sorted=True)
        return s, i.to(torch.int64)

    def capture(self, queries, k: int) -> "GraphedSearch":
        """Record the whole search step for `queries`' shape in a CUDA graph and return a callable that replays it:
        `s, i = graphed(new_queries)`; results live in static output buffers until the next replay."""
        return GraphedSearch(self, queries, k)

    def _topk_generic(self, queries, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Same plumbing with separate score / row tensors and injectable stages (CPU tests; corpora without source;
        re-score per shard)."""
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


def _cuda_local_topk_i64(queries, shard, k, phases=1):
    from . import joint, ops
    if isinstance(shard, joint.JointCorpus):
        return joint.topk_joint(queries, shard, k)
    return ops.topk(queries, shard, k, phases=phases)


class GraphedSearch:
    """A ShardedCorpus search step frozen into a CUDA graph (all kernels of all stages, both internal streams), replayed per
    query batch.  The flag sequence numbers of the exchange live in device memory, so a replay is a valid next step."""

    def __init__(self, sc: ShardedCorpus, queries, k: int):
        from . import ops
        if sc._injected or not sc._has_source():
            raise RuntimeError("capture() needs the CUDA path with source embeddings kept (exact re-score)")
        if sc._is_joint:
            raise RuntimeError("capture() supports single-modality corpora")
        self.sc, self.k = sc, k
        dev = sc.shard.device
        q = ops._as_rows(queries, dev)
        self.q_static = q.clone()
        self.peer = sc.world > 1
        sc.topk(self.q_static, k)                                  # warm-up: lazy initialisation (attributes, symmetric memory)
        if self.peer and sc.exchange != "peer":
            raise RuntimeError("capture() with several ranks needs exchange='peer' (NCCL collectives are not captured here)")
        torch.cuda.synchronize(dev)
        if self.peer:
            dist.barrier(group=sc.group)
        stream = torch.cuda.Stream(device=dev)
        self.graphs, self.outs = [], []
        g = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(g, stream=stream):
            out = sc.topk(self.q_static, k)
        self.launches_per_step = ops.launch_count() - n0          # this library's kernels inside one replay
        self.graphs.append(g)
        self.outs.append(out)
        self.calls = 0

    def __call__(self, queries=None):
        if queries is not None:
            self.sc.upload_queries(queries, out=self.q_static)
        self.calls += 1
        self.graphs[0].replay()
        return self.outs[0]
