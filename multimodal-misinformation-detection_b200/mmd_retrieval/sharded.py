"""Row-sharded corpus across the GPUs of one box.

One process per GPU (torchrun).  The corpus rows are split contiguously over the ranks; queries are
replicated; every rank runs the fused top-K on its shard and emits (score f32, GLOBAL row i32) lists; one
all-gather of Q*K*8 bytes per rank over NCCL (NVLink 5 / NVSwitch) and a `world`-way merge on the device
(K4, mmd_topk_merge) give every rank the global top-K.  Scoring never crosses GPUs; the all-gather is the
path's only exchange step.  The reference has no counterpart (single process, single device).

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


class ShardedCorpus:
    """This rank's shard of a row-sharded corpus plus the collective that merges local top-K lists."""

    def __init__(self, local_rows, n_total: int, start: int, group=None, dtype: str = "bf16", metric: str = "cos",
                 eps: float = 1e-12, keep_source: bool = True,
                 local_topk: Optional[Callable] = None, merge: Optional[Callable] = None, prepare: Optional[Callable] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = int(n_total)
        self.start = int(start)
        self._local_topk = local_topk or _cuda_local_topk
        self._merge = merge or _cuda_merge
        if prepare is None:
            from . import ops
            self.shard = ops.prepare_corpus(local_rows, dtype=dtype, metric=metric, eps=eps, keep_source=keep_source,
                                            idx_offset=self.start)
            self.n_local = self.shard.n
        else:
            self.shard = prepare(local_rows, self.start)
            self.n_local = int(local_rows.shape[0])

    @classmethod
    def from_full(cls, corpus: torch.Tensor, group=None, **kw) -> "ShardedCorpus":
        """Every rank passes the same full corpus (or a view of it); each keeps only its own rows."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        lo, hi = shard_bounds(corpus.shape[0], world, rank)
        return cls(corpus[lo:hi], corpus.shape[0], lo, group=group, **kw)

    def topk(self, queries, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Global top-k over all shards: (scores f32 [Q,k'], global rows i64 [Q,k']), k' = min(k, n_total)."""
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
