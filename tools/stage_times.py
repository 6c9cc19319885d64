"""Developer tool (torchrun): per-stage device times of the row-sharded search step, for every exchange x stage-order
combination.  python -m torch.distributed.run --nproc-per-node N tools/stage_times.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-misinformation-detection_b200"))
import torch
import torch.distributed as dist
import mmd_retrieval as m
from mmd_retrieval import ops
from mmd_retrieval.sharded import ShardedCorpus, shard_bounds

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
Q, N, D, k = 16384, 1_000_000, 768, 10
lo, hi = shard_bounds(N, world, rank)
g = torch.Generator(device=dev).manual_seed(17 + rank)
corpus = torch.randn(hi - lo, D, device=dev, generator=g)
queries = torch.randn(Q, D, device=dev, generator=torch.Generator(device=dev).manual_seed(5))


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


out = {}
configs = {"peer/global (thresholds shared across GPUs, scatter fused into the strip merge)": dict(exchange="peer", rescore="global"),
           "peer/global, thresholds NOT shared": dict(exchange="peer", rescore="global", share_thresholds=False),
           "peer/local": dict(exchange="peer", rescore="local"),
           "nccl/global": dict(exchange="nccl", rescore="global"),
           "nccl/local": dict(exchange="nccl", rescore="local")}
scs = {name: ShardedCorpus(corpus, N, lo, **kw) for name, kw in configs.items()}
for rnd in range(3):                       # interleaved rounds, best of three: no configuration owns the cold / hot GPU
    for name, sc_i in scs.items():
        t = timed(lambda: sc_i.topk(queries, k))
        out["step " + name] = min(out.get("step " + name, 1e9), t)
del scs
sc = ShardedCorpus(corpus, N, lo, exchange="peer")
sc.topk(queries, k)
shard = sc.shard
out["candidates (K1 + fused + strip merge), local thresholds"] = timed(lambda: ops.topk_candidates(queries, shard, k, overfetch=18))
qd, q_inv, raw_s, cand = ops.topk_candidates(queries, shard, k, overfetch=18)
peer = sc._peer
ptrs, local = peer.slot(0)
out["scatter_pairs to all peers"] = timed(lambda: ops.scatter_pairs(raw_s, cand, ptrs, rank * peer.cap))
out["scatter_pairs local only"] = timed(lambda: ops.scatter_pairs(raw_s, cand, [ptrs[rank]], rank * peer.cap))
out["signal-pad barrier"] = timed(lambda: peer.barrier())
out["merge_pairs world x 18 -> 18"] = timed(lambda: ops.merge_pairs(local, 18, n_queries=Q, k_in=18))
_, cg = ops.merge_pairs(local, 18, n_queries=Q, k_in=18)
out["rescore_pairs global cand -> peers"] = timed(lambda: ops.rescore_pairs(qd, q_inv, shard, cg, k, ptrs, dst_offset_pairs=rank * peer.cap + Q * 18))
out["rescore_pairs local cand -> local buffer"] = timed(lambda: ops.rescore_pairs(qd, q_inv, shard, cand, k, [ptrs[rank]], dst_offset_pairs=rank * peer.cap + Q * 18))
out["rescore_pairs local cand -> peers"] = timed(lambda: ops.rescore_pairs(qd, q_inv, shard, cand, k, ptrs, dst_offset_pairs=rank * peer.cap + Q * 18))
send = torch.empty((Q, k, 2), dtype=torch.int32, device=dev)
gath = torch.empty((world * Q, k, 2), dtype=torch.int32, device=dev)
out["nccl all_gather Q*k pairs"] = timed(lambda: dist.all_gather_into_tensor(gath, send))
send2 = torch.empty((Q, 18, 2), dtype=torch.int32, device=dev)
gath2 = torch.empty((world * Q, 18, 2), dtype=torch.int32, device=dev)
out["nccl all_gather Q*18 pairs"] = timed(lambda: dist.all_gather_into_tensor(gath2, send2))
out["merge_pairs world x 10 -> 10"] = timed(lambda: ops.merge_pairs(gath.view(world, Q, k, 2), k))
if rank == 0:
    for kk, v in out.items():
        print(f"[stage world={world}] {kk:48s} {v*1e3:9.1f} us")
dist.barrier()
dist.destroy_process_group()
