"""Developer tool (torchrun): where does the end-to-end step time go at N ranks?  Times, per variant, K batches of C3 shape.
python -m torch.distributed.run --nproc-per-node N tools/e2e_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-misinformation-detection_b200"))
import torch
import torch.distributed as dist
import mmd_retrieval as m
from mmd_retrieval.sharded import ShardedCorpus, shard_bounds

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
Q, N, D, k = 16384, 1_000_000, 768, 10
lo, hi = shard_bounds(N, world, rank)
g = torch.Generator(device=dev).manual_seed(17 + rank)
corpus = torch.randn(hi - lo, D, device=dev, generator=g)
queries = torch.randn(Q, D, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
q_host = queries.cpu().pin_memory()
K = 20


def wall(fn):
    fn(); fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e3 / K


# ---- cold sequence: how many batches until the end-to-end stream reaches its steady state?  (bench.py's e2e region follows
# a short warm-up; one-time costs -- pinned / device block allocation of the pipeline, NCCL buffers -- must not fall into it)
sc0 = ShardedCorpus(corpus, N, lo)
sc0.topk(queries, k)
for rnd in range(8):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    stamps = []
    for _ in sc0.topk_stream((q_host for _ in range(10)), k, to_host=True):
        stamps.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    if rank == 0:
        print(f"[probe world={world}] cold round {rnd}: {(time.perf_counter() - t0) * 100:.3f} ms/batch; yields at " + " ".join(f"{x:.1f}" for x in stamps), flush=True)
del sc0

out = {}
for sub in (1,):
    sc = ShardedCorpus(corpus, N, lo, sub_batches=sub)
    sc.topk(queries, k)
    out[f"sub={sub} device stream (topk_stream, device queries)"] = wall(lambda: [0 for _ in sc.topk_stream((queries for _ in range(K)), k)])
    out[f"sub={sub} device loop of topk() calls"] = wall(lambda: [sc.topk(queries, k) for _ in range(K)])
    out[f"sub={sub} e2e stream to_host"] = wall(lambda: [0 for _ in sc.topk_stream((q_host for _ in range(K)), k, to_host=True)])
    out[f"sub={sub} e2e stream host in, device out"] = wall(lambda: [0 for _ in sc.topk_stream((q_host for _ in range(K)), k)])
    out[f"sub={sub} stream device in, host out"] = wall(lambda: [0 for _ in sc.topk_stream((queries for _ in range(K)), k, to_host=True)])
    out[f"sub={sub} e2e loop of topk(host) + sync copy"] = wall(lambda: [[t.cpu() for t in sc.topk(q_host, k)] for _ in range(K)])
    h2d = torch.cuda.Stream()

    def up_only():
        for _ in range(K):
            with torch.cuda.stream(h2d):
                sc.upload_queries(q_host)
    out[f"sub={sub} upload_queries only (side stream)"] = wall(up_only)
    out[f"sub={sub} upload_queries only (current stream)"] = wall(lambda: [sc.upload_queries(q_host) for _ in range(K)])
    out[f"sub={sub} full-batch H2D per rank"] = wall(lambda: [q_host.to(dev, non_blocking=True) for _ in range(K)])
if rank == 0:
    for kk, v in out.items():
        print(f"[probe world={world}] {kk:56s} {v:9.3f} ms/batch")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
