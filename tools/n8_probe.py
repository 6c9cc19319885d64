"""Developer tool (torchrun): device-resident C3 step time at N ranks under a few host-side policies, same process, interleaved.
python -m torch.distributed.run --nproc-per-node N tools/n8_probe.py"""
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
dist.init_process_group("nccl", device_id=dev)
Q, N, D, k, K = 16384, 1_000_000, 768, 10, 20
lo, hi = shard_bounds(N, world, rank)
corpus = torch.randn(hi - lo, D, device=dev, generator=torch.Generator(device=dev).manual_seed(17 + rank))
queries = torch.randn(Q, D, device=dev, generator=torch.Generator(device=dev).manual_seed(5))


def timed(fn):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) / K


variants = {
    "stream (pipelined batches)": (dict(), "stream"),
    "loop of topk() (tail behind its own contraction)": (dict(), "loop"),
    "stream, thresholds not shared across GPUs": (dict(share_thresholds=False), "stream"),
    "loop, thresholds not shared across GPUs": (dict(share_thresholds=False), "loop"),
    "loop, NCCL exchange": (dict(exchange="nccl"), "loop"),
}
scs = {name: ShardedCorpus(corpus, N, lo, **kw) for name, (kw, _) in variants.items()}
res = {name: [] for name in variants}
for rnd in range(4):
    for name, (kw, mode) in variants.items():
        sc = scs[name]
        if mode == "stream":
            fn = lambda sc=sc: [0 for _ in sc.topk_stream((queries for _ in range(K)), k)]
        else:
            fn = lambda sc=sc: [sc.topk(queries, k) for _ in range(K)]
        if rnd == 0:
            fn()
        res[name].append(timed(fn))
if rank == 0:
    for name, v in res.items():
        print(f"[n8probe world={world} levels={os.environ.get('MMD_LEVELS', 'default')}] {name:52s} min {min(v[1:]):.3f}  all {' '.join(f'{x:.3f}' for x in v)} ms/batch", flush=True)
dist.barrier()
dist.destroy_process_group()
