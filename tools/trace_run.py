"""Developer tool: run one workload a few times with an MMD_STATS build to dump the per-tile timeline of block 0.
MMD_LIB_PATH=.../libmmd_stats_w64.so python tools/trace_run.py Q N D k [text|image] [bf16|fp8] [topk|filter]
mode "filter": every pruning threshold preset to +inf, i.e. the epilogue runs its filter pass only (its floor per tile)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-misinformation-detection_b200"))
import torch
import mmd_retrieval as m
from mmd_retrieval import ops, _lib
Q, N, D, k = [int(x) for x in sys.argv[1:5]]
g = torch.Generator(device="cuda").manual_seed(1)
kind = sys.argv[5] if len(sys.argv) > 5 else "text"
q = torch.randn(Q, D, device="cuda", generator=g)
c = torch.randn(N, D, device="cuda", generator=g)
if kind == "image":
    q, c = torch.relu(q), torch.relu(c)
op = sys.argv[6] if len(sys.argv) > 6 else "bf16"
mode = sys.argv[7] if len(sys.argv) > 7 else "topk"
pc = m.prepare_corpus(c, dtype=op, keep_source=False)
print(f"[trace] {_lib.build_info()} Q={Q} N={N} D={D} k={k} {kind} {op} {mode}", file=sys.stderr)
if mode == "filter":
    q_rows, _ = ops.normalize_cast(q, op, _lib.SIDE_QUERY, True, 1e-12)
    thr = torch.full((Q,), -8388608, dtype=torch.int32, device="cuda")          # ordered(+inf) = 0xff800000
    for _ in range(5):
        ops.topk_prepared(q_rows, Q, pc, k, shared_thr=(thr.data_ptr(), [thr.data_ptr()]))
else:
    for _ in range(5):
        m.topk(q, pc, k, rescore_exact=False)
torch.cuda.synchronize()
