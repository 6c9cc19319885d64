"""Developer tool: run one workload a few times with the MMD_STATS build to dump the per-tile timeline."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-misinformation-detection_b200"))
import torch
import mmd_retrieval as m
Q, N, D, k = [int(x) for x in sys.argv[1:5]]
g = torch.Generator(device="cuda").manual_seed(1)
kind = sys.argv[5] if len(sys.argv) > 5 else "text"
q = torch.randn(Q, D, device="cuda", generator=g)
c = torch.randn(N, D, device="cuda", generator=g)
if kind == "image":
    q, c = torch.relu(q), torch.relu(c)
op = sys.argv[6] if len(sys.argv) > 6 else "bf16"
pc = m.prepare_corpus(c, dtype=op, keep_source=False)
for _ in range(5):
    m.topk(q, pc, k, rescore_exact=False)
torch.cuda.synchronize()
