"""Developer tool: fused-kernel time of a few corner shapes under the current tuning knobs / library (one process per
setting: the knobs are read once).  python tools/epi_sweep.py [case ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-misinformation-detection_b200"))
import torch
import mmd_retrieval as m

CASES = {  # name: Q, N, D, k (list length kept by the kernel), kind, op
    "c3_k18": (16384, 1000000, 768, 18, "text", "bf16"),
    "c2_k18": (4096, 50000, 2048, 18, "image", "bf16"),
    "fp8_k18": (16384, 1000000, 768, 18, "text", "fp8"),
    "bf16_k100": (16384, 1000000, 768, 100, "text", "bf16"),
    "fp8_k100": (16384, 1000000, 768, 100, "text", "fp8"),
    "fp8_k104_4m": (16384, 4000000, 768, 104, "text", "fp8"),
    "c3_n8share": (16384, 125000, 768, 18, "text", "bf16"),     # what one of 8 GPUs sees of C3
    "q100": (100, 1000000, 768, 18, "text", "bf16"),
    "q1": (1, 1000000, 768, 18, "text", "bf16"),
}
names = sys.argv[1:] or list(CASES)
tag = os.environ.get("SWEEP_TAG", "default")
for name in names:
    Q, N, D, k, kind, op = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(Q, D, device="cuda", generator=g)
    c = torch.randn(N, D, device="cuda", generator=g)
    if kind == "image":
        q, c = torch.relu(q), torch.relu(c)
    pc = m.prepare_corpus(c, dtype=op, keep_source=False)
    del c
    short = Q * N * D < 4e12
    for _ in range(40 if short else 3):
        m.topk(q, pc, k, rescore_exact=False)
    torch.cuda.synchronize()
    m.profile_enable(True)
    n = 40 if short else 6
    for _ in range(n):
        s, i = m.topk(q, pc, k, rescore_exact=False)
    torch.cuda.synchronize()
    fused = m.profile_collect()
    m.profile_enable(False)
    fm = sorted(fused)[len(fused) // 2]
    print(f"[sweep {tag:28s}] {name:12s} fused median {fm:8.3f} ms = {2.0 * Q * N * D / fm / 1e9:6.0f} TFLOP/s  (min {min(fused):.3f})  chk {float(s.double().sum()):.6f} {int(i.long().sum())}", flush=True)
    del pc
