"""Developer tool: repeat the K=100 search of the text-evaluation test and compare with the fp32 oracle (which rows differ?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-misinformation-detection_b200")); sys.path.insert(0, ROOT)
import torch
import mmd_retrieval as m
from mmd_retrieval import ops
from oracle import exact
def _data(kind, n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g)
n_claims = 120
train, test = _data("text", 2500, 768, 101), _data("text", n_claims, 768, 102)
claims = test + 0.9 * _data("text", n_claims, 768, 103)
train[:30] = test[:30]
for name, corp in (("train", train), ("test", test)):
    pc = ops.prepare_corpus(corp.cuda(), dtype="bf16")
    full = exact.exact_scores(claims, corp)
    bad_total = 0
    for rep in range(30):
        s, i = ops.topk(claims.cuda(), pc, 100)
        cmp = exact.compare_topk(s, i, full, 100, tie_tol=2e-6)
        if not cmp.ok:
            bad_total += 1
            want_s, want_i = exact.exact_topk(claims, corp, 100)
            ii = i.cpu()
            rows = [r for r in range(n_claims) if set(ii[r].tolist()) != set(want_i[r].tolist())]
            print(name, "rep", rep, cmp, "rows", rows[:8])
            r = rows[0]
            miss = sorted(set(want_i[r].tolist()) - set(ii[r].tolist())); extra = sorted(set(ii[r].tolist()) - set(want_i[r].tolist()))
            rank_of = {int(x): k for k, x in enumerate(want_i[r].tolist())}
            print("   row", r, "missing", miss, "at oracle ranks", [rank_of[x] for x in miss], "extra", extra, "n_unique", len(set(ii[r].tolist())))
    print(name, "levels", os.environ.get("MMD_LEVELS"), "bad reps:", bad_total, "of 30", flush=True)
