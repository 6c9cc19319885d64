"""GPU (B200): BASELINE.json's full-size configurations through size-independent properties, plus sampled
exact parity against the CPU oracle (a full 4k x 50k x 2048 float64 oracle would take minutes)."""
import pytest
import torch

from oracle import exact

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import mmd_retrieval
    return mmd_retrieval


def _gen(kind, rows, dim, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(rows, dim, device="cuda", generator=g)
    return torch.relu(x) if kind == "image" else x


def _check_properties(m, kind, q_n, c_n, dim, k, eps, seed):
    c = _gen(kind, c_n, dim, seed)
    q = _gen(kind, q_n, dim, seed + 1)
    planted = torch.randperm(c_n, device="cuda")[:q_n]
    q = c[planted] + 0.25 * q                                   # planted positives: row planted[j] must be rank 0 of query j
    if kind == "image":
        q = torch.relu(q)
    pc = m.prepare_corpus(c, dtype="bf16", eps=eps)
    s, i = m.topk(q, pc, k)
    # sorted, in range, unique
    assert bool((s[:, :-1] >= s[:, 1:]).all())
    assert int(i.min()) >= 0 and int(i.max()) < c_n
    assert bool((torch.sort(i, dim=1).values.diff(dim=1) > 0).all())
    # planted positive found first (hits@1 == 1)
    assert torch.equal(i[:, 0], planted)
    # prefix property: top-(k/2) is the head of top-k
    s2, i2 = m.topk(q, pc, k // 2)
    assert torch.equal(i2, i[:, : k // 2]) and torch.equal(s2, s[:, : k // 2])
    # scores are the fp32 cosine of the returned rows (recomputed in float64 on the device)
    qn = torch.nn.functional.normalize(q.double(), dim=1, eps=eps)
    sub = torch.randperm(q_n, device="cuda")[:256]
    cn = torch.nn.functional.normalize(c[i[sub]].double(), dim=2, eps=eps)
    want = torch.einsum("qd,qkd->qk", qn[sub], cn)
    assert float(((s[sub].double() - want).abs() / want.abs().clamp_min(1e-6)).max()) <= 1e-5
    # permutation invariance: shuffling the corpus rows permutes indices and leaves the score lists unchanged
    perm = torch.randperm(c_n, device="cuda")
    s3, i3 = m.topk(q[sub], m.prepare_corpus(c[perm], dtype="bf16", eps=eps), k)
    assert torch.equal(perm[i3], i[sub]) or float((s3 - s[sub]).abs().max()) <= 1e-6
    assert float((s3 - s[sub]).abs().max()) <= 1e-6
    # sampled exact parity with the CPU oracle
    pick = sub[:32].cpu()
    full = exact.exact_scores(q.cpu()[pick], c.cpu(), "cos", eps)
    cmp = exact.compare_topk(s.cpu()[pick], i.cpu()[pick], full, k, tie_tol=2e-6)
    assert cmp.ok and cmp.max_rel_score_err <= 1e-5, cmp


def test_config2_im2im_4k_x_50k_x_2048(m):
    _check_properties(m, "image", 4096, 50000, 2048, 10, 1e-6, 2)


def test_config3_text_16k_x_1m_x_768(m):
    _check_properties(m, "text", 16384, 1_000_000, 768, 10, 1e-12, 3)


def test_config1_text_1k_x_10k_fp32_full_oracle(m):
    g = torch.Generator().manual_seed(1)
    q, c = torch.randn(1000, 768, generator=g), torch.randn(10000, 768, generator=g)
    s, i = m.topk(q.cuda(), m.prepare_corpus(c.cuda(), dtype="fp32", keep_source=False), 5, rescore_exact=False)
    full = exact.exact_scores(q, c)
    cmp = exact.compare_topk(s, i, full, 5, tie_tol=2e-6)
    assert cmp.ok and cmp.max_rel_score_err <= 1e-5, cmp
    assert cmp.identical_order >= 995


def test_config4_joint_one_gpu_share_of_10m_x_512x2(m):
    """configs[3] as one of 8 GPUs sees it: 1.25 M corpus rows x (512 text + 512 image), 16384 query pairs, fused top-10.
    Properties: planted positives first, sorted, fused scores recomputed in float64, sampled parity with the fusion oracle."""
    from oracle import fusion
    n_c, n_q, k, w = 1_250_000, 16384, 10, (0.5, 0.5)
    ct, ci = _gen("text", n_c, 512, 4), _gen("text", n_c, 512, 5)
    planted = torch.randperm(n_c, device="cuda")[:n_q]
    qt = ct[planted] + 0.5 * _gen("text", n_q, 512, 6)
    qi = ci[planted] + 0.5 * _gen("text", n_q, 512, 7)
    jc = m.prepare_joint([ct, ci], w, dtype="bf16")
    s, i = m.topk_joint([qt, qi], jc, k)
    assert bool((s[:, :-1] >= s[:, 1:]).all()) and torch.equal(i[:, 0], planted)
    sub = torch.randperm(n_q, device="cuda")[:128]
    nrm = lambda x: torch.nn.functional.normalize(x.double(), dim=-1, eps=1e-12)   # noqa: E731
    want = 0.5 * torch.einsum("qd,qkd->qk", nrm(qt[sub]), nrm(ct[i[sub]])) + 0.5 * torch.einsum("qd,qkd->qk", nrm(qi[sub]), nrm(ci[i[sub]]))
    assert float(((s[sub].double() - want).abs() / want.abs().clamp_min(1e-6)).max()) <= 1e-5
    pick = sub[:8].cpu()
    full = fusion.fused_scores([qt.cpu()[pick], qi.cpu()[pick]], [ct.cpu(), ci.cpu()], w)
    cmp = exact.compare_topk(s.cpu()[pick], i.cpu()[pick], full, k, tie_tol=2e-6)
    assert cmp.ok, cmp


def test_config5_fp8_one_gpu_share_streamed_top100(m):
    """configs[4] as one of 8 GPUs sees it, scaled to fit the test budget: a corpus shard streamed chunk-wise through K1
    straight to fp8 tiles (never resident in fp32), fp16 source kept for the exact re-score, 8192 queries, top-100.
    Properties: planted positives first, sorted, unique, prefix property; recall@100 against a float64 ranking of
    sampled queries (BASELINE.json states no fp8 tolerance: e4m3 operand noise is ~1e-3 per score against a
    rank-100 spacing of ~1e-4 in a 4 M-row Gaussian corpus, so the bar is recall, not set equality)."""
    n_c, n_q, dim, k, chunk = 4_000_000, 8192, 768, 100, 500_000
    planted = torch.randperm(n_c, device="cuda")[:n_q]
    keep = {}

    def chunks():
        for lo in range(0, n_c, chunk):
            x = _gen("text", min(chunk, n_c - lo), dim, 100 + lo // chunk)
            sel = (planted >= lo) & (planted < lo + x.shape[0])
            keep[lo] = (sel.nonzero().flatten(), x[planted[sel] - lo].clone())
            yield x

    pc = m.prepare_streamed(chunks(), n_c, dim, dtype="fp8", keep_source=torch.float16)
    assert pc.rows.shape == (n_c, 768) and pc.source.dtype == torch.float16
    q = torch.empty(n_q, dim, device="cuda")
    for sel, rows in keep.values():
        q[sel] = rows
    q = q + 0.3 * _gen("text", n_q, dim, 999)
    s, i = m.topk(q, pc, k)
    assert tuple(i.shape) == (n_q, k) and bool((s[:, :-1] >= s[:, 1:]).all())
    assert bool((torch.sort(i, dim=1).values.diff(dim=1) > 0).all())
    assert torch.equal(i[:, 0], planted)
    s2, i2 = m.topk(q, pc, 50)
    agree = (i2 == i[:, :50]).float().mean().item()
    assert agree >= 0.99           # prefix property up to what the fp8 selection (75 vs 104 candidates) lets through
    # recall of the fp8 selection against the float64 ranking of the fp16-stored corpus, sampled queries
    sub = torch.randperm(n_q, device="cuda")[:16]
    qn = torch.nn.functional.normalize(q[sub].double(), dim=1)
    best = torch.full((16, k), -2.0, dtype=torch.float64, device="cuda")
    best_i = torch.zeros((16, k), dtype=torch.int64, device="cuda")
    for lo in range(0, n_c, chunk):
        cn = torch.nn.functional.normalize(pc.source[lo:lo + chunk].double(), dim=1)
        sc = qn @ cn.T
        v, j = torch.topk(sc, k, dim=1)
        allv, alli = torch.cat([best, v], 1), torch.cat([best_i, j + lo], 1)
        o = torch.argsort(allv, dim=1, descending=True)[:, :k]
        best, best_i = torch.gather(allv, 1, o), torch.gather(alli, 1, o)
    recall = sum(len(set(a) & set(b)) for a, b in zip(i[sub].cpu().tolist(), best_i.cpu().tolist())) / (16 * k)
    print(f"fp8 top-100 recall vs float64: {recall:.4f}")
    assert recall >= 0.90, recall
