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
