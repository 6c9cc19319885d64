"""The cross-encoder re-rank stage (src/evidence/text2text_retrieval.py:67-95) against the Hugging Face
BertForSequenceClassification the reference's CrossEncoder wraps (oracle/cross_encoder.py), on seeded random weights of the
MiniLM-L6 geometry (the checkpoint cannot be fetched here).  CPU: the encoder arithmetic in fp32.  GPU: the batched class with
the linear layers on the tcgen05 contraction (gemm="mmd") and in PyTorch (gemm="torch"), bf16 operands."""
import pytest
import torch

import mmd_retrieval as m
from mmd_retrieval import cross_encoder as ce
from oracle import cross_encoder as oce


def _tokens(cfg, n_pairs, max_len, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, cfg.vocab_size, (n_pairs, max_len), generator=g)
    lens = torch.randint(5, max_len + 1, (n_pairs,), generator=g)
    lens[0] = max_len
    am = (torch.arange(max_len)[None, :] < lens[:, None]).long()
    tt = ((torch.arange(max_len)[None, :] > (lens[:, None] // 3)) & am.bool()).long()
    ids = ids * am                                                 # [PAD] = 0 behind the pair
    ids[:, 0] = 101
    return ids, tt, am


def test_encoder_arithmetic_matches_hf_bert_fp32():
    cfg = ce.EncoderConfig(vocab_size=3000)                        # MiniLM-L6 geometry, small vocabulary
    sd = ce.random_state_dict(cfg, seed=5)
    model = oce.hf_model(sd, cfg.vocab_size, cfg.hidden, cfg.layers, cfg.heads, cfg.intermediate)
    ids, tt, am = _tokens(cfg, 24, 80, 9)
    want = oce.predict_logits(model, ids, tt, am)
    w = ce._load(cfg, sd, "cpu", "torch", torch.float32)
    got = ce._forward(cfg, w, ids, tt, am)
    assert float((got - want).abs().max()) <= 2e-5, float((got - want).abs().max())
    # padding does not leak: a pair scored alone at its own length gives the same logit
    n5 = int(am[5].sum())
    alone = ce._forward(cfg, w, ids[5:6, :n5], tt[5:6, :n5], am[5:6, :n5])
    assert abs(float(alone) - float(want[5])) <= 2e-5
    # order of the batch does not matter
    perm = torch.randperm(24, generator=torch.Generator().manual_seed(1))
    assert float((ce._forward(cfg, w, ids[perm], tt[perm], am[perm]) - want[perm]).abs().max()) <= 2e-5


def test_hashing_tokenizer_layout_and_no_cpu_path():
    tok = ce.hashing_tokenizer(3000, 32)
    enc = tok(["a b c", "x " * 40], ["d e", "y " * 40])
    ids, tt, am = enc["input_ids"], enc["token_type_ids"], enc["attention_mask"]
    assert ids.shape == (2, 32) and int(am[0].sum()) == 3 + 3 + 2 and int(am[1].sum()) == 32
    assert ids[0, 0] == 101 and ids[0, 4] == 102 and tt[0, :5].sum() == 0 and tt[0, 5:8].sum() == 3
    assert tok(["a b c"], ["d e"])["input_ids"].tolist() == tok(["a b c"], ["d e"])["input_ids"].tolist()
    if not torch.cuda.is_available():
        with pytest.raises(m.MmdError):
            m.BatchedCrossEncoder(ce.random_state_dict(ce.EncoderConfig(vocab_size=2000, layers=1)), cfg=ce.EncoderConfig(vocab_size=2000, layers=1))


def test_search_batch_uses_batched_predict():
    """SemanticSimilarity hands ALL pairs of a claim batch to a cross-encoder that offers predict(pairs) (host logic only:
    the retrieval itself is stubbed out here, it needs the GPU)."""
    from mmd_retrieval import text_corpus

    class FakePC:
        def __init__(self, n):
            self.n = n

    calls = []

    class Rer:
        def predict(self, pairs):
            calls.append(len(pairs))
            return [float(len(t)) + 0.001 * len(q) for q, t in pairs]

    ss = text_corpus.SemanticSimilarity.__new__(text_corpus.SemanticSimilarity)
    ss.train, ss.test = FakePC(4), FakePC(3)
    ss.train_ids, ss.test_ids = [f"train_{i}".encode() for i in range(4)], [f"test_{i}".encode() for i in range(3)]
    ss.bi_encoder, ss.cross_encoder = None, Rer()
    ss.train_texts, ss.test_texts = ["t" * (i + 1) for i in range(4)], ["u" * (10 + i) for i in range(3)]
    real = text_corpus.ops.topk
    try:
        text_corpus.ops.topk = lambda emb, corpus, k, **kw: (torch.linspace(1, 0, corpus.n).repeat(emb.shape[0], 1),
                                                            torch.arange(corpus.n).repeat(emb.shape[0], 1))
        out = ss.search_batch(torch.zeros(2, 8), top_k=2, query_texts=["q", "qq"])
    finally:
        text_corpus.ops.topk = real
    assert calls == [2 * (4 + 3)]
    assert [k for k, _ in out[0]] == ["test_2", "test_1"] and out[0][0][1] == pytest.approx(12.001)
    assert out[1][0][1] == pytest.approx(12.002)


@pytest.mark.gpu
@pytest.mark.parametrize("gemm", ["torch", "mmd"])
def test_batched_cross_encoder_on_gpu(gemm):
    cfg = ce.EncoderConfig(vocab_size=3000)
    sd = ce.random_state_dict(cfg, seed=6)
    model = oce.hf_model(sd, cfg.vocab_size, cfg.hidden, cfg.layers, cfg.heads, cfg.intermediate)
    ids, tt, am = _tokens(cfg, 40, 96, 10)
    want = oce.predict_logits(model, ids, tt, am)
    enc = m.BatchedCrossEncoder(sd, cfg=cfg, gemm=gemm, max_tokens=1024)            # several length buckets
    got = enc.score_tokens(ids, tt, am).cpu()
    err = float((got - want).abs().max())
    # bf16 operands in 36 linear layers: ~1e-2 absolute on logits of spread ~0.15 (fp32 reference); fp32 accumulation
    assert err <= 5e-2, (gemm, err, float(want.std()))
    assert float(torch.corrcoef(torch.stack([got, want]))[0, 1]) >= 0.99
    # the same pairs as text through a tokenizer, and the (query, texts) callable SemanticSimilarity accepts
    enc2 = m.BatchedCrossEncoder(sd, tokenize=ce.hashing_tokenizer(cfg.vocab_size, 64), cfg=cfg, gemm=gemm)
    texts = ["flood waters rose in the old town " * (1 + i % 3) for i in range(7)]
    a = enc2.predict([("did the town flood", t) for t in texts])
    b = enc2("did the town flood", texts)
    assert len(a) == 7 and a == b
