"""Batched cross-encoder re-rank stage (the stage that FOLLOWS the hot path).

Reference: src/evidence/text2text_retrieval.py:67-95 -- for ONE claim, the 2 * top_k*5 retrieved passages are scored by
`CrossEncoder("cross-encoder/ms-marco-MiniLM-L-6-v2").predict([[query, passage], ...])` (a 6-layer, 384-wide BERT sequence
classifier with one label; `predict` applies no activation for that checkpoint: the raw logit is the score), twice per
claim (train hits, test hits), one claim per call.  Here ALL (claim, passage) pairs of a whole claim batch go through
the encoder in length-bucketed batches on the GPU.

What is native and what is not: the stage is 12 attention heads of 32 dims over <= 512 tokens plus six small linear
layers per block -- not the path this repository exists for.  The linear layers (95 % of its flops) can run on this
repository's tcgen05 contraction (`gemm="mmd"`: `mmd_scores_dense`, bf16 operands, fp32 accumulation; the same kernel as
the dense score pass); attention, layer norm and GELU are PyTorch ops.  `gemm="torch"` keeps the linears in PyTorch too.
The tokenizer is NOT part of this module (the checkpoint's WordPiece vocabulary cannot be fetched here): pass any callable
`tokenize(queries, passages) -> {"input_ids", "token_type_ids", "attention_mask"}` (a Hugging Face tokenizer called with
`padding=True, truncation=True, return_tensors="pt"` has exactly that shape).

Parity: `oracle/cross_encoder.py` builds the Hugging Face `BertForSequenceClassification` from the same state dict and
runs it in fp32 on the host; tests/test_cross_encoder.py compares logits (random, seeded weights of the MiniLM-L6
geometry: the checkpoint itself is not available offline).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import _lib


@dataclass(frozen=True)
class EncoderConfig:
    """Geometry of cross-encoder/ms-marco-MiniLM-L-6-v2 (config.json of the checkpoint)."""
    vocab_size: int = 30522
    hidden: int = 384
    layers: int = 6
    heads: int = 12
    intermediate: int = 1536
    max_positions: int = 512
    type_vocab: int = 2
    ln_eps: float = 1e-12


def random_state_dict(cfg: EncoderConfig = EncoderConfig(), seed: int = 0, std: float = 0.05) -> Dict[str, torch.Tensor]:
    """A seeded state dict with the Hugging Face BERT key names (for tests and benchmarks: no checkpoint can be downloaded here)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def mat(name, *shape):
        sd[name] = torch.randn(*shape, generator=g) * std

    def ln(prefix):
        sd[prefix + ".weight"] = 1.0 + 0.1 * torch.randn(cfg.hidden, generator=g)
        sd[prefix + ".bias"] = 0.1 * torch.randn(cfg.hidden, generator=g)

    mat("bert.embeddings.word_embeddings.weight", cfg.vocab_size, cfg.hidden)
    mat("bert.embeddings.position_embeddings.weight", cfg.max_positions, cfg.hidden)
    mat("bert.embeddings.token_type_embeddings.weight", cfg.type_vocab, cfg.hidden)
    ln("bert.embeddings.LayerNorm")
    for i in range(cfg.layers):
        p = f"bert.encoder.layer.{i}."
        for nm in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
            mat(p + nm + ".weight", cfg.hidden, cfg.hidden)
            mat(p + nm + ".bias", cfg.hidden)
        ln(p + "attention.output.LayerNorm")
        mat(p + "intermediate.dense.weight", cfg.intermediate, cfg.hidden)
        mat(p + "intermediate.dense.bias", cfg.intermediate)
        mat(p + "output.dense.weight", cfg.hidden, cfg.intermediate)
        mat(p + "output.dense.bias", cfg.hidden)
        ln(p + "output.LayerNorm")
    mat("bert.pooler.dense.weight", cfg.hidden, cfg.hidden)
    mat("bert.pooler.dense.bias", cfg.hidden)
    mat("classifier.weight", 1, cfg.hidden)
    mat("classifier.bias", 1)
    return sd


class _Linear:
    """y = x W^T + b.  gemm="mmd": W is prepared once as bf16 operand tiles and x W^T runs on the tcgen05 contraction."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, device, gemm: str, compute_dtype: torch.dtype):
        self.gemm = gemm
        self.bias = bias.to(device=device, dtype=torch.float32)
        if gemm == "mmd":
            from . import ops
            self.prepared = ops.prepare_corpus(weight.to(device=device, dtype=torch.float32).contiguous(), dtype="bf16", metric="dot",
                                               keep_source=False)
        else:
            self.weight = weight.to(device=device, dtype=compute_dtype)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self.gemm == "mmd":
            from . import ops
            lead = x.shape[:-1]
            y = ops.dense_scores(x.reshape(-1, x.shape[-1]).float().contiguous(), self.prepared)      # fp32 [tokens, out]
            return (y + self.bias).reshape(*lead, -1)
        return (F.linear(x.to(self.weight.dtype), self.weight).float() + self.bias)


def _forward(cfg: EncoderConfig, w: Dict[str, object], input_ids: torch.Tensor, token_type_ids: torch.Tensor,
             attention_mask: torch.Tensor) -> torch.Tensor:
    """The encoder proper: logits [B] for token batches [B, L].  Activations are fp32 between the linear layers (residual
    stream, layer norm, softmax statistics); device-agnostic torch -- the public class below insists on CUDA."""
    B, L = input_ids.shape
    pos = torch.arange(L, device=input_ids.device)
    x = w["word"][input_ids] + w["pos"][pos][None, :, :] + w["type"][token_type_ids]
    x = F.layer_norm(x, (cfg.hidden,), w["emb_ln_w"], w["emb_ln_b"], cfg.ln_eps)
    hd = cfg.hidden // cfg.heads
    # additive key mask: padded keys get -inf (every row has at least its [CLS] token)
    key_mask = torch.zeros((B, 1, 1, L), dtype=torch.float32, device=x.device).masked_fill(~attention_mask.bool()[:, None, None, :], float("-inf"))
    for layer in w["layers"]:
        qkv = layer["qkv"](x)                                                         # [B, L, 3 * hidden]
        q, k, v = (t.reshape(B, L, cfg.heads, hd).transpose(1, 2) for t in qkv.split(cfg.hidden, dim=-1))
        att = F.scaled_dot_product_attention(q, k, v, attn_mask=key_mask)             # [B, heads, L, hd]
        att = att.transpose(1, 2).reshape(B, L, cfg.hidden)
        x = F.layer_norm(layer["attn_out"](att) + x, (cfg.hidden,), layer["ln1_w"], layer["ln1_b"], cfg.ln_eps)
        h = F.gelu(layer["ffn_in"](x))
        x = F.layer_norm(layer["ffn_out"](h) + x, (cfg.hidden,), layer["ln2_w"], layer["ln2_b"], cfg.ln_eps)
    pooled = torch.tanh(w["pooler"](x[:, 0]))
    return w["classifier"](pooled).reshape(B)


def _load(cfg: EncoderConfig, sd: Dict[str, torch.Tensor], device, gemm: str, compute_dtype: torch.dtype) -> Dict[str, object]:
    f32 = lambda name: sd[name].to(device=device, dtype=torch.float32)                # noqa: E731
    w: Dict[str, object] = {
        "word": f32("bert.embeddings.word_embeddings.weight"), "pos": f32("bert.embeddings.position_embeddings.weight"),
        "type": f32("bert.embeddings.token_type_embeddings.weight"),
        "emb_ln_w": f32("bert.embeddings.LayerNorm.weight"), "emb_ln_b": f32("bert.embeddings.LayerNorm.bias"), "layers": [],
    }
    for i in range(cfg.layers):
        p = f"bert.encoder.layer.{i}."
        qkv_w = torch.cat([sd[p + f"attention.self.{n}.weight"] for n in ("query", "key", "value")], dim=0)      # one GEMM for Q, K, V
        qkv_b = torch.cat([sd[p + f"attention.self.{n}.bias"] for n in ("query", "key", "value")], dim=0)
        w["layers"].append({
            "qkv": _Linear(qkv_w, qkv_b, device, gemm, compute_dtype),
            "attn_out": _Linear(sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"], device, gemm, compute_dtype),
            "ln1_w": f32(p + "attention.output.LayerNorm.weight"), "ln1_b": f32(p + "attention.output.LayerNorm.bias"),
            "ffn_in": _Linear(sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"], device, gemm, compute_dtype),
            "ffn_out": _Linear(sd[p + "output.dense.weight"], sd[p + "output.dense.bias"], device, gemm, compute_dtype),
            "ln2_w": f32(p + "output.LayerNorm.weight"), "ln2_b": f32(p + "output.LayerNorm.bias"),
        })
    # the two single-row heads stay in PyTorch: B x 384 x 384 and B x 384 x 1
    w["pooler"] = _Linear(sd["bert.pooler.dense.weight"], sd["bert.pooler.dense.bias"], device, "torch", torch.float32)
    w["classifier"] = _Linear(sd["classifier.weight"], sd["classifier.bias"], device, "torch", torch.float32)
    return w


class BatchedCrossEncoder:
    """`predict(pairs)` / `score_tokens(...)` over whole batches of (claim, passage) pairs on the GPU.

    state_dict: Hugging Face BERT-for-sequence-classification weights (one label).  tokenize: see the module docstring.
    gemm: "mmd" (this repository's tcgen05 contraction for the linear layers, bf16 operands) or "torch".
    max_tokens: padded tokens per forward batch (pairs are sorted by length and cut into batches of at most this many).
    """

    def __init__(self, state_dict: Dict[str, torch.Tensor], tokenize: Optional[Callable] = None, cfg: EncoderConfig = EncoderConfig(),
                 device=None, gemm: str = "mmd", compute_dtype: torch.dtype = torch.bfloat16, max_tokens: int = 65536):
        if gemm not in ("mmd", "torch"):
            raise ValueError("gemm must be 'mmd' or 'torch'")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise _lib.MmdError("the batched cross-encoder needs a CUDA device (no CPU path)")
        self.cfg, self.device, self.tokenize, self.max_tokens, self.gemm = cfg, dev, tokenize, int(max_tokens), gemm
        self.w = _load(cfg, state_dict, dev, gemm, compute_dtype)

    @torch.no_grad()
    def score_tokens(self, input_ids: torch.Tensor, token_type_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        """Logits f32 [B] (on the device) of already tokenised pairs [B, L]; pairs are re-batched by length."""
        ids, tt, am = (t.to(self.device) for t in (input_ids, token_type_ids, attention_mask))
        n = ids.shape[0]
        out = torch.empty((n,), dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        lengths = am.long().sum(dim=1).clamp_min(1)
        order = torch.argsort(lengths, descending=True)
        lens_sorted = lengths[order].tolist()
        start = 0
        while start < n:
            width = min(int(lens_sorted[start]), self.cfg.max_positions)
            count = max(1, min(n - start, self.max_tokens // max(width, 1)))
            sel = order[start:start + count]
            out[sel] = _forward(self.cfg, self.w, ids[sel, :width], tt[sel, :width], am[sel, :width])
            start += count
        return out

    def predict(self, pairs: Sequence[Tuple[str, str]]) -> List[float]:
        """CrossEncoder.predict for a list of [query, passage] pairs (any number, any mix of queries)."""
        if self.tokenize is None:
            raise RuntimeError("no tokenizer attached: pass tokenize=..., or call score_tokens() with token tensors")
        if len(pairs) == 0:
            return []
        enc = self.tokenize([p[0] for p in pairs], [p[1] for p in pairs])
        return self.score_tokens(enc["input_ids"], enc["token_type_ids"], enc["attention_mask"]).cpu().tolist()

    # the (query, texts) -> scores callable SemanticSimilarity(cross_encoder=...) accepts; SemanticSimilarity.search_batch uses
    # predict() directly when it finds one, so that the pairs of ALL claims of a batch go through the encoder together
    def __call__(self, query: str, texts: Sequence[str]) -> List[float]:
        return self.predict([(query, t) for t in texts])


def hashing_tokenizer(vocab_size: int = 30522, max_length: int = 512) -> Callable:
    """A stand-in tokenizer for synthetic runs (no WordPiece vocabulary is available offline): whitespace tokens hashed
    into the vocabulary, [CLS] a [SEP] b [SEP] layout, token types 0 / 1, right padding -- the tensor layout of a BERT pair
    encoding, NOT the checkpoint's segmentation."""
    cls_id, sep_id = 101, 102

    def ids_of(text: str) -> List[int]:
        return [1000 + (hash_str(tok) % (vocab_size - 1000)) for tok in text.split()]

    def tokenize(queries: Sequence[str], passages: Sequence[str]) -> Dict[str, torch.Tensor]:
        rows, types = [], []
        for a, b in zip(queries, passages):
            ia, ib = ids_of(a), ids_of(b)
            room = max_length - 3
            ia = ia[: max(1, room // 2)] if len(ia) + len(ib) > room else ia
            ib = ib[: room - len(ia)]
            rows.append([cls_id] + ia + [sep_id] + ib + [sep_id])
            types.append([0] * (len(ia) + 2) + [1] * (len(ib) + 1))
        width = max(len(r) for r in rows)
        ids = torch.zeros((len(rows), width), dtype=torch.long)
        tt = torch.zeros_like(ids)
        am = torch.zeros_like(ids)
        for i, (r, t) in enumerate(zip(rows, types)):
            ids[i, :len(r)] = torch.tensor(r)
            tt[i, :len(t)] = torch.tensor(t)
            am[i, :len(r)] = 1
        return {"input_ids": ids, "token_type_ids": tt, "attention_mask": am}

    return tokenize


def hash_str(s: str) -> int:
    """Deterministic across processes (Python's hash() is salted): FNV-1a."""
    h = 0xcbf29ce484222325
    for ch in s.encode("utf-8"):
        h = ((h ^ ch) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def flops_per_pair(cfg: EncoderConfig, length: int) -> float:
    """Multiply-add flops (x2) of one pair of `length` tokens: linear layers + attention."""
    lin = cfg.layers * (4 * cfg.hidden * cfg.hidden + 2 * cfg.hidden * cfg.intermediate)
    att = cfg.layers * 2 * length * cfg.hidden
    return 2.0 * length * (lin + att) + 2.0 * cfg.hidden * (cfg.hidden + 1)

