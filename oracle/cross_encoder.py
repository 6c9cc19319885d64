"""TEST INFRASTRUCTURE -- CPU oracle of the cross-encoder re-rank stage.  Only tests/ may import this.

Reference: src/evidence/text2text_retrieval.py:25,67-95 -- `CrossEncoder("cross-encoder/ms-marco-MiniLM-L-6-v2").predict(pairs)`.
sentence-transformers (3.3.1, un-vendored) wraps `transformers.AutoModelForSequenceClassification` (here: BertForSequenceClassification,
one label) and, for this checkpoint, applies no activation to the logit (`sbert_ce_default_activation_function: Identity` in its
config.json).  So the oracle IS the Hugging Face model: built from a config of the checkpoint's geometry, loaded with the state
dict under test, run in fp32 on the host.  The checkpoint's weights and vocabulary cannot be fetched here (no network):
parity is pinned on the reference's own model CODE (transformers is installed) with seeded random weights.
"""
from __future__ import annotations

from typing import Dict

import torch


def hf_model(state_dict: Dict[str, torch.Tensor], vocab_size: int, hidden: int, layers: int, heads: int, intermediate: int,
             max_positions: int = 512, type_vocab: int = 2, ln_eps: float = 1e-12):
    from transformers import BertConfig, BertForSequenceClassification
    cfg = BertConfig(vocab_size=vocab_size, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                     intermediate_size=intermediate, max_position_embeddings=max_positions, type_vocab_size=type_vocab,
                     layer_norm_eps=ln_eps, hidden_act="gelu", num_labels=1, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = BertForSequenceClassification(cfg).eval()
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not [k for k in missing if "position_ids" not in k], missing
    assert not unexpected, unexpected
    return model


@torch.no_grad()
def predict_logits(model, input_ids: torch.Tensor, token_type_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """What CrossEncoder.predict returns for this checkpoint: the raw logit of every pair, fp32."""
    out = model(input_ids=input_ids, token_type_ids=token_type_ids, attention_mask=attention_mask)
    return out.logits.reshape(-1).float()
