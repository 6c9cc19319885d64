"""mmd_retrieval -- B200-native evidence retrieval (top-K cosine / inner product) for
sakdag/multimodal-misinformation-detection's `src/evidence` path.

Public surface (mirrors the reference's entry points; see DESIGN.md / INTEGRATION.md):

    topk(queries, corpus, k, metric, dtype)        raw tensor contract  (scores, indices)
    prepare_corpus(...) -> PreparedCorpus          normalise + cast once, keep resident in HBM
    semantic_search(q, corpus, top_k=..)           drop-in for sentence_transformers.util.semantic_search
    cos_sim / dot_score                            drop-in score functions
    ImageCorpus / ImageSimilarity                  drop-in for src/evidence/im2im_retrieval.py
    SemanticSimilarity                             drop-in for src/evidence/text2text_retrieval.py's search (encoders pluggable)
    ShardedCorpus                                  row-sharded corpus over the GPUs of one box
    prepare_joint / topk_joint                     joint image+text retrieval, weighted score fusion in one contraction
    BatchedCrossEncoder                            the re-rank stage behind SemanticSimilarity(cross_encoder=...), batched on the GPU

All arithmetic runs in libmmd.so (hand-written sm_100a CUDA behind a C ABI).  There is no CPU fallback.
"""
from ._lib import MmdError, LIB_PATH
from .ops import (PreparedCorpus, prepare_corpus, topk, dense_scores, merge_topk, normalize_cast, max_k,
                  profile_enable, profile_collect, launch_count)
from .postfilter import dedupe_by_score, hits_at_k
from .semantic_search import semantic_search, cos_sim, dot_score, clear_cache
from .image_corpus import ImageCorpus, ImageSimilarity, calculate_topk_accuracy_image_retrieval
from .text_corpus import SemanticSimilarity, calculate_topk_accuracy_text_retrieval
from .sharded import ShardedCorpus, shard_bounds
from .joint import JointCorpus, prepare_joint, topk_joint
from .corpus_io import prepare_streamed, load_text_corpus, load_image_corpus
from .cross_encoder import BatchedCrossEncoder, EncoderConfig

__all__ = [
    "MmdError", "LIB_PATH", "PreparedCorpus", "prepare_corpus", "topk", "dense_scores", "merge_topk", "normalize_cast",
    "max_k", "profile_enable", "profile_collect", "launch_count", "dedupe_by_score", "hits_at_k", "semantic_search",
    "cos_sim", "dot_score", "clear_cache", "ImageCorpus", "ImageSimilarity", "SemanticSimilarity", "calculate_topk_accuracy_text_retrieval", "calculate_topk_accuracy_image_retrieval",
    "ShardedCorpus", "shard_bounds", "JointCorpus", "prepare_joint", "topk_joint", "prepare_streamed", "load_text_corpus", "load_image_corpus",
    "BatchedCrossEncoder", "EncoderConfig",
]
