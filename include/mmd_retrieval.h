/*
 * mmd_retrieval.h -- C ABI of libmmd.so, the B200-native (sm_100a) evidence-retrieval hot path.
 *
 * This is the drop-in boundary for the ONE path this repo accelerates: the embedding-similarity
 * evidence retrieval of sakdag/multimodal-misinformation-detection (paths below are relative to the
 * reference checkout):
 *
 *   src/evidence/im2im_retrieval.py:38-42     ImageSimilarity.similarity  (pairwise cosine, eps=1e-6)
 *   src/evidence/im2im_retrieval.py:80-106    ImageCorpus.retrieve_similar_images (score all, sort, dedupe)
 *   src/evidence/text2text_retrieval.py:56-64 util.semantic_search(q, corpus, top_k=...)  (third-party
 *                                             sentence-transformers==3.3.1: normalise -> mm -> topk -> heap)
 *   src/evidence/experiment_text.py:25-33, src/evidence/experiment_image.py:25-33   same, in the eval loops
 *
 * The reference is pure Python and has no FFI of its own; this header is what a ctypes binding on the
 * reference side would load (see INTEGRATION.md for that stub).  All pointers are DEVICE pointers unless
 * the name ends in _host; the caller owns every buffer; the library allocates nothing persistent except
 * a small per-device status word.  Every call is asynchronous on `stream` (a cudaStream_t passed as
 * void*), returns 0 on success or a negative mmd_status, and never falls back to the CPU: on a device
 * that is not sm_100 the compute entry points return MMD_ERR_DEVICE.
 *
 * Contract of the path: (query embeddings [Q,D], corpus embeddings [N,D], K) -> (scores f32 [Q,K] sorted
 * descending, indices i32 [Q,K] = 0-based corpus rows).  Ties are ordered by ascending corpus row.
 * Slots beyond min(K,N) hold score = -inf and index = -1.
 */
#ifndef MMD_RETRIEVAL_H_
#define MMD_RETRIEVAL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMD_ABI_VERSION 1

#if defined(__GNUC__)
#define MMD_API __attribute__((visibility("default")))
#else
#define MMD_API
#endif

typedef enum mmd_status {
  MMD_OK = 0,
  MMD_ERR_ARG = -1,       /* bad argument (null pointer, non-positive size, unsupported K, misalignment) */
  MMD_ERR_DEVICE = -2,    /* no CUDA device, or device is not compute capability 10.x */
  MMD_ERR_CUDA = -3,      /* a CUDA runtime / driver call failed; see mmd_last_error() */
  MMD_ERR_WORKSPACE = -4, /* workspace too small; call mmd_topk_workspace_bytes() */
  MMD_ERR_KERNEL = -5     /* a kernel reported a pipeline fault through the device status word */
} mmd_status;

/* dtype of caller-side embeddings (what the reference stores: fp32 image features
 * im2im_retrieval.py:29-36, fp16 text embeddings text2text_retrieval.py:44,53). */
typedef enum mmd_src_dtype { MMD_SRC_F32 = 0, MMD_SRC_F16 = 1, MMD_SRC_BF16 = 2 } mmd_src_dtype;

/* operand format of the prepared (normalised, cast) rows the tensor-core contraction reads. */
typedef enum mmd_op_dtype {
  MMD_OP_BF16 = 0,   /* bf16 operands, fp32 accumulate (tcgen05 kind::f16)                        */
  MMD_OP_F16 = 1,    /* fp16 operands, fp32 accumulate (tcgen05 kind::f16)                        */
  MMD_OP_E4M3 = 2,   /* fp8 e4m3 operands scaled by 2^8, fp32 accumulate (tcgen05 kind::f8f6f4)   */
  MMD_OP_BF16X3 = 3  /* fp32-accurate: every value split into 3 bf16 limbs, 6 cross terms laid
                        out along K, fp32 accumulate (the "fp32" configuration of the path)      */
} mmd_op_dtype;

typedef enum mmd_side { MMD_SIDE_QUERY = 0, MMD_SIDE_CORPUS = 1 } mmd_side;

/* ---- introspection ------------------------------------------------------------------------- */
MMD_API int mmd_abi_version(void);
/* Human-readable text of the last failure on the calling thread ("" if none). */
MMD_API const char* mmd_last_error(void);
/* 0 if the current CUDA device can run the path (compute capability 10.x), else MMD_ERR_DEVICE. */
MMD_API int mmd_device_check(void);

/* ---- K1: fused normalise-and-cast ---------------------------------------------------------- */
/* Layout of prepared rows for (op_dtype, dim): *kdim = length of the contraction axis in operand
 * elements (dim padded to 16 bytes; x6 for BF16X3), *row_bytes = pitch of one prepared row. */
MMD_API int mmd_prepared_layout(int op_dtype, int dim, int64_t* kdim, int64_t* row_bytes);

/* dst[r,:] = cast(src[r,:] * (normalize ? 1/max(||src[r,:]||_2, eps) : 1))  for r in [0,rows)
 * (F.normalize semantics, sentence_transformers.util.cos_sim; nn.CosineSimilarity per-norm clamp,
 * im2im_retrieval.py:40).  inv_norm (nullable) receives the fp32 row multipliers.  src rows are
 * src_row_stride elements apart.  side selects the limb order for MMD_OP_BF16X3. */
MMD_API int mmd_normalize_cast(const void* src, int src_dtype, int64_t rows, int dim, int64_t src_row_stride,
                       int normalize, float eps, int op_dtype, int side, void* dst, float* inv_norm,
                       void* stream);

/* One SEGMENT of a joint (multi-modality) prepared row: like mmd_normalize_cast, but every value is also multiplied by
 * `scale` and the rows are written with pitch dst_row_bytes starting at `dst` (= row 0 of the joint buffer + the
 * segment's byte offset, a multiple of 16).  Laying the modalities of a claim / an evidence side by side along K, with
 * the query side scaled by the modality weight w_m, makes ONE contraction compute sum_m w_m * cos(q_m, c_m) -- the joint
 * image+text score (BASELINE.json configs[3]; the reference's closest counterpart is the concat + sort of the two
 * hit lists, src/evidence/text2text_retrieval.py:97-118).  Not available for MMD_OP_BF16X3. */
MMD_API int mmd_normalize_cast_segment(const void* src, int src_dtype, int64_t rows, int dim, int64_t src_row_stride,
                               int normalize, float eps, float scale, int op_dtype, int side, void* dst,
                               int64_t dst_row_bytes, float* inv_norm, void* stream);

/* ---- K2+K3: tensor-core score contraction with fused top-K ---------------------------------- */
/* Largest K the fused selection supports. */
MMD_API int mmd_topk_max_k(void);
/* Bytes of scratch mmd_topk_scores needs for this problem: one K-list of 64-bit keys per (query, strip of the planned
 * schedule), the per-query pruning bounds (uint32 [Q]) and the quantile slots of the merged bounds (uint32 [30][Q]).  Device
 * memory, 8-byte aligned, private to one call in flight; the library zeroes what it reads. */
MMD_API size_t mmd_topk_workspace_bytes(int64_t Q, int64_t N, int dim, int op_dtype, int k);

/* q_prep [Q rows], c_prep [N rows]: prepared by mmd_normalize_cast with the same op_dtype / dim.
 * out_scores f32 [Q,k] descending, out_idx i32 [Q,k] = corpus row + idx_offset.  The Q x N score
 * matrix is never written to memory. */
MMD_API int mmd_topk_scores(const void* q_prep, const void* c_prep, int op_dtype, int64_t Q, int64_t N, int dim, int k,
                    int64_t idx_offset, float* out_scores, int32_t* out_idx, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Row-sharded corpora: same call, but the per-query pruning thresholds live in caller-owned arrays shared by ALL ranks.
 * thr_local = this rank's array (uint32 [Q], zeroed by the caller before the step on every rank, with a barrier between
 * the zeroing and the first launch); thr_all_host = HOST array of n_thr (1..8) DEVICE pointers to every rank's array (own one
 * included; peer-mapped memory).  A shard's K-th best score is a lower bound of the global K-th best, so every bound a
 * CTA learns is published to all ranks with system-scope atomics over NVLink and prunes every shard.  The lists returned
 * then hold each shard's candidates for the GLOBAL top-k (possibly fewer than k valid entries), not its complete local
 * top-k; the merged global top-k is unchanged.  pair_dst_host (n_pair_dst = 0..16 device pointers, may be NULL/0):
 * the strip merge additionally stores the list as packed {score bits, row} pairs at pair offset pair_offset + q * k + rank
 * into every destination -- every rank's gather buffer -- so the candidate exchange needs no kernel of its own. */
MMD_API int mmd_topk_scores_shared(const void* q_prep, const void* c_prep, int op_dtype, int64_t Q, int64_t N, int dim, int k,
                           int64_t idx_offset, float* out_scores, int32_t* out_idx, void* workspace,
                           size_t workspace_bytes, uint32_t* thr_local, void* const* thr_all_host, int n_thr,
                           void* const* pair_dst_host, int n_pair_dst, int64_t pair_offset, void* stream);

/* Same contraction, dense output scores f32 [Q, ld_scores] (small shapes: pairwise similarity,
 * tests, custom post-processing). */
MMD_API int mmd_scores_dense(const void* q_prep, const void* c_prep, int op_dtype, int64_t Q, int64_t N, int dim,
                     float* out_scores, int64_t ld_scores, void* stream);

/* ---- K3b/K4: merge of partial top-K lists ---------------------------------------------------- */
/* scores/idx: [parts, Q, k_in] (e.g. the all-gathered per-rank lists of a row-sharded corpus);
 * out: [Q, k_out] best by (score desc, idx asc).  parts*k_in <= 4096. */
MMD_API int mmd_topk_merge(const float* scores, const int32_t* idx, int parts, int64_t Q, int k_in, int k_out,
                   float* out_scores, int32_t* out_idx, void* stream);

/* Same merge over packed lists: pairs [parts][Q, k_in] of {int32 score bits (IEEE f32), int32 row} -- the layout
 * the exchange step of a row-sharded corpus moves (mmd_rescore_pairs / mmd_scatter_pairs write it).  Part p starts
 * part_stride_pairs * p pairs after `pairs` (0 = Q * k_in, i.e. densely packed).  parts_sorted != 0 promises that every
 * list is already in (score descending, row ascending) order with its empty slots last -- what mmd_rescore_pairs and
 * mmd_topk_scores produce -- and skips the per-list sort; 0 accepts lists in any order. */
MMD_API int mmd_topk_merge_pairs(const void* pairs, int parts, int64_t part_stride_pairs, int64_t Q, int k_in, int k_out,
                         int parts_sorted, float* out_scores, int32_t* out_idx, void* stream);

/* Pack ranked lists (scores f32 [Q,k], idx i32 [Q,k]) into {score bits, row} pairs and store them to n_dst (1..16)
 * destination buffers at pair offset dst_offset_pairs (dst_host: HOST array of DEVICE pointers, e.g. every peer's
 * gather buffer): the candidate exchange of the row-sharded path when the exact re-score runs AFTER the global merge. */
MMD_API int mmd_scatter_pairs(const float* scores, const int32_t* idx, int64_t Q, int k, void* const* dst_host, int n_dst,
                      int64_t dst_offset_pairs, void* stream);

/* ---- K5: exact re-score of the selected candidates ------------------------------------------ */
/* For each query q and candidate j < k_in with cand_idx[q,j] >= 0 :
 *   s = (sum_i q_src[q,i] * c_src[cand_idx[q,j] - idx_offset, i]) * q_inv[q] * c_inv[row]   (fp32)
 * then the k_out best by (s desc, idx asc) are written.  q_inv / c_inv may be NULL (= 1). */
MMD_API int mmd_rescore(const void* q_src, int q_dtype, int64_t q_stride, const float* q_inv, const void* c_src,
                int c_dtype, int64_t c_stride, const float* c_inv, int64_t Q, int64_t N, int dim,
                const int32_t* cand_idx, int k_in, int64_t idx_offset, int k_out, float* out_scores,
                int32_t* out_idx, void* stream);

/* Same re-score, but the k_out best are written as packed {score bits, row} pairs to n_dst (1..16) destination
 * buffers at pair offset dst_offset_pairs + q * k_out + rank.  dst_host is a HOST array of n_dst DEVICE pointers:
 * one local send buffer for an NCCL all-gather, or every peer's gather buffer ([world][Q][k_out] pairs, offset =
 * rank * Q * k_out) when the buffers are peer-mapped (NVLink): then the exchange step of the row-sharded path
 * happens inside this kernel's stores and only a barrier remains.  No reference counterpart (single process). */
MMD_API int mmd_rescore_pairs(const void* q_src, int q_dtype, int64_t q_stride, const float* q_inv, const void* c_src,
                      int c_dtype, int64_t c_stride, const float* c_inv, int64_t Q, int64_t N, int dim,
                      const int32_t* cand_idx, int k_in, int64_t idx_offset, int k_out, void* const* dst_host,
                      int n_dst, int64_t dst_offset_pairs, void* stream);

/* Joint re-score over n_seg (1..4) modalities: s = sum_m weight[m] * <q_m, c_m> * q_inv[m][q] * c_inv[m][row] in fp32.
 * Every *_host argument is a HOST array of n_seg entries (device pointers / dtypes / strides / dims / weights);
 * q_inv_host / c_inv_host (or single entries) may be NULL (= 1). */
MMD_API int mmd_rescore_joint(int n_seg, const void* const* q_src_host, const int* q_dtype_host, const int64_t* q_stride_host,
                      const float* const* q_inv_host, const void* const* c_src_host, const int* c_dtype_host,
                      const int64_t* c_stride_host, const float* const* c_inv_host, const int* dim_host,
                      const float* weight_host, int64_t Q, int64_t N, const int32_t* cand_idx, int k_in,
                      int64_t idx_offset, int k_out, float* out_scores, int32_t* out_idx, void* stream);

/* mmd_rescore_joint with the packed multi-destination output of mmd_rescore_pairs. */
MMD_API int mmd_rescore_joint_pairs(int n_seg, const void* const* q_src_host, const int* q_dtype_host,
                            const int64_t* q_stride_host, const float* const* q_inv_host, const void* const* c_src_host,
                            const int* c_dtype_host, const int64_t* c_stride_host, const float* const* c_inv_host,
                            const int* dim_host, const float* weight_host, int64_t Q, int64_t N, const int32_t* cand_idx,
                            int k_in, int64_t idx_offset, int k_out, void* const* dst_host, int n_dst,
                            int64_t dst_offset_pairs, void* stream);

/* ---- row-sharded corpus: one search step = three kernels per rank, synchronised through flags in peer memory -------
 * One process per GPU; the corpus rows are split over `world` ranks, the queries are replicated.  No reference
 * counterpart (single process): the step stands where sentence_transformers.util.semantic_search merges its corpus
 * chunks with a per-query heap (call sites src/evidence/text2text_retrieval.py:56-64, experiment_text.py:25-33).
 *
 * Synchronisation contract shared by the three calls: a STAGE owns a device word pair sync_state = {launches completed,
 * blocks finished} (zero-initialised by the caller, then only touched by the library).  A producing stage writes
 * (launches completed + 1) into arrive_flags_host[i] (one DEVICE pointer per rank: this rank's word in that rank's flag
 * array, peer-mapped) when its last block is done; a consuming stage spins at its start until wait_flags[0..n_wait)
 * (LOCAL flag array, one word per rank) have all reached (its own launches completed + 1).  Every rank must launch the
 * same sequence of stages.  A wait that lasts more than 4 s traps the kernel (CUDA error at the next synchronisation).
 *
 * Stage C: mmd_topk_scores_shared, plus: the strip merge stores this rank's candidate list as packed pairs at pair
 * pair_offset + q * pair_width + i (i < pair_width; entries beyond the k the shard can fill are empties) into every
 * destination and arrives on arrive_flags_host.  n_thr = 0 keeps the pruning thresholds private to the launch; with
 * n_thr > 0 and reset_thr != 0 thr_local[0..Q) is zeroed on `stream` first (bounds a faster peer has already published
 * for this step are lost, which only weakens the pruning). */
MMD_API int mmd_sharded_candidates(const void* q_prep, const void* c_prep, int op_dtype, int64_t Q, int64_t N, int dim, int k,
                           int64_t idx_offset, float* out_scores, int32_t* out_idx, void* workspace,
                           size_t workspace_bytes, uint32_t* thr_local, void* const* thr_all_host, int n_thr, int reset_thr,
                           void* const* pair_dst_host, int n_pair_dst, int64_t pair_offset, int pair_width,
                           void* const* arrive_flags_host, int n_arrive, uint32_t* sync_state, void* stream,
                           void* merge_stream);
/* (merge_stream: NULL, or a second stream on which the strip merge -- and with it the stores to the peers and the flag --
 * runs behind an event recorded after the contraction, so that the next contraction on `stream` need not wait for it.)
 * Zero n 32-bit words on `stream` (the shared threshold array of a step, reset before anyone publishes into it). */
MMD_API int mmd_zero_u32(uint32_t* dst, int64_t n, void* stream);

/* Stage X: wait; merge the `parts` sorted candidate lists of every query (gathered: part p starts p * part_stride_pairs
 * pairs in, [Q][k_in] pairs each) into the global candidate list of kc entries (identical on every rank); re-score in
 * fp32 -- modalities / source tables exactly as mmd_rescore_joint, n_seg = 1 with weight 1 is the plain cosine -- the
 * candidates whose global row lies in [idx_offset, idx_offset + N), and store each {exact score bits, row} at pair
 * dst_offset_pairs + q * kc + (position in the global list) into every destination (every rank's re-score buffer);
 * empty positions are written to dst_host[own_dst] only.  Arrives on arrive_flags_host. */
MMD_API int mmd_exchange_rescore(const void* gathered, int parts, int64_t part_stride_pairs, int64_t Q, int k_in, int kc,
                         int n_seg, const void* const* q_src_host, const int* q_dtype_host, const int64_t* q_stride_host,
                         const float* const* q_inv_host, const void* const* c_src_host, const int* c_dtype_host,
                         const int64_t* c_stride_host, const float* const* c_inv_host, const int* dim_host,
                         const float* weight_host, int64_t N, int64_t idx_offset, void* const* dst_host, int n_dst,
                         int own_dst, int64_t dst_offset_pairs, const uint32_t* wait_flags, int n_wait,
                         void* const* arrive_flags_host, int n_arrive, uint32_t* sync_state, void* stream);

/* Stage F: wait; sort the kc re-scored candidates of every query (rescored: [Q][kc] pairs) and write the k_out best:
 * out_scores f32 [Q,k_out], out_idx i32 or (idx_is_i64 != 0) i64 [Q,k_out]; missing entries are (-inf, -1). */
MMD_API int mmd_exchange_finish(const void* rescored, int64_t Q, int kc, int k_out, float* out_scores, void* out_idx,
                        int idx_is_i64, const uint32_t* wait_flags, int n_wait, uint32_t* sync_state, void* stream);

/* ---- distinct-score filter of ranked lists ---------------------------------------------------- */
/* scores/idx [Q, k_in] descending (as every entry point above returns them).  Keeps, per query, the first entry of every
 * distinct score -- and any entry whose row equals gold_idx[q] (gold_idx nullable; -1 = none) -- until top_k are kept:
 * the walk of src/evidence/im2im_retrieval.py:94-104 / text2text_retrieval.py:105-118 and its gold-exempting variants
 * experiment_image.py:41-50 / experiment_text.py:79-87.  out_* [Q, top_k] padded with (-inf, -1); out_count[q]
 * (nullable) = entries kept, so the caller can over-fetch more when a duplicate-heavy corpus leaves a list short. */
MMD_API int mmd_dedupe_scores(const float* scores, const int32_t* idx, const int32_t* gold_idx, int64_t Q, int k_in, int top_k,
                      float* out_scores, int32_t* out_idx, int32_t* out_count, void* stream);

/* ---- measurement hooks ------------------------------------------------------------------------ */
/* While enabled, every fused contraction launch of mmd_topk_scores is bracketed by CUDA events on its own
 * stream (up to 512 launches are kept).  mmd_profile_collect synchronises those events, writes up to `cap`
 * durations in milliseconds (oldest first), clears the list and returns how many it wrote.  A launch recorded while
 * the stream is being captured into a CUDA graph gets external event-record nodes instead: each later collect also
 * reports the fused-kernel duration of that graph's most recent replay. */
MMD_API int mmd_profile_enable(int on);
MMD_API int mmd_profile_collect(float* ms_host, int cap);

/* Launch counter: number of kernels this library has launched in this process (bench evidence). */
MMD_API int64_t mmd_launch_count(void);

/* Provenance of this binary: "src=<sha256 of csrc/*, include/*, build flags> nvcc=<version> arch=sm_100a ..." */
MMD_API const char* mmd_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* MMD_RETRIEVAL_H_ */
