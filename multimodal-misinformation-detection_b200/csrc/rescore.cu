// K5: exact fp32 re-score of the few candidates the low-precision tensor-core pass selected.
//
// The bf16 / fp8 contraction decides WHICH corpus rows are candidates (over-fetched to k_in > k_out); this
// kernel recomputes their cosine from the caller's original embeddings in fp32 -- the arithmetic of the
// reference (F.normalize + mm in sentence_transformers.util.cos_sim; nn.CosineSimilarity in
// src/evidence/im2im_retrieval.py:38-42) -- and re-ranks.  Returned scores therefore agree with the fp32
// reference to rounding, and the order inside the returned list is the fp32 order.
//
// Up to 32 candidates: one warp per query (rescore_warp_kernel).  Longer lists: one 128-thread block per query, warps
// stride over candidates, each computing one D-long dot product with 128-bit gathered loads (HBM/L2-bound gather:
// Q * k_in * D * sizeof(src) bytes), then a rank-by-counting pass orders the <= 1024 candidates.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "rescore_common.cuh"

namespace mmd {
namespace {

constexpr int kMaxCand = 1024;

// Where a re-scored list goes: separate score / index arrays (the plain contract) and/or packed
// {score bits, global row} pairs written to up to kMaxPairDst buffers -- the local send buffer of the all-gather,
// or every peer's gather buffer directly (peer-mapped pointers, stores travel over NVLink).
constexpr int kMaxPairDst = 16;
struct PairDst {
  int2* dst[kMaxPairDst];
  int n;
  int64_t offset;   // in pairs: rank * Q * k_out for a gather buffer laid out [world][Q][k_out]
};

__device__ __forceinline__ void emit(float* out_s, int32_t* out_i, const PairDst& pairs, int64_t pos, float sc, int32_t ix) {
  if (out_s != nullptr) {
    out_s[pos] = sc;
    out_i[pos] = ix;
  }
  const int2 v = make_int2(__float_as_int(sc), ix);
  for (int d = 0; d < pairs.n; ++d) pairs.dst[d][pairs.offset + pos] = v;
}

template <typename TQ, typename TC>
__global__ void __launch_bounds__(128)
rescore_kernel(const TQ* __restrict__ q_src, int64_t q_stride, const float* __restrict__ q_inv,
               const TC* __restrict__ c_src, int64_t c_stride, const float* __restrict__ c_inv, int64_t N, int dim,
               const int32_t* __restrict__ cand_idx, int k_in, int64_t idx_offset, int k_out,
               float* __restrict__ out_s, int32_t* __restrict__ out_i, PairDst pairs, bool vec_ok) {
  __shared__ uint64_t keys[kMaxCand];
  const int64_t q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TQ* qrow = q_src + q * q_stride;
  const float qi = q_inv != nullptr ? q_inv[q] : 1.0f;
  for (int j = warp; j < k_in; j += 4) {
    const int32_t gi = cand_idx[q * k_in + j];
    uint64_t key = 0ull;
    const int64_t row = static_cast<int64_t>(gi) - idx_offset;
    if (gi >= 0 && row >= 0 && row < N) {
      const float d = warp_dot<TQ, TC>(qrow, c_src + row * c_stride, dim, lane, vec_ok);
      const float ci = c_inv != nullptr ? c_inv[row] : 1.0f;
      key = make_key(d * qi * ci, static_cast<uint32_t>(gi));
    }
    if (lane == 0) keys[j] = key;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < k_in; j += 128) {
    const uint64_t mine = keys[j];
    int rank = 0;
    for (int t = 0; t < k_in; ++t) {
      const uint64_t o = keys[t];
      rank += (o > mine) || (o == mine && t < j);
    }
    if (rank < k_out) {
      const float sc = mine == 0ull ? __int_as_float(0xff800000) : key_score(mine);
      const int32_t ix = mine == 0ull ? -1 : static_cast<int32_t>(key_row(mine));
      emit(out_s, out_i, pairs, q * k_out + rank, sc, ix);
    }
  }
  for (int i = k_in + threadIdx.x; i < k_out; i += 128) emit(out_s, out_i, pairs, q * k_out + i, __int_as_float(0xff800000), -1);
}

// Lists of at most 32 candidates (every default configuration: K' = K + 8): ONE WARP per query, four queries per block,
// no shared memory and no block barrier.  The 32 candidate rows arrive with one coalesced load; the rows that live in this
// shard (all of them for an unsharded corpus, K'/world on average after a global candidate merge) are re-scored one after
// the other, each lane keeps the key of "its" candidate, and the ranks come from a shuffle sweep.
template <typename TQ, typename TC>
__global__ void __launch_bounds__(128)
rescore_warp_kernel(const TQ* __restrict__ q_src, int64_t q_stride, const float* __restrict__ q_inv,
                    const TC* __restrict__ c_src, int64_t c_stride, const float* __restrict__ c_inv, int64_t Q, int64_t N,
                    int dim, const int32_t* __restrict__ cand_idx, int k_in, int64_t idx_offset, int k_out,
                    float* __restrict__ out_s, int32_t* __restrict__ out_i, PairDst pairs, bool vec_ok) {
  const int lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (q >= Q) return;
  const TQ* qrow = q_src + q * q_stride;
  const float qi = q_inv != nullptr ? q_inv[q] : 1.0f;
  const int32_t gi = lane < k_in ? cand_idx[q * k_in + lane] : -1;
  const int64_t row = static_cast<int64_t>(gi) - idx_offset;
  uint32_t m = __ballot_sync(0xffffffffu, gi >= 0 && row >= 0 && row < N);
  uint64_t key = 0ull;
  while (m) {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    const int32_t gb = __shfl_sync(0xffffffffu, gi, b);
    const int64_t rb = static_cast<int64_t>(gb) - idx_offset;
    const float d = warp_dot<TQ, TC>(qrow, c_src + rb * c_stride, dim, lane, vec_ok);
    const float ci = c_inv != nullptr ? c_inv[rb] : 1.0f;
    if (lane == b) key = make_key(d * qi * ci, static_cast<uint32_t>(gb));
  }
  int rank = 0;
  for (int t = 0; t < k_in; ++t) {
    const uint64_t o = __shfl_sync(0xffffffffu, key, t);
    rank += (o > key) || (o == key && t < lane);
  }
  if (lane < k_in && rank < k_out) {
    const float sc = key == 0ull ? __int_as_float(0xff800000) : key_score(key);
    const int32_t ix = key == 0ull ? -1 : static_cast<int32_t>(key_row(key));
    emit(out_s, out_i, pairs, q * k_out + rank, sc, ix);
  }
  for (int i = k_in + lane; i < k_out; i += 32) emit(out_s, out_i, pairs, q * k_out + i, __int_as_float(0xff800000), -1);
}

template <typename TQ, typename TC>
int launch(const void* q_src, int64_t q_stride, const float* q_inv, const void* c_src, int64_t c_stride,
           const float* c_inv, int64_t Q, int64_t N, int dim, const int32_t* cand_idx, int k_in, int64_t idx_offset,
           int k_out, float* out_s, int32_t* out_i, const PairDst& pairs, cudaStream_t stream) {
  const bool vec_ok = dim % 8 == 0 && reinterpret_cast<uintptr_t>(q_src) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(c_src) % 16 == 0 && (q_stride * sizeof(TQ)) % 16 == 0 &&
                      (c_stride * sizeof(TC)) % 16 == 0;
  if (k_in <= 32) {
    rescore_warp_kernel<TQ, TC><<<static_cast<unsigned>((Q + 3) / 4), 128, 0, stream>>>(
        static_cast<const TQ*>(q_src), q_stride, q_inv, static_cast<const TC*>(c_src), c_stride, c_inv, Q, N, dim, cand_idx,
        k_in, idx_offset, k_out, out_s, out_i, pairs, vec_ok);
  } else {
    rescore_kernel<TQ, TC><<<static_cast<unsigned>(Q), 128, 0, stream>>>(
        static_cast<const TQ*>(q_src), q_stride, q_inv, static_cast<const TC*>(c_src), c_stride, c_inv, N, dim, cand_idx,
        k_in, idx_offset, k_out, out_s, out_i, pairs, vec_ok);
  }
  count_launch();
  MMD_CUDA_OK(cudaGetLastError());
  return MMD_OK;
}

template <typename TQ>
int dispatch_c(int c_dtype, const void* q_src, int64_t q_stride, const float* q_inv, const void* c_src, int64_t c_stride,
               const float* c_inv, int64_t Q, int64_t N, int dim, const int32_t* cand_idx, int k_in, int64_t idx_offset,
               int k_out, float* out_s, int32_t* out_i, const PairDst& pairs, cudaStream_t stream) {
  switch (c_dtype) {
    case MMD_SRC_F32: return launch<TQ, float>(q_src, q_stride, q_inv, c_src, c_stride, c_inv, Q, N, dim, cand_idx, k_in, idx_offset, k_out, out_s, out_i, pairs, stream);
    case MMD_SRC_F16: return launch<TQ, __half>(q_src, q_stride, q_inv, c_src, c_stride, c_inv, Q, N, dim, cand_idx, k_in, idx_offset, k_out, out_s, out_i, pairs, stream);
    case MMD_SRC_BF16: return launch<TQ, __nv_bfloat16>(q_src, q_stride, q_inv, c_src, c_stride, c_inv, Q, N, dim, cand_idx, k_in, idx_offset, k_out, out_s, out_i, pairs, stream);
  }
  set_last_error("mmd_rescore: unknown c_dtype %d", c_dtype);
  return MMD_ERR_ARG;
}

}  // namespace
}  // namespace mmd

namespace mmd {
namespace {

// ---------------------------------------------------------------- joint (multi-modality) re-score
// score(q, c) = sum_m weight_m * <q_m, c_m> * q_inv_m[q] * c_inv_m[c]  over up to kMaxSeg modalities, each with its own
// embeddings (own dim / dtype / stride).  Element types are switched at run time (uniform per segment).
__global__ void __launch_bounds__(128)
rescore_multi_kernel(Segments sg, int64_t N, const int32_t* __restrict__ cand_idx, int k_in, int64_t idx_offset, int k_out,
                     float* __restrict__ out_s, int32_t* __restrict__ out_i, PairDst pairs) {
  __shared__ uint64_t keys[kMaxCand];
  const int64_t q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < k_in; j += 4) {
    const int32_t gi = cand_idx[q * k_in + j];
    uint64_t key = 0ull;
    const int64_t row = static_cast<int64_t>(gi) - idx_offset;
    if (gi >= 0 && row >= 0 && row < N) {
      const float total = segments_score(sg, q, row, lane);
      key = make_key(total, static_cast<uint32_t>(gi));
    }
    if (lane == 0) keys[j] = key;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < k_in; j += 128) {
    const uint64_t mine = keys[j];
    int rank = 0;
    for (int t = 0; t < k_in; ++t) {
      const uint64_t o = keys[t];
      rank += (o > mine) || (o == mine && t < j);
    }
    if (rank < k_out) {
      const float sc = mine == 0ull ? __int_as_float(0xff800000) : key_score(mine);
      const int32_t ix = mine == 0ull ? -1 : static_cast<int32_t>(key_row(mine));
      emit(out_s, out_i, pairs, q * k_out + rank, sc, ix);
    }
  }
  for (int i = k_in + threadIdx.x; i < k_out; i += 128) emit(out_s, out_i, pairs, q * k_out + i, __int_as_float(0xff800000), -1);
}

}  // namespace
}  // namespace mmd

namespace mmd {
namespace {
int rescore_joint_any(int n_seg, const void* const* q_src_host, const int* q_dtype_host, const int64_t* q_stride_host,
                      const float* const* q_inv_host, const void* const* c_src_host, const int* c_dtype_host,
                      const int64_t* c_stride_host, const float* const* c_inv_host, const int* dim_host,
                      const float* weight_host, int64_t Q, int64_t N, const int32_t* cand_idx, int k_in,
                      int64_t idx_offset, int k_out, float* out_scores, int32_t* out_idx, const PairDst& pairs, void* stream,
                      const char* who) {
  MMD_REQUIRE(Q >= 0 && N >= 0 && k_in > 0 && k_out > 0 && k_in <= kMaxCand, "%s: Q=%lld N=%lld k_in=%d k_out=%d", who,
              (long long)Q, (long long)N, k_in, k_out);
  if (Q == 0) return MMD_OK;
  MMD_REQUIRE(cand_idx != nullptr, "%s: null candidate list", who);
  MMD_REQUIRE((out_scores != nullptr && out_idx != nullptr) || (out_scores == nullptr && out_idx == nullptr && pairs.n > 0),
              "%s: no output buffer", who);
  Segments sg{};
  int rc = fill_segments(&sg, n_seg, q_src_host, q_dtype_host, q_stride_host, q_inv_host, c_src_host, c_dtype_host, c_stride_host,
                         c_inv_host, dim_host, weight_host, N, who);
  if (rc != MMD_OK) return rc;
  rc = mmd_device_check();
  if (rc != MMD_OK) return rc;
  rescore_multi_kernel<<<static_cast<unsigned>(Q), 128, 0, static_cast<cudaStream_t>(stream)>>>(sg, N, cand_idx, k_in, idx_offset,
                                                                                                 k_out, out_scores, out_idx, pairs);
  count_launch();
  MMD_CUDA_OK(cudaGetLastError());
  return MMD_OK;
}
}  // namespace
}  // namespace mmd

extern "C" int mmd_rescore_joint(int n_seg, const void* const* q_src_host, const int* q_dtype_host, const int64_t* q_stride_host,
                                 const float* const* q_inv_host, const void* const* c_src_host, const int* c_dtype_host,
                                 const int64_t* c_stride_host, const float* const* c_inv_host, const int* dim_host,
                                 const float* weight_host, int64_t Q, int64_t N, const int32_t* cand_idx, int k_in,
                                 int64_t idx_offset, int k_out, float* out_scores, int32_t* out_idx, void* stream) {
  mmd::PairDst none{};
  MMD_REQUIRE(Q == 0 || (out_scores != nullptr && out_idx != nullptr), "mmd_rescore_joint: null output");
  return mmd::rescore_joint_any(n_seg, q_src_host, q_dtype_host, q_stride_host, q_inv_host, c_src_host, c_dtype_host, c_stride_host,
                                c_inv_host, dim_host, weight_host, Q, N, cand_idx, k_in, idx_offset, k_out, out_scores, out_idx,
                                none, stream, "mmd_rescore_joint");
}

extern "C" int mmd_rescore_joint_pairs(int n_seg, const void* const* q_src_host, const int* q_dtype_host,
                                       const int64_t* q_stride_host, const float* const* q_inv_host,
                                       const void* const* c_src_host, const int* c_dtype_host, const int64_t* c_stride_host,
                                       const float* const* c_inv_host, const int* dim_host, const float* weight_host,
                                       int64_t Q, int64_t N, const int32_t* cand_idx, int k_in, int64_t idx_offset, int k_out,
                                       void* const* dst_host, int n_dst, int64_t dst_offset_pairs, void* stream) {
  using namespace mmd;
  MMD_REQUIRE(dst_host != nullptr && n_dst >= 1 && n_dst <= kMaxPairDst, "mmd_rescore_joint_pairs: n_dst=%d (1..%d)", n_dst,
              kMaxPairDst);
  MMD_REQUIRE(dst_offset_pairs >= 0, "mmd_rescore_joint_pairs: negative offset");
  PairDst pairs{};
  pairs.n = n_dst;
  pairs.offset = dst_offset_pairs;
  for (int d = 0; d < n_dst; ++d) {
    MMD_REQUIRE(dst_host[d] != nullptr && reinterpret_cast<uintptr_t>(dst_host[d]) % 8 == 0,
                "mmd_rescore_joint_pairs: destination %d is null or not 8-byte aligned", d);
    pairs.dst[d] = static_cast<int2*>(dst_host[d]);
  }
  return rescore_joint_any(n_seg, q_src_host, q_dtype_host, q_stride_host, q_inv_host, c_src_host, c_dtype_host, c_stride_host,
                           c_inv_host, dim_host, weight_host, Q, N, cand_idx, k_in, idx_offset, k_out, nullptr, nullptr, pairs,
                           stream, "mmd_rescore_joint_pairs");
}

namespace mmd {
namespace {
int rescore_any(const void* q_src, int q_dtype, int64_t q_stride, const float* q_inv, const void* c_src, int c_dtype,
                int64_t c_stride, const float* c_inv, int64_t Q, int64_t N, int dim, const int32_t* cand_idx, int k_in,
                int64_t idx_offset, int k_out, float* out_scores, int32_t* out_idx, const PairDst& pairs, void* stream,
                const char* who) {
  MMD_REQUIRE(Q >= 0 && N >= 0 && dim > 0 && k_in > 0 && k_out > 0, "%s: Q=%lld N=%lld dim=%d k_in=%d k_out=%d", who,
              (long long)Q, (long long)N, dim, k_in, k_out);
  MMD_REQUIRE(k_in <= kMaxCand, "%s: k_in %d exceeds %d", who, k_in, kMaxCand);
  if (Q == 0) return MMD_OK;
  MMD_REQUIRE(q_src != nullptr && cand_idx != nullptr, "%s: null buffer", who);
  MMD_REQUIRE((out_scores != nullptr && out_idx != nullptr) || (out_scores == nullptr && out_idx == nullptr && pairs.n > 0),
              "%s: no output buffer", who);
  MMD_REQUIRE(c_src != nullptr || N == 0, "%s: null corpus", who);
  MMD_REQUIRE(q_stride >= dim && (c_stride >= dim || N == 0), "%s: row stride smaller than dim", who);
  int rc = mmd_device_check();
  if (rc != MMD_OK) return rc;
  auto st = static_cast<cudaStream_t>(stream);
  switch (q_dtype) {
    case MMD_SRC_F32: return dispatch_c<float>(c_dtype, q_src, q_stride, q_inv, c_src, c_stride, c_inv, Q, N, dim, cand_idx, k_in, idx_offset, k_out, out_scores, out_idx, pairs, st);
    case MMD_SRC_F16: return dispatch_c<__half>(c_dtype, q_src, q_stride, q_inv, c_src, c_stride, c_inv, Q, N, dim, cand_idx, k_in, idx_offset, k_out, out_scores, out_idx, pairs, st);
    case MMD_SRC_BF16: return dispatch_c<__nv_bfloat16>(c_dtype, q_src, q_stride, q_inv, c_src, c_stride, c_inv, Q, N, dim, cand_idx, k_in, idx_offset, k_out, out_scores, out_idx, pairs, st);
  }
  set_last_error("%s: unknown q_dtype %d", who, q_dtype);
  return MMD_ERR_ARG;
}
}  // namespace
}  // namespace mmd

extern "C" int mmd_rescore(const void* q_src, int q_dtype, int64_t q_stride, const float* q_inv, const void* c_src,
                           int c_dtype, int64_t c_stride, const float* c_inv, int64_t Q, int64_t N, int dim,
                           const int32_t* cand_idx, int k_in, int64_t idx_offset, int k_out, float* out_scores,
                           int32_t* out_idx, void* stream) {
  mmd::PairDst none{};
  MMD_REQUIRE(out_scores != nullptr && out_idx != nullptr, "mmd_rescore: null output");
  return mmd::rescore_any(q_src, q_dtype, q_stride, q_inv, c_src, c_dtype, c_stride, c_inv, Q, N, dim, cand_idx, k_in,
                          idx_offset, k_out, out_scores, out_idx, none, stream, "mmd_rescore");
}

extern "C" int mmd_rescore_pairs(const void* q_src, int q_dtype, int64_t q_stride, const float* q_inv, const void* c_src,
                                 int c_dtype, int64_t c_stride, const float* c_inv, int64_t Q, int64_t N, int dim,
                                 const int32_t* cand_idx, int k_in, int64_t idx_offset, int k_out, void* const* dst_host,
                                 int n_dst, int64_t dst_offset_pairs, void* stream) {
  using namespace mmd;
  MMD_REQUIRE(dst_host != nullptr && n_dst >= 1 && n_dst <= kMaxPairDst, "mmd_rescore_pairs: n_dst=%d (1..%d)", n_dst,
              kMaxPairDst);
  MMD_REQUIRE(dst_offset_pairs >= 0, "mmd_rescore_pairs: negative offset");
  PairDst pairs{};
  pairs.n = n_dst;
  pairs.offset = dst_offset_pairs;
  for (int d = 0; d < n_dst; ++d) {
    MMD_REQUIRE(dst_host[d] != nullptr && reinterpret_cast<uintptr_t>(dst_host[d]) % 8 == 0,
                "mmd_rescore_pairs: destination %d is null or not 8-byte aligned", d);
    pairs.dst[d] = static_cast<int2*>(dst_host[d]);
  }
  return rescore_any(q_src, q_dtype, q_stride, q_inv, c_src, c_dtype, c_stride, c_inv, Q, N, dim, cand_idx, k_in,
                     idx_offset, k_out, nullptr, nullptr, pairs, stream, "mmd_rescore_pairs");
}
