// K3b / K4: fold partial top-K lists into one list per query.
//
// Two producers feed it: the strips of the fused kernel (keys, [Q][parts][k_in]) and the per-rank lists of
// a row-sharded corpus after the all-gather ([parts][Q][k_in] decoded scores + rows).  The reference's
// counterpart is the per-query heapq push / pushpop over 500k-row corpus chunks followed by sorted(...)
// inside sentence_transformers.util.semantic_search (call sites src/evidence/text2text_retrieval.py:56-64),
// and the list concat + sort of src/evidence/text2text_retrieval.py:97-110.
//
// Lists of up to 128 entries (every list the fused kernel produces): one query per warp keeps a running sorted
// list in registers and folds one partial list at a time into it -- max against the reversed newcomer (the
// bitonic half-cleaner: the result holds the best 32*E of both) plus one log2-depth bitonic merge through
// shuffles.  O(parts * log L) per query instead of a full sort of parts * k_in keys.  Longer lists: one query
// per 256-thread block, bitonic sort in shared memory (<= 4096 candidates).
// Order: (score descending, row ascending).
#include "common.cuh"
#include "peer_sync.cuh"
#include "warp_sort.cuh"

namespace mmd {
namespace {

struct MergeSrc {
  PairOut po;             // optional packed copies of the result (see PairOut)
  PeerArrive arrive;      // optional: signal every rank once the whole grid has stored its lists (row-sharded path)
  const uint64_t* keys;   // [Q][parts * k_in]                    (keys != nullptr)
  const int2* pairs;      // [parts][Q][k_in] {score bits, row}    (keys == nullptr, pairs != nullptr)
  const float* scores;    // [parts][Q][k_in]                      (otherwise)
  const int32_t* idx;
  int parts, k_in;
  int64_t Q;
  int64_t part_stride;    // pairs / entries between consecutive parts of the [parts][Q][k_in] layouts (>= Q * k_in)
};

__device__ __forceinline__ uint64_t load_candidate(const MergeSrc& s, int64_t q, int i) {
  if (i >= s.parts * s.k_in) return 0ull;
  if (s.keys != nullptr) return s.keys[q * (static_cast<int64_t>(s.parts) * s.k_in) + i];
  const int part = i / s.k_in, j = i - part * s.k_in;
  const int64_t off = static_cast<int64_t>(part) * s.part_stride + q * s.k_in + j;
  if (s.pairs != nullptr) {
    const int2 v = s.pairs[off];
    return v.y < 0 ? 0ull : make_key(__int_as_float(v.x), static_cast<uint32_t>(v.y));
  }
  const int32_t r = s.idx[off];
  return r < 0 ? 0ull : make_key(s.scores[off], static_cast<uint32_t>(r));
}

// Entry i of query q's merged list: to out_s/out_i [Q, k_out] when i < k_out, and -- row-sharded corpora -- packed, straight
// into every rank's gather buffer at pair q * width + i for i < width (entries beyond k_out are empties there).
__device__ __forceinline__ void store_result(uint64_t key, float scale, int64_t idx_offset, float* out_s, int32_t* out_i,
                                             int64_t q, int i, int k_out, const PairOut& po) {
  if (i >= k_out) key = 0ull;
  const float sc = key == 0ull ? __int_as_float(0xff800000) : key_score(key) * scale;
  const int32_t ix = key == 0ull ? -1 : static_cast<int32_t>(static_cast<int64_t>(key_row(key)) + idx_offset);
  if (i < k_out) {
    out_s[q * k_out + i] = sc;
    out_i[q * k_out + i] = ix;
  }
  if (po.n > 0) {
    const int width = po.width > 0 ? po.width : k_out;
    if (i < width) {
      const int2 v = make_int2(__float_as_int(sc), ix);
      const int64_t pos = po.offset + q * width + i;
      for (int d = 0; d < po.n; ++d) po.dst[d][pos] = v;
    }
  }
}
__device__ __forceinline__ int store_extent(int k_out, const PairOut& po) {
  return (po.n > 0 && po.width > k_out) ? po.width : k_out;
}

// Streaming merge, one query per warp, lists of at most L = 32 * E entries.
//   kSortedParts = true : every partial list is already descending (the fused kernel's strips).
//   kSortedParts = false: partial lists in any order (caller-provided lists): each is sorted first.
template <int E, bool kSortedParts>
__global__ void __launch_bounds__(128)
merge_stream_kernel(MergeSrc src, int k_out, float scale, int64_t idx_offset, float* __restrict__ out_s,
                    int32_t* __restrict__ out_i) {
  constexpr int L = 32 * E;
  const int lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (q < src.Q) {
  uint64_t best[E];
#pragma unroll
  for (int e = 0; e < E; ++e) best[e] = 0ull;
  for (int part = 0; part < src.parts; ++part) {
    uint64_t nw[E];
    if constexpr (kSortedParts) {
      // position i of the ascending newcomer = entry L-1-i of the descending list (empty slots first)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int j = L - 1 - (e * 32 + lane);
        nw[e] = j < src.k_in ? load_candidate(src, q, part * src.k_in + j) : 0ull;
      }
    } else {
      uint64_t tmp[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int j = e * 32 + lane;
        tmp[e] = j < src.k_in ? load_candidate(src, q, part * src.k_in + j) : 0ull;
      }
      warp_bitonic_desc<E>(tmp, lane);
#pragma unroll
      for (int e = 0; e < E; ++e) nw[e] = __shfl_sync(kWarpFull, tmp[E - 1 - e], 31 - lane);
    }
#pragma unroll
    for (int e = 0; e < E; ++e) best[e] = best[e] > nw[e] ? best[e] : nw[e];
    warp_bitonic_merge_desc<E>(best, lane);
  }
#pragma unroll
  for (int e = 0; e < E; ++e) store_result(best[e], scale, idx_offset, out_s, out_i, q, e * 32 + lane, k_out, src.po);
  const int extent = store_extent(k_out, src.po);
  for (int i = L + lane; i < extent; i += 32) store_result(0ull, scale, idx_offset, out_s, out_i, q, i, k_out, src.po);
  }
  peer_arrive_all(src.arrive);
}

template <int L>
__global__ void __launch_bounds__(256)
merge_block_kernel(MergeSrc src, int k_out, float scale, int64_t idx_offset, float* __restrict__ out_s,
                   int32_t* __restrict__ out_i) {
  __shared__ uint64_t keys[L];
  const int64_t q = blockIdx.x;
  for (int i = threadIdx.x; i < L; i += 256) keys[i] = load_candidate(src, q, i);
  __syncthreads();
  for (int size = 2; size <= L; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < L / 2; t += 256) {
        const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));   // lower index of the pair
        const int j = i | stride;
        const bool desc = (i & size) == 0;
        const uint64_t a = keys[i], b = keys[j];
        if ((a < b) == desc) { keys[i] = b; keys[j] = a; }
      }
      __syncthreads();
    }
  }
  const int extent = store_extent(k_out, src.po);
  for (int i = threadIdx.x; i < extent; i += 256)
    store_result(i < L ? keys[i] : 0ull, scale, idx_offset, out_s, out_i, q, i, k_out, src.po);
  peer_arrive_all(src.arrive);
}

template <bool kSortedParts>
int launch_merge(const MergeSrc& src, int k_out, float scale, int64_t idx_offset, float* out_s, int32_t* out_i,
                 cudaStream_t stream) {
  const int total = src.parts * src.k_in;
  const unsigned qblocks = static_cast<unsigned>(ceil_div(src.Q, 4));
  const int longest = src.k_in > k_out ? src.k_in : k_out;
  if (longest <= 32) merge_stream_kernel<1, kSortedParts><<<qblocks, 128, 0, stream>>>(src, k_out, scale, idx_offset, out_s, out_i);
  else if (longest <= 64) merge_stream_kernel<2, kSortedParts><<<qblocks, 128, 0, stream>>>(src, k_out, scale, idx_offset, out_s, out_i);
  else if (longest <= 128) merge_stream_kernel<4, kSortedParts><<<qblocks, 128, 0, stream>>>(src, k_out, scale, idx_offset, out_s, out_i);
  else if (total <= 256) merge_block_kernel<256><<<static_cast<unsigned>(src.Q), 256, 0, stream>>>(src, k_out, scale, idx_offset, out_s, out_i);
  else if (total <= 512) merge_block_kernel<512><<<static_cast<unsigned>(src.Q), 256, 0, stream>>>(src, k_out, scale, idx_offset, out_s, out_i);
  else if (total <= 1024) merge_block_kernel<1024><<<static_cast<unsigned>(src.Q), 256, 0, stream>>>(src, k_out, scale, idx_offset, out_s, out_i);
  else if (total <= 2048) merge_block_kernel<2048><<<static_cast<unsigned>(src.Q), 256, 0, stream>>>(src, k_out, scale, idx_offset, out_s, out_i);
  else if (total <= 4096) merge_block_kernel<4096><<<static_cast<unsigned>(src.Q), 256, 0, stream>>>(src, k_out, scale, idx_offset, out_s, out_i);
  else {
    set_last_error("merge: %d candidates per query exceeds 4096 (lists longer than 128 entries)", total);
    return MMD_ERR_ARG;
  }
  count_launch();
  MMD_CUDA_OK(cudaGetLastError());
  return MMD_OK;
}

}  // namespace

int merge_partial_keys(const uint64_t* partial, int parts, int64_t Q, int k_in, int k_out, float scale,
                       int64_t idx_offset, float* out_scores, int32_t* out_idx, cudaStream_t stream, const PairOut* po,
                       const PeerArrive* arrive) {
  MergeSrc src{};
  if (po != nullptr) src.po = *po;
  if (arrive != nullptr) src.arrive = *arrive;
  src.keys = partial;
  src.pairs = nullptr;
  src.scores = nullptr;
  src.idx = nullptr;
  src.parts = partial == nullptr ? 0 : parts;
  src.k_in = k_in;
  src.Q = Q;
  src.part_stride = Q * k_in;
  if (partial == nullptr) {
    // no candidates at all: make load_candidate return "empty" for every slot
    src.keys = reinterpret_cast<const uint64_t*>(out_scores);   // never dereferenced (parts == 0)
  }
  return launch_merge<true>(src, k_out, scale, idx_offset, out_scores, out_idx, stream);
}

}  // namespace mmd

extern "C" int mmd_topk_merge(const float* scores, const int32_t* idx, int parts, int64_t Q, int k_in, int k_out,
                              float* out_scores, int32_t* out_idx, void* stream) {
  using namespace mmd;
  MMD_REQUIRE(parts > 0 && Q >= 0 && k_in > 0 && k_out > 0, "mmd_topk_merge: parts=%d Q=%lld k_in=%d k_out=%d", parts,
              (long long)Q, k_in, k_out);
  if (Q == 0) return MMD_OK;
  MMD_REQUIRE(scores != nullptr && idx != nullptr && out_scores != nullptr && out_idx != nullptr,
              "mmd_topk_merge: null buffer");
  MMD_REQUIRE((k_in <= 128 && k_out <= 128) || static_cast<int64_t>(parts) * k_in <= 4096,
              "mmd_topk_merge: parts*k_in = %lld exceeds 4096 (only lists of <= 128 entries stream)", (long long)parts * k_in);
  MMD_REQUIRE(k_out <= 4096, "mmd_topk_merge: k_out %d exceeds 4096", k_out);
  int rc = mmd_device_check();
  if (rc != MMD_OK) return rc;
  MergeSrc src{};
  src.keys = nullptr;
  src.pairs = nullptr;
  src.scores = scores;
  src.idx = idx;
  src.parts = parts;
  src.k_in = k_in;
  src.Q = Q;
  src.part_stride = Q * k_in;
  return launch_merge<false>(src, k_out, 1.0f, 0, out_scores, out_idx, static_cast<cudaStream_t>(stream));
}

extern "C" int mmd_topk_merge_pairs(const void* pairs, int parts, int64_t part_stride_pairs, int64_t Q, int k_in, int k_out,
                                    int parts_sorted, float* out_scores, int32_t* out_idx, void* stream) {
  using namespace mmd;
  MMD_REQUIRE(parts > 0 && Q >= 0 && k_in > 0 && k_out > 0, "mmd_topk_merge_pairs: parts=%d Q=%lld k_in=%d k_out=%d", parts,
              (long long)Q, k_in, k_out);
  if (Q == 0) return MMD_OK;
  MMD_REQUIRE(pairs != nullptr && out_scores != nullptr && out_idx != nullptr, "mmd_topk_merge_pairs: null buffer");
  MMD_REQUIRE(reinterpret_cast<uintptr_t>(pairs) % 8 == 0, "mmd_topk_merge_pairs: pairs must be 8-byte aligned");
  if (part_stride_pairs == 0) part_stride_pairs = Q * k_in;
  MMD_REQUIRE(part_stride_pairs >= Q * k_in, "mmd_topk_merge_pairs: part stride %lld < Q*k_in", (long long)part_stride_pairs);
  MMD_REQUIRE((k_in <= 128 && k_out <= 128) || static_cast<int64_t>(parts) * k_in <= 4096,
              "mmd_topk_merge_pairs: parts*k_in = %lld exceeds 4096 (only lists of <= 128 entries stream)", (long long)parts * k_in);
  MMD_REQUIRE(k_out <= 4096, "mmd_topk_merge_pairs: k_out %d exceeds 4096", k_out);
  int rc = mmd_device_check();
  if (rc != MMD_OK) return rc;
  MergeSrc src{};
  src.keys = nullptr;
  src.pairs = static_cast<const int2*>(pairs);
  src.scores = nullptr;
  src.idx = nullptr;
  src.parts = parts;
  src.k_in = k_in;
  src.Q = Q;
  src.part_stride = part_stride_pairs;
  // lists this library produced are already descending: skip the per-part sort (lists longer than 128 are always sorted)
  if (parts_sorted) return launch_merge<true>(src, k_out, 1.0f, 0, out_scores, out_idx, static_cast<cudaStream_t>(stream));
  return launch_merge<false>(src, k_out, 1.0f, 0, out_scores, out_idx, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------- pack + scatter of ranked lists
namespace mmd {
namespace {
constexpr int kMaxScatterDst = 16;
struct ScatterDst {
  int2* dst[kMaxScatterDst];
  int n;
  int64_t offset;
};
__global__ void __launch_bounds__(256)
scatter_pairs_kernel(const float* __restrict__ scores, const int32_t* __restrict__ idx, int64_t total, ScatterDst d) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * 256) {
    const int2 v = make_int2(__float_as_int(scores[i]), idx[i]);
    for (int t = 0; t < d.n; ++t) d.dst[t][d.offset + i] = v;
  }
}
}  // namespace
}  // namespace mmd

extern "C" int mmd_scatter_pairs(const float* scores, const int32_t* idx, int64_t Q, int k, void* const* dst_host, int n_dst,
                                 int64_t dst_offset_pairs, void* stream) {
  using namespace mmd;
  MMD_REQUIRE(Q >= 0 && k > 0 && n_dst >= 1 && n_dst <= kMaxScatterDst && dst_offset_pairs >= 0,
              "mmd_scatter_pairs: Q=%lld k=%d n_dst=%d", (long long)Q, k, n_dst);
  if (Q == 0) return MMD_OK;
  MMD_REQUIRE(scores != nullptr && idx != nullptr && dst_host != nullptr, "mmd_scatter_pairs: null buffer");
  int rc = mmd_device_check();
  if (rc != MMD_OK) return rc;
  ScatterDst d{};
  d.n = n_dst;
  d.offset = dst_offset_pairs;
  for (int t = 0; t < n_dst; ++t) {
    MMD_REQUIRE(dst_host[t] != nullptr && reinterpret_cast<uintptr_t>(dst_host[t]) % 8 == 0,
                "mmd_scatter_pairs: destination %d is null or not 8-byte aligned", t);
    d.dst[t] = static_cast<int2*>(dst_host[t]);
  }
  const int64_t total = Q * k;
  const unsigned grid = static_cast<unsigned>(ceil_div(total, 256) < 2368 ? ceil_div(total, 256) : 2368);
  scatter_pairs_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(scores, idx, total, d);
  count_launch();
  MMD_CUDA_OK(cudaGetLastError());
  return MMD_OK;
}
