// K4 of the row-sharded path: the exchange step done by the kernels themselves over peer-mapped memory.
//
// One sharded search step is three stages per rank (the reference has no counterpart: single process, single device; the
// step replaces its per-corpus-chunk heap merge inside sentence_transformers.util.semantic_search, called from
// src/evidence/text2text_retrieval.py:56-64, and the list concat + sort of text2text_retrieval.py:97-110):
//
//   stage C  (topk_fused.cu + topk_merge.cu) tensor-core pass over this rank's shard; the strip merge stores the rank's raw
//            candidate list {score bits, global row} straight into EVERY rank's gather buffer and arrives on flag set 1.
//   stage X  exchange_rescore_kernel (here): waits for flag set 1, merges the `world` candidate lists of each query into the
//            global candidate list (identical on every rank), re-scores in fp32 -- from the original embeddings -- exactly
//            those candidates that live in THIS rank's shard, and stores each exact {score, row} at the candidate's position
//            in every rank's re-score buffer [Q][kc]; arrives on flag set 2.  Every global candidate is owned by exactly one
//            rank, so the second exchange moves kc pairs per query in total instead of world * k.
//   stage F  exchange_finish_kernel (here): waits for flag set 2, sorts each query's kc exact keys, writes the final top-k.
//
// No barrier launches, no collective: the only cross-GPU traffic is the stores of stages C and X (NVLink 5 / NVSwitch)
// and one flag word per rank and stage.  See peer_sync.cuh for the synchronisation and its safety argument.
#include "common.cuh"
#include "peer_sync.cuh"
#include "rescore_common.cuh"
#include "warp_sort.cuh"

namespace mmd {
namespace {

struct ExchangeArgs {
  const int2* gathered;          // local gather buffer: part p (rank p's list) starts p * part_stride pairs in; [Q][k_in] each
  int parts, k_in, kc;
  int64_t part_stride, Q;
  Segments sg;                   // query / corpus source embeddings of this rank (1..4 modalities)
  int64_t N, idx_offset;         // this rank's shard: global rows [idx_offset, idx_offset + N)
  int2* dst[kMaxPeers];          // every rank's re-score buffer [Q][kc]
  int n_dst, own;                // dst[own] is the local one: it also receives the empty slots
  int64_t dst_offset;
  PeerWait wait;
  PeerArrive arrive;
};

__device__ __forceinline__ uint64_t pair_to_key(int2 v) {
  return v.y < 0 ? 0ull : make_key(__int_as_float(v.x), static_cast<uint32_t>(v.y));
}

template <int E>
__global__ void __launch_bounds__(128) exchange_rescore_kernel(const ExchangeArgs a) {
  constexpr int L = 32 * E;
  peer_wait_all(a.wait);
  const int lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (q < a.Q) {
    // ---- merge the ranks' sorted candidate lists (streaming bitonic merge, as merge_stream_kernel)
    uint64_t best[E];
#pragma unroll
    for (int e = 0; e < E; ++e) best[e] = 0ull;
    for (int part = 0; part < a.parts; ++part) {
      const int2* list = a.gathered + static_cast<int64_t>(part) * a.part_stride + q * a.k_in;
      uint64_t nw[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int j = L - 1 - (e * 32 + lane);
        nw[e] = j < a.k_in ? pair_to_key(__ldcg(list + j)) : 0ull;
      }
#pragma unroll
      for (int e = 0; e < E; ++e) best[e] = best[e] > nw[e] ? best[e] : nw[e];
      warp_bitonic_merge_desc<E>(best, lane);
    }
    // ---- exact fp32 score of the candidates this rank owns, stored at the candidate's position on every rank
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      const uint64_t key = i < a.kc ? best[e] : 0ull;
      const int32_t grow = static_cast<int32_t>(key_row(key));
      const int64_t lrow = static_cast<int64_t>(grow) - a.idx_offset;
      const bool mine = key != 0ull && lrow >= 0 && lrow < a.N;
      uint32_t m = __ballot_sync(kWarpFull, mine);
      float sc = 0.0f;
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int32_t gb = __shfl_sync(kWarpFull, grow, b);
        const float s = segments_score(a.sg, q, static_cast<int64_t>(gb) - a.idx_offset, lane);
        if (lane == b) sc = s;
      }
      const int64_t pos = a.dst_offset + q * a.kc + i;
      if (mine) {
        const int2 v = make_int2(__float_as_int(key_score(make_key(sc, static_cast<uint32_t>(grow)))), grow);
        for (int d = 0; d < a.n_dst; ++d) a.dst[d][pos] = v;
      } else if (i < a.kc && key == 0ull) {
        a.dst[a.own][pos] = make_int2(static_cast<int>(0xff800000u), -1);
      }
    }
  }
  peer_arrive_all(a.arrive);
}

struct FinishArgs {
  const int2* resc;              // local re-score buffer [Q][kc]
  int64_t Q;
  int kc, k_out;
  float* out_s;                  // [Q, k_out]
  int32_t* out_i32;              // exactly one of the two index outputs is non-null
  int64_t* out_i64;
  PeerWait wait;
  PeerArrive bump;               // n = 0: only counts this stage's launches
};

template <int E>
__global__ void __launch_bounds__(128) exchange_finish_kernel(const FinishArgs a) {
  peer_wait_all(a.wait);
  const int lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (q < a.Q) {
    uint64_t k[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      k[e] = i < a.kc ? pair_to_key(__ldcg(a.resc + q * a.kc + i)) : 0ull;
    }
    warp_bitonic_desc<E>(k, lane);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      if (i < a.k_out) {
        const uint64_t key = k[e];
        const int64_t o = q * a.k_out + i;
        a.out_s[o] = key == 0ull ? __int_as_float(0xff800000) : key_score(key);
        const int32_t ix = key == 0ull ? -1 : static_cast<int32_t>(key_row(key));
        if (a.out_i64 != nullptr) a.out_i64[o] = ix; else a.out_i32[o] = ix;
      }
    }
    for (int i = 32 * E + lane; i < a.k_out; i += 32) {
      const int64_t o = q * a.k_out + i;
      a.out_s[o] = __int_as_float(0xff800000);
      if (a.out_i64 != nullptr) a.out_i64[o] = -1; else a.out_i32[o] = -1;
    }
  }
  peer_arrive_all(a.bump);
}

int fill_wait(PeerWait* w, const uint32_t* flags, int n, const uint32_t* state, const char* who) {
  MMD_REQUIRE(n >= 0 && n <= kMaxPeers && (n == 0 || (flags != nullptr && state != nullptr)), "%s: wait flags null or n_wait=%d (0..%d)",
              who, n, kMaxPeers);
  w->flags = flags;
  w->n = n;
  w->state = state;
  return MMD_OK;
}

}  // namespace

int fill_arrive(PeerArrive* a, void* const* flags_host, int n, uint32_t* state, const char* who) {
  MMD_REQUIRE(n >= 0 && n <= kMaxPeers && (n == 0 || flags_host != nullptr), "%s: arrive flags null or n_arrive=%d (0..%d)", who, n,
              kMaxPeers);
  MMD_REQUIRE(n == 0 || state != nullptr, "%s: arrive flags need a state word", who);
  a->n = n;
  a->state = state;
  for (int i = 0; i < n; ++i) {
    MMD_REQUIRE(flags_host[i] != nullptr && reinterpret_cast<uintptr_t>(flags_host[i]) % 4 == 0, "%s: arrive flag %d null or misaligned",
                who, i);
    a->flag[i] = static_cast<uint32_t*>(flags_host[i]);
  }
  return MMD_OK;
}

}  // namespace mmd

extern "C" int mmd_exchange_rescore(const void* gathered, int parts, int64_t part_stride_pairs, int64_t Q, int k_in, int kc,
                                    int n_seg, const void* const* q_src_host, const int* q_dtype_host,
                                    const int64_t* q_stride_host, const float* const* q_inv_host,
                                    const void* const* c_src_host, const int* c_dtype_host, const int64_t* c_stride_host,
                                    const float* const* c_inv_host, const int* dim_host, const float* weight_host, int64_t N,
                                    int64_t idx_offset, void* const* dst_host, int n_dst, int own_dst, int64_t dst_offset_pairs,
                                    const uint32_t* wait_flags, int n_wait, void* const* arrive_flags_host, int n_arrive,
                                    uint32_t* sync_state, void* stream) {
  using namespace mmd;
  const char* who = "mmd_exchange_rescore";
  MMD_REQUIRE(parts >= 1 && parts <= kMaxPeers && Q >= 0 && k_in > 0 && kc > 0 && kc <= 128 && N >= 0 && idx_offset >= 0,
              "%s: parts=%d Q=%lld k_in=%d kc=%d (kc <= 128) N=%lld", who, parts, (long long)Q, k_in, kc, (long long)N);
  MMD_REQUIRE(k_in <= 128, "%s: k_in %d exceeds 128", who, k_in);
  if (Q == 0) return MMD_OK;
  MMD_REQUIRE(gathered != nullptr && reinterpret_cast<uintptr_t>(gathered) % 8 == 0, "%s: gather buffer null or misaligned", who);
  if (part_stride_pairs == 0) part_stride_pairs = Q * k_in;
  MMD_REQUIRE(part_stride_pairs >= Q * k_in, "%s: part stride %lld < Q*k_in", who, (long long)part_stride_pairs);
  MMD_REQUIRE(dst_host != nullptr && n_dst >= 1 && n_dst <= kMaxPeers && own_dst >= 0 && own_dst < n_dst && dst_offset_pairs >= 0,
              "%s: n_dst=%d own_dst=%d", who, n_dst, own_dst);
  MMD_REQUIRE(sync_state != nullptr || (n_wait == 0 && n_arrive == 0), "%s: flags need a sync_state word pair", who);
  ExchangeArgs a{};
  int rc = fill_segments(&a.sg, n_seg, q_src_host, q_dtype_host, q_stride_host, q_inv_host, c_src_host, c_dtype_host, c_stride_host,
                         c_inv_host, dim_host, weight_host, N, who);
  if (rc != MMD_OK) return rc;
  for (int d = 0; d < n_dst; ++d) {
    MMD_REQUIRE(dst_host[d] != nullptr && reinterpret_cast<uintptr_t>(dst_host[d]) % 8 == 0, "%s: destination %d null or misaligned", who, d);
    a.dst[d] = static_cast<int2*>(dst_host[d]);
  }
  rc = fill_wait(&a.wait, wait_flags, n_wait, sync_state, who);
  if (rc != MMD_OK) return rc;
  rc = fill_arrive(&a.arrive, arrive_flags_host, n_arrive, sync_state, who);
  if (rc != MMD_OK) return rc;
  rc = mmd_device_check();
  if (rc != MMD_OK) return rc;
  a.gathered = static_cast<const int2*>(gathered);
  a.parts = parts; a.k_in = k_in; a.kc = kc; a.part_stride = part_stride_pairs; a.Q = Q;
  a.N = N; a.idx_offset = idx_offset;
  a.n_dst = n_dst; a.own = own_dst; a.dst_offset = dst_offset_pairs;
  const unsigned grid = static_cast<unsigned>(ceil_div(Q, 4));
  auto st = static_cast<cudaStream_t>(stream);
  const int longest = k_in > kc ? k_in : kc;
  if (longest <= 32) exchange_rescore_kernel<1><<<grid, 128, 0, st>>>(a);
  else if (longest <= 64) exchange_rescore_kernel<2><<<grid, 128, 0, st>>>(a);
  else exchange_rescore_kernel<4><<<grid, 128, 0, st>>>(a);
  count_launch();
  MMD_CUDA_OK(cudaGetLastError());
  return MMD_OK;
}

extern "C" int mmd_exchange_finish(const void* rescored, int64_t Q, int kc, int k_out, float* out_scores, void* out_idx,
                                   int idx_is_i64, const uint32_t* wait_flags, int n_wait, uint32_t* sync_state, void* stream) {
  using namespace mmd;
  const char* who = "mmd_exchange_finish";
  MMD_REQUIRE(Q >= 0 && kc > 0 && kc <= 128 && k_out > 0, "%s: Q=%lld kc=%d (<= 128) k_out=%d", who, (long long)Q, kc, k_out);
  if (Q == 0) return MMD_OK;
  MMD_REQUIRE(rescored != nullptr && reinterpret_cast<uintptr_t>(rescored) % 8 == 0 && out_scores != nullptr && out_idx != nullptr,
              "%s: null or misaligned buffer", who);
  MMD_REQUIRE(sync_state != nullptr || n_wait == 0, "%s: flags need a sync_state word pair", who);
  FinishArgs a{};
  int rc = fill_wait(&a.wait, wait_flags, n_wait, sync_state, who);
  if (rc != MMD_OK) return rc;
  a.bump.n = 0;
  a.bump.state = sync_state;
  rc = mmd_device_check();
  if (rc != MMD_OK) return rc;
  a.resc = static_cast<const int2*>(rescored);
  a.Q = Q; a.kc = kc; a.k_out = k_out;
  a.out_s = out_scores;
  a.out_i32 = idx_is_i64 ? nullptr : static_cast<int32_t*>(out_idx);
  a.out_i64 = idx_is_i64 ? static_cast<int64_t*>(out_idx) : nullptr;
  const unsigned grid = static_cast<unsigned>(ceil_div(Q, 4));
  auto st = static_cast<cudaStream_t>(stream);
  if (kc <= 32) exchange_finish_kernel<1><<<grid, 128, 0, st>>>(a);
  else if (kc <= 64) exchange_finish_kernel<2><<<grid, 128, 0, st>>>(a);
  else exchange_finish_kernel<4><<<grid, 128, 0, st>>>(a);
  count_launch();
  MMD_CUDA_OK(cudaGetLastError());
  return MMD_OK;
}
