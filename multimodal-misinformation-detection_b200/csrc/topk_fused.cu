// K2+K3: query x corpus score contraction on tcgen05 tensor cores with the top-K selection fused
// into the TMEM epilogue, so the Q x N score matrix never reaches HBM.
//
// Reference arithmetic this replaces (paths relative to the reference checkout):
//   torch.mm(a_norm, b_norm.T) + torch.topk + per-query heapq merge inside
//     sentence_transformers.util.semantic_search   (call sites src/evidence/text2text_retrieval.py:56-64,
//                                                   src/evidence/experiment_text.py:25-33)
//   the O(N) python loop + full sort of               src/evidence/im2im_retrieval.py:84-92,
//                                                     src/evidence/experiment_image.py:25-33
//
// Shape of the kernel (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0 (one lane)  TMA producer: {128 queries x 128 B} and {256 corpus rows x 128 B} boxes, 128B swizzle,
//                      multi-stage ring of mbarriers.
//   warp 1 (one lane)  MMA issuer: tcgen05.mma 128x256x(32 B) into one of two 256-column TMEM accumulators.
//   warps 2..5         epilogue: thread <-> query row.  Per tile a branch-free FILTER pass (tcgen05.ld 32 lanes x 32 columns,
//                      two loads in flight; the maxima of the tile's 64 groups of 4 columns against the row thresholds,
//                      OR-ed over the warp) names the groups worth a second look; a COLLECT pass re-reads only those
//                      (tcgen05.ld .x4) and appends scores above the row's threshold to a per-row candidate buffer in
//                      shared memory (predicated stores).  A full row is compacted alone: short lists by a shuffle bitonic
//                      sort, long lists by a 4-ary selection on the score word.  At the end of a unit every thread ranks
//                      its own row and stores the sorted K-list.
// Work decomposition: unit = (128-query tile, strip of T corpus tiles), query tile fastest so that the
// CTAs running at the same time read the same corpus rows (L2 reuse; HBM sees the corpus ~once).
// Every unit leaves a sorted K-list per query; topk_merge.cu folds the strips together.
#include <cuda_bf16.h>

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <utility>
#include <vector>

#include "common.cuh"
#include "peer_sync.cuh"
#include "ptx.cuh"

namespace mmd {

int merge_partial_keys(const uint64_t* partial, int parts, int64_t Q, int k_in, int k_out, float scale,
                       int64_t idx_offset, float* out_scores, int32_t* out_idx, cudaStream_t stream,
                       const PairOut* po = nullptr, const PeerArrive* arrive = nullptr);
int fill_arrive(PeerArrive* a, void* const* flags_host, int n, uint32_t* state, const char* who);

namespace {

constexpr int kTileM = 128;          // queries per tile  (UMMA M, TMEM lanes)
constexpr int kTileN = 256;          // corpus rows per tile (UMMA N, TMEM columns)
constexpr int kBlockKBytes = 128;    // one swizzle atom along K per stage
// Corpus bytes per CTA and stage: the whole 256-row tile alone (32 KB), or half of it (16 KB) when two
// CTAs of a pair run one cta_group::2 MMA over 256 queries x 256 corpus rows.
__host__ __device__ constexpr int b_bytes(int cta) { return (kTileN / cta) * kBlockKBytes; }
constexpr int kMaxStages = 8;
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;       // two 256-column fp32 accumulators
constexpr int kMaxSmem = 232448;     // 227 KB opt-in limit per CTA on sm_100
constexpr int kMaxLevels = 4;        // merged bounds: quantile levels 1..4 (2 + 4 + 8 + 16 slots per query)

struct FusedParams {
  int64_t Q, N;
  int kblocks;           // ceil(row_bytes / 128)
  int n_m, n_n;          // query tiles, corpus tiles
  int tiles_per_strip;   // T
  int n_strips;          // S
  int n_units;           // n_m * S
  int kprime;            // list length kept per (query, strip)
  int stages;
  int a_rows;            // query rows staged per CTA and K block: 128, or the 32-row groups a batch of < 128 queries fills
  int key_warps;         // epilogue warps (lane quarters 0 .. key_warps-1) that own candidate buffers: a_rows / 32
  int keep_hi;           // long lists (CAP > 64): a compaction by selection leaves between kprime and keep_hi entries
  int early_level;       //   ... and takes along every row of the warp that holds more than this many
  int levels;            // merged bounds: levels 1..levels of quantile slots in lvl (0 = off)
  uint64_t* partial;     // [Q][S][kprime] keys
  uint32_t* thr_global;  // [Q] shared lower bounds of every query's K-th best (ordered uint, 0 = none yet)
  uint32_t* lvl;         // [level_slots(levels)][Q] ordered score words, 0 = empty (see "merged bounds")
  uint32_t* thr_peer[8]; // row-sharded corpus: the same array on EVERY rank (peer-mapped, own one included); a bound
  int n_thr_peers;       //   learnt on one shard prunes all of them.  0 = publish to thr_global only.
  float* dense;          // dense mode: [Q][ldd] scores
  int64_t ldd;
  float out_scale;       // dense mode: multiplier applied to the accumulators
  DeviceStatus* status;
  unsigned long long* stats;   // developer builds (-DMMD_STATS): per-CTA wait-cycle counters, else unused
};

#ifdef MMD_STATS
// developer timeline: block 0 records clock64() at a few points of every tile (up to 128 tiles):
//   stats[tile*8 + 0] MMA: before tmem_empty wait   [1] MMA: after wait   [2] MMA: all MMAs of the tile issued
//   stats[tile*8 + 4] EPI(warp 2): before tmem_full wait   [5] after wait   [6] tile processed
#define MMD_TRACE(tile, slot) do { if (p.stats && blockIdx.x == 0 && (tile) < 128) p.stats[(tile) * 8 + (slot)] = clock64(); } while (0)
// ... and whole-launch aggregates (summed over CTAs / epilogue warps) in stats[1024 + slot]:
//   MMA issuer:  0 cycles waiting for a free accumulator (epilogue-bound)   1 waiting for operands (load-bound)   2 busy in total
//   epilogue:    8 waiting for an accumulator   9 cold start   10 filter pass   11 collect pass without compactions
//                12 compactions   13 end of unit (final sort, publication)   14 busy in total
//                16 tiles   17 hit chunks   18 group visits   19 compaction calls   20 rows compacted   21 appends   22 units
enum { kStMmaEmpty = 0, kStMmaFull = 1, kStMmaTotal = 2, kStEpiWait = 8, kStEpiCold = 9, kStEpiFilter = 10, kStEpiCollect = 11,
       kStEpiCompact = 12, kStEpiFinal = 13, kStEpiTotal = 14, kStTiles = 16, kStHitChunks = 17, kStGroups = 18, kStCompCalls = 19,
       kStCompRows = 20, kStAppends = 21, kStUnits = 22, kStSlots = 24 };
#define MMD_ST_DECL long long st_acc[kStSlots] = {0}; long long st_t = 0; (void)st_t
#define MMD_ST_T0() do { st_t = clock64(); } while (0)
#define MMD_ST_ACC(slot) do { const long long now_ = clock64(); st_acc[slot] += now_ - st_t; st_t = now_; } while (0)
#define MMD_ST_ADD(slot, v) do { st_acc[slot] += (v); } while (0)
#define MMD_ST_FLUSH() do { if (p.stats && lane == 0) for (int i_ = 0; i_ < kStSlots; ++i_) if (st_acc[i_]) atomicAdd(p.stats + 1024 + i_, (unsigned long long)st_acc[i_]); } while (0)
#else
#define MMD_TRACE(tile, slot) do { } while (0)
#define MMD_ST_DECL do { } while (0)
#define MMD_ST_T0() do { } while (0)
#define MMD_ST_ACC(slot) do { } while (0)
#define MMD_ST_ADD(slot, v) do { } while (0)
#define MMD_ST_FLUSH() do { } while (0)
#endif

struct SmemLayout {
  uint32_t stage_off;    // stages x {A,B}
  uint32_t keys_off;     // 4 warps x CAP x 32 keys
  uint32_t bars_off;     // full[8], empty[8], tmem_full[2], tmem_empty[2]
  uint32_t tmem_ptr_off;
  uint32_t total;        // including 1024 B of alignment slack
};
// a_rows / key_warps: see FusedParams.  A small batch (the reference's one-claim-per-call pattern) is an HBM-bound corpus
// stream; staging only the query rows that exist (4 KB instead of 16 KB per stage for <= 32 queries) and only their
// candidate buffers (16 KB instead of 64 KB) buys two more TMA stages, i.e. 160 KB instead of 96 KB of corpus in flight per SM.
__host__ __device__ inline SmemLayout smem_layout(int stages, int cap, int cta, int a_rows = kTileM, int key_warps = 4) {
  SmemLayout l;
  l.stage_off = 0;
  l.keys_off = stages * (a_rows * kBlockKBytes + b_bytes(cta));
  l.bars_off = l.keys_off + cap * 32 * key_warps * 8;
  l.tmem_ptr_off = l.bars_off + (2 * kMaxStages + 4) * 8;
  l.total = l.tmem_ptr_off + 16 + 1024;
  return l;
}

// ---------------------------------------------------------------- warp-cooperative list maintenance
// Candidate (slot, row r of this warp) lives at wkeys[slot * 32 + r], as a RAW entry {lo = ~column, hi = fp32 score bits}:
// an append is then four instructions (compare, address, predicated 8-byte store, predicated count) with no
// bank conflicts for any mix of per-row counts (slots are 256 B apart, a multiple of the 128 B bank cycle).  Entries
// become ordered keys only when a row is compacted; the warp-wide read of one row there is a 32-way bank conflict,
// paid twice per compaction instead of once per appended score.
__device__ __forceinline__ int key_slot_index(int slot, int r) { return slot * 32 + r; }
__device__ __forceinline__ uint64_t raw_to_key(uint64_t raw) {
  return raw == 0ull ? 0ull
                     : (static_cast<uint64_t>(float_to_ordered(__uint_as_float(static_cast<uint32_t>(raw >> 32)))) << 32) |
                           (raw & 0xffffffffull);
}
__device__ __forceinline__ uint64_t key_to_raw(uint64_t key) {
  return key == 0ull ? 0ull
                     : (static_cast<uint64_t>(__float_as_uint(ordered_to_float(static_cast<uint32_t>(key >> 32)))) << 32) |
                           (key & 0xffffffffull);
}

// R rows at a time go through the same bitonic network (independent dependency chains: one epilogue warp per
// scheduler has nobody else to hide the shuffle latency behind).  Element index i = e * 32 + lane over 32 * E
// elements per row; result descending in i.
template <int E, int R>
__device__ __forceinline__ void bitonic_sort_desc(uint64_t (&k)[R][E], int lane) {
#pragma unroll
  for (int size = 2; size <= 32 * E; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {
        const int es = stride >> 5;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & es) == 0) {
            const bool desc = ((e * 32 + lane) & size) == 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const uint64_t a = k[r][e], b = k[r][e | es];
              const uint64_t mx = a > b ? a : b, mn = a > b ? b : a;
              k[r][e] = desc ? mx : mn;
              k[r][e | es] = desc ? mn : mx;
            }
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const bool lower = (lane & stride) == 0;
          const bool desc = ((e * 32 + lane) & size) == 0;
          const bool take_max = (lower == desc);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const uint64_t other = __shfl_xor_sync(kFullMask, k[r][e], stride);
            const uint64_t mine = k[r][e];
            const bool mine_gt = mine > other;
            k[r][e] = (mine_gt == take_max) ? mine : other;
          }
        }
      }
    }
  }
}

struct RowState {
  int cnt;
  float thr;
};

// Quantile bounds a finished unit leaves behind (see "merged bounds" above the kernel): where to publish them.
struct LevelPub {
  uint32_t* lvl;      // nullptr: nothing to publish
  int64_t Q;
  int64_t row0;       // query row of this warp's row 0
  int strip;
  int levels;
};
__host__ __device__ constexpr int level_slot0(int j) { return (1 << j) - 2; }        // first slot of level j >= 1
__host__ __device__ constexpr int level_slots(int levels) { return (2 << levels) - 2; }

// Compaction by sorting (short lists): the row's entries spread over the lanes (E per lane), shuffle bitonic network, the best
// `kprime` written back in order; the row's threshold becomes its exact kprime-th best -- the tightest filter a row can have.
template <int CAP>
__device__ __forceinline__ void sort_row(uint64_t* wkeys, int row, int lane, int kprime, int& cnt, float& thr) {
  constexpr int E = CAP / 32;
  uint64_t k[1][E];
  const int n = __shfl_sync(kFullMask, cnt, row);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    k[0][e] = (i < n) ? raw_to_key(wkeys[key_slot_index(i, row)]) : 0ull;
  }
  bitonic_sort_desc<E, 1>(k, lane);
  const int keep = n < kprime ? n : kprime;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    if (i < keep) wkeys[key_slot_index(i, row)] = key_to_raw(k[0][e]);
  }
  const int last = kprime - 1;
  uint64_t kk = k[0][0];
#pragma unroll
  for (int e = 1; e < E; ++e) kk = ((last >> 5) == e) ? k[0][e] : kk;
  kk = __shfl_sync(kFullMask, kk, last & 31);
  if (lane == row) {
    cnt = keep;
    if (n >= kprime) thr = fmaxf(thr, key_score(kk));   // never below a bound learnt from other CTAs
  }
}

// Selection instead of a sort (long lists, where a compaction frees few slots and must be cheap): the row's entries are
// spread over the lanes (E per lane); bisection on the order-preserving score word, counting "entries above the midpoint"
// with one warp reduction per step, finds a bound with between kprime and keep_hi entries above it; those are written back
// densely (ballot prefix), the bound becomes the row's threshold.  The buffer stays unsorted -- only the end of a unit
// needs the order.  ~10 steps of E compares + 1 reduction for R rows at once instead of a 32*E-wide bitonic network.
// Returns the mask of rows that did not converge (more equal scores around the kprime-th best than the window holds):
// the caller sorts those.
template <int CAP, int R>
__device__ __forceinline__ uint32_t select_batch(uint64_t* wkeys, const int (&rows)[R], int lane, int kprime, int keep_hi,
                                                 int& cnt, float& thr) {
  constexpr int E = CAP / 32;
  uint32_t ord[R][E], col[R][E];
  uint32_t lo[R], hi[R];
  int c_lo[R];
  bool done[R], ok[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int n = __shfl_sync(kFullMask, cnt, rows[r]);
    uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      const uint64_t raw = (i < n) ? wkeys[key_slot_index(i, rows[r])] : 0ull;
      const uint32_t o = (i < n) ? float_to_ordered(__uint_as_float(static_cast<uint32_t>(raw >> 32))) : 0u;
      ord[r][e] = o;                                   // real scores map to >= 0x007fffff; 0 = empty slot
      col[r][e] = static_cast<uint32_t>(raw);
      mn = (i < n) ? min(mn, o) : mn;
      mx = max(mx, o);
    }
    // invariant: count(> lo) = c_lo >= kprime (the row holds at least kprime entries when it is compacted), count(> hi) < kprime
    lo[r] = __reduce_min_sync(kFullMask, mn) - 1u;
    hi[r] = __reduce_max_sync(kFullMask, mx);
    c_lo[r] = n;
    ok[r] = n <= keep_hi || n < kprime;
    done[r] = ok[r];
  }
  // three probes per step (the interval shrinks to a quarter; the three warp reductions are independent, and a step is a
  // chain of dependent instructions with nothing else to issue in between)
#pragma unroll 1
  for (int it = 0; it < 40; ++it) {
    bool all_done = true;
#pragma unroll
    for (int r = 0; r < R; ++r) all_done = all_done && done[r];
    if (all_done) break;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint32_t d = hi[r] - lo[r];
      const uint32_t q = d >> 2;
      // d < 4: one midpoint (three equal probes); d = 1: the midpoint is lo itself -> equal scores straddle the kprime-th best
      const uint32_t t1 = q ? lo[r] + q : lo[r] + (d >> 1), t2 = q ? t1 + q : t1, t3 = q ? t2 + q : t1;
      int c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        c1 += (ord[r][e] > t1) ? 1 : 0;
        c2 += (ord[r][e] > t2) ? 1 : 0;
        c3 += (ord[r][e] > t3) ? 1 : 0;
      }
      c1 = __reduce_add_sync(kFullMask, c1);
      c2 = __reduce_add_sync(kFullMask, c2);
      c3 = __reduce_add_sync(kFullMask, c3);
      if (!done[r]) {
        if (t1 == lo[r]) {
          done[r] = true;
        } else {
          if (c3 >= kprime) { lo[r] = t3; c_lo[r] = c3; }
          else if (c2 >= kprime) { lo[r] = t2; c_lo[r] = c2; hi[r] = t3; }
          else if (c1 >= kprime) { lo[r] = t1; c_lo[r] = c1; hi[r] = t2; }
          else { hi[r] = t1; }
          done[r] = ok[r] = c_lo[r] <= keep_hi;
        }
      }
    }
  }
  uint32_t failed = 0u;
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (!ok[r]) {
      failed |= 1u << rows[r];
      continue;
    }
    if (c_lo[r] < kprime || lo[r] + 1u == 0u) continue;      // (short row: left alone)
    int base = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool keep = ord[r][e] > lo[r];
      const uint32_t b = __ballot_sync(kFullMask, keep);
      const int pos = base + __popc(b & lt);
      if (keep)
        wkeys[key_slot_index(pos, rows[r])] =
            (static_cast<uint64_t>(__float_as_uint(ordered_to_float(ord[r][e]))) << 32) | col[r][e];
      base += __popc(b);
    }
    if (lane == rows[r]) {
      cnt = base;
      thr = fmaxf(thr, ordered_to_float(lo[r]));   // kprime entries of the row lie strictly above it
    }
  }
  return failed;
}

// End of a unit, a warp with only a few query rows (the reference's one-claim-per-call pattern: ONE row): the rows one at a
// time, entries spread over the lanes (E per lane), every lane ranks its entries against the whole row (broadcast loads).
template <int CAP>
__device__ __forceinline__ void final_row(uint64_t* wkeys, int row, int lane, int kprime, int& cnt, float& thr, uint64_t* out,
                                          const LevelPub& pub) {
  constexpr int E = CAP / 32;
  const int n = __shfl_sync(kFullMask, cnt, row);
  uint64_t mine[E];
  int rank[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    mine[e] = (i < n) ? raw_to_key(wkeys[key_slot_index(i, row)]) : 0ull;
    rank[e] = 0;
  }
#pragma unroll 4
  for (int j = 0; j < n; ++j) {
    const uint64_t kj = raw_to_key(wkeys[key_slot_index(j, row)]);       // same address for the whole warp: a broadcast
#pragma unroll
    for (int e = 0; e < E; ++e) rank[e] += (kj > mine[e]) ? 1 : 0;
  }
  const int last = kprime - 1;
  uint32_t kth = 0u;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (mine[e] == 0ull) continue;
    if (rank[e] < kprime) out[rank[e]] = mine[e];
    const uint32_t ord = static_cast<uint32_t>(mine[e] >> 32);
    if (rank[e] == last) kth = ord;
    if (pub.lvl != nullptr)
      for (int j = 1; j <= pub.levels; ++j)
        if (rank[e] == ((kprime + (1 << j) - 1) >> j) - 1)
          atomicMax(pub.lvl + static_cast<int64_t>(level_slot0(j) + (pub.strip & ((1 << j) - 1))) * pub.Q + pub.row0 + row, ord);
  }
  for (int i = n + lane; i < kprime; i += 32) out[i] = 0ull;               // a short list ends in empty slots
  kth = __reduce_max_sync(kFullMask, kth);
  if (lane == row) {
    cnt = n < kprime ? n : kprime;
    if (n >= kprime) thr = fmaxf(thr, ordered_to_float(kth));
  }
}

// End of a unit: every row's sorted K'-list, by RANKING, thread <-> row, all 32 rows of the warp at once.  A thread reads
// only its own row (conflict-free in the [slot][row] layout): it turns its entries into ordered keys in place, then, eight
// entries at a time, counts how many entries of the row are larger (one shared-memory load per entry of the row, eight
// independent 64-bit compares per load) and stores each entry straight to its rank in global memory.  n^2 compares per
// row, but 32 rows in lockstep and no shuffles: ~4x fewer cycles per unit than the warp-cooperative network (which
// handled one row at a time at ~2 k cycles each, during which the accumulators of the next unit's first tiles wait).
template <int CAP>
__device__ __noinline__ RowState final_lists(uint64_t* wkeys, int lane, int kprime, bool valid, int cnt, float thr,
                                             uint64_t* out, int64_t out_row_stride, LevelPub pub) {
  uint32_t vmask = __ballot_sync(kFullMask, valid);
  if (__popc(vmask) <= 6) {
    // few rows: a lockstep pass over the longest row would be one thread's work at a warp's cost
#pragma unroll 1
    while (vmask) {
      const int row = __ffs(vmask) - 1;
      vmask &= vmask - 1;
      final_row<CAP>(wkeys, row, lane, kprime, cnt, thr, out + static_cast<int64_t>(row - lane) * out_row_stride, pub);
    }
    __syncwarp();
    return RowState{cnt, thr};
  }
  uint64_t* own = wkeys + lane;                         // slot s of this thread's row: own[s * 32]
  const int n = valid ? cnt : 0;
  const int nmax = __reduce_max_sync(kFullMask, n);
#pragma unroll 4
  for (int i = 0; i < nmax; ++i) {
    const uint64_t raw = own[i * 32];
    own[i * 32] = (i < n) ? raw_to_key(raw) : 0ull;     // slots beyond the own count read as empty below
  }
  const int last = kprime - 1;
  int lvl_idx[kMaxLevels];
#pragma unroll
  for (int j = 1; j <= kMaxLevels; ++j) lvl_idx[j - 1] = (pub.lvl != nullptr && j <= pub.levels) ? ((kprime + (1 << j) - 1) >> j) - 1 : -1;
  uint32_t kth = 0u;
#pragma unroll 1
  for (int i0 = 0; i0 < nmax; i0 += 8) {
    uint64_t ki[8];
    int rank[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      ki[t] = (i0 + t < nmax) ? own[(i0 + t) * 32] : 0ull;
      rank[t] = 0;
    }
#pragma unroll 2
    for (int j = 0; j < nmax; ++j) {
      const uint64_t kj = own[j * 32];
#pragma unroll
      for (int t = 0; t < 8; ++t) rank[t] += (kj > ki[t]) ? 1 : 0;
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (ki[t] == 0ull) continue;                       // keys of a row are distinct: its ranks are a permutation
      if (rank[t] < kprime) out[rank[t]] = ki[t];
      const uint32_t ord = static_cast<uint32_t>(ki[t] >> 32);
      if (rank[t] == last) kth = ord;
#pragma unroll
      for (int j = 1; j <= kMaxLevels; ++j)              // the score at rank ceil(K' / 2^j), for every level j: see merged bounds
        if (rank[t] == lvl_idx[j - 1])
          atomicMax(pub.lvl + static_cast<int64_t>(level_slot0(j) + (pub.strip & ((1 << j) - 1))) * pub.Q + pub.row0 + lane, ord);
    }
  }
  if (valid)
    for (int i = n; i < kprime; ++i) out[i] = 0ull;      // a short list ends in empty slots
  RowState st;
  st.cnt = n < kprime ? n : kprime;
  st.thr = (n >= kprime) ? fmaxf(thr, ordered_to_float(kth)) : thr;          // never below a bound learnt from other CTAs
  __syncwarp();
  return st;
}

// Compaction of the rows named in `mask`, one row at a time in a loop: the rarely executed paths of this kernel run out of a cold
// instruction cache with one warp per scheduler, and one row's worth of code that the following rows reuse beat the
// four-rows-interleaved variants by a wide margin (and so did one instantiation of the network instead of four).
// Short lists are sorted (sort_row); long lists go through the selection (select_batch; keep_hi: see there), rows it
// cannot split through the sort.  Returns the calling lane's own (count, threshold).
template <int CAP>
__device__ __noinline__ RowState compact_rows(uint64_t* wkeys, uint32_t mask, int lane, int kprime, int keep_hi, int cnt, float thr) {
  __syncwarp();
  if constexpr (CAP > 64) {
    uint32_t failed = 0u;
#pragma unroll 1
    while (mask) {
      int rows[1];
      rows[0] = __ffs(mask) - 1;
      mask &= mask - 1;
      failed |= select_batch<CAP, 1>(wkeys, rows, lane, kprime, keep_hi, cnt, thr);
    }
    mask = failed;
    __syncwarp();
  }
#pragma unroll 1
  while (mask) {
    const int row = __ffs(mask) - 1;
    mask &= mask - 1;
    sort_row<CAP>(wkeys, row, lane, kprime, cnt, thr);
  }
  __syncwarp();
  RowState st;
  st.cnt = cnt;
  st.thr = thr;
  return st;
}

// Per-thread bitonic sort (descending) of a small register array; all indices are compile-time constants.
template <int N>
__device__ __forceinline__ void thread_sort_desc(float (&v)[N]) {
#pragma unroll
  for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int j = i ^ stride;
        if (j > i) {
          const bool desc = (i & size) == 0;
          const float a = v[i], b = v[j];
          const float mx = fmaxf(a, b), mn = fminf(a, b);
          v[i] = desc ? mx : mn;
          v[j] = desc ? mn : mx;
        }
      }
    }
  }
}

// Bit g set <=> one of columns 4g .. 4g+3 of the 32 loaded scores lies above the row's threshold (NaN never does).
__device__ __forceinline__ uint32_t group_mask(const uint32_t (&r)[32], float thr) {
  uint32_t m = 0u;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float mx = fmaxf(fmaxf(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1])),
                           fmaxf(__uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3])));
    m |= (mx > thr) ? (1u << g) : 0u;
  }
  return m;
}

// Thresholds are shared between CTAs through thr_global[q] (order-preserving uint, atomicMax).  A published
// value is the next float BELOW a row's current K-th best, so that "score > bound" still admits a score equal
// to that K-th best (the (score desc, row asc) tie rule is decided later, by the sorts).  Any K-th best of a
// subset of the corpus is a valid lower bound for the K-th best of the whole corpus.
// With a row-sharded corpus the K-th best of one GPU's shard is equally a lower bound for the global K-th best, so a
// publication goes to the threshold array of every rank (system-scope atomics over NVLink on peer-mapped memory).
// `seen` = the largest shared bound of this row the thread has read or written so far: most bounds a unit learns are below
// what is already known (the max over all units and ranks), and on a row-sharded corpus a publication is one system-scope
// atomic per rank over NVLink -- those are not sent.  (A register compare: a load here would put an L2 round trip on the
// path of every compaction of a young row.)
__device__ __noinline__ uint32_t publish_threshold(const FusedParams& p, int64_t qrow, float thr, uint32_t seen) {
  const uint32_t v = float_to_ordered(thr) - 1u;
  if (v <= seen) return seen;
  if (p.n_thr_peers == 0) {
    atomicMax(p.thr_global + qrow, v);
  } else {
    for (int i = 0; i < p.n_thr_peers; ++i) atomicMax_system(p.thr_peer[i] + qrow, v);
  }
  return v;
}

// Merged bounds.  thr_global[q] carries the best K'-th best any ONE unit (strip) has seen; the K'-th best of the UNION of
// the strips processed so far is much higher (with s finished strips of n rows each roughly the (K'/s)-th best of one
// strip), and the fraction of scores that pass the filter -- hence collect passes and compactions -- falls with it.  A full
// merge of finished lists inside the sweep would be expensive; quantiles are enough: a finished unit publishes, for
// levels j = 1..J, its score at rank r_j = ceil(K' / 2^j) into slot (strip mod 2^j) of level j (atomicMax).  The slots of a
// level belong to disjoint sets of strips, each slot's value has r_j scores of ONE strip at or above it, so the minimum
// over the 2^j slots of a level has 2^j * r_j >= K' distinct corpus rows at or above it: a valid lower bound of the
// query's K'-th best.  Every finishing unit folds the levels (max over levels of the min over slots) into thr_global,
// where the running units pick it up with their per-tile read.
// ---------------------------------------------------------------- the kernel
// kCta = 1: one CTA per SM, UMMA 128 x 256.
// kCta = 2: clusters of two CTAs (one TPC); the pair runs ONE tcgen05.mma.cta_group::2 of 256 queries x 256
//           corpus rows per K step: each CTA stages its own 128 queries and HALF of the corpus tile, so the
//           L2 -> shared-memory traffic per flop drops by a third and the smem operand reads per SM by a
//           quarter.  CTA 0 (the leader) issues the MMAs; both CTAs own 128 accumulator rows in their TMEM.
template <int CAP, bool kF8, bool kDense, int kCta>
__global__ void __launch_bounds__(kThreads, 1)
fused_score_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                        const FusedParams p, const uint32_t idesc) {
  const int a_bytes = p.a_rows * kBlockKBytes;          // B starts right after the staged query rows (1 KB multiple)
  const int stage_sz = a_bytes + b_bytes(kCta);
  constexpr int kUnitRows = kTileM * kCta;        // queries per unit (per CTA pair)
  constexpr int kBRows = kTileN / kCta;           // corpus rows this CTA stages per tile
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled operand tiles need 1024-byte alignment.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const SmemLayout L = smem_layout(p.stages, kDense ? 0 : CAP, kCta, p.a_rows, p.key_warps);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars_off);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L.tmem_ptr_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = kCta == 2 ? cluster_ctarank() : 0u;
  const int unit0 = static_cast<int>(blockIdx.x) / kCta;          // first unit of this CTA (pair)
  const int unit_stride = static_cast<int>(gridDim.x) / kCta;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);        // the leader's expect_tx arrival; TMA bytes of both CTAs complete it
      mbar_init(&empty_bar[s], 1);       // one tcgen05.commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);       // one tcgen05.commit
      mbar_init(&tmem_empty[b], 4 * kCta);   // one arrival per epilogue warp of every CTA of the pair
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kCta>(tmem_ptr, kTmemCols);
  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int T = p.tiles_per_strip;

  if (warp == 0) {
    // ===================================================== TMA producer (one lane per CTA)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // in a pair every CTA's bytes are counted on the LEADER's full barrier
      uint32_t full_addr0 = smem_u32(&full_bar[0]);
      if constexpr (kCta == 2) full_addr0 = mapa_shared_cluster(full_addr0, 0);
      for (int u = unit0; u < p.n_units; u += unit_stride) {
        const int m_tile = u % p.n_m;
        const int strip = u / p.n_m;
        const int t0 = strip * T;
        const int t1 = min(t0 + T, p.n_n);
        const int q_row = m_tile * kUnitRows + static_cast<int>(cta_rank) * kTileM;
        for (int t = t0; t < t1; ++t) {
          const int c_row = t * kTileN + static_cast<int>(cta_rank) * kBRows;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1, p.status, 1);
            uint8_t* sa = smem + L.stage_off + stage * stage_sz;
            const int kelem = kb * (kF8 ? kBlockKBytes : kBlockKBytes / 2);
            if constexpr (kCta == 1) {
              mbar_arrive_expect_tx(&full_bar[stage], stage_sz);
              tma_load_2d(sa, &tmap_q, &full_bar[stage], kelem, q_row, kEvictLast);
              tma_load_2d(sa + a_bytes, &tmap_c, &full_bar[stage], kelem, c_row, kEvictNormal);
            } else {
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_sz);
              const uint32_t bar = full_addr0 + stage * 8;
              tma_load_2d_2sm(sa, &tmap_q, bar, kelem, q_row, kEvictLast);
              tma_load_2d_2sm(sa + a_bytes, &tmap_c, bar, kelem, c_row, kEvictNormal);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (single thread of the leader CTA)
    if (lane == 0 && cta_rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      MMD_ST_DECL;
#ifdef MMD_STATS
      const long long st_begin = clock64();
#endif
      for (int u = unit0; u < p.n_units; u += unit_stride) {
        const int strip = u / p.n_m;
        const int t0 = strip * T;
        const int t1 = min(t0 + T, p.n_n);
        for (int t = t0; t < t1; ++t, ++it) {
          const uint32_t buf = it & 1;
          MMD_TRACE(it, 0);
          MMD_ST_T0();
          mbar_wait<kCta == 2>(&tmem_empty[buf], ((it >> 1) & 1) ^ 1, p.status, 2);
          MMD_ST_ACC(kStMmaEmpty);
          MMD_TRACE(it, 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * kTileN;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            MMD_ST_T0();
            mbar_wait(&full_bar[stage], phase, p.status, 3);
            MMD_ST_ACC(kStMmaFull);
            tc_fence_after();
            // (with a_rows = 32 the 128-row A descriptor also covers what follows the staged rows: those accumulator
            //  lanes belong to query rows >= Q, which no epilogue thread ever reads a candidate from)
            const uint32_t sa = smem_u32(smem + L.stage_off + stage * stage_sz);
            const uint64_t adesc = make_smem_desc_sw128(sa);
            const uint64_t bdesc = make_smem_desc_sw128(sa + a_bytes);
#pragma unroll
            for (int k = 0; k < kBlockKBytes / 32; ++k) {
              // advance 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
              umma<kCta, kF8>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            // frees the smem stage (in both CTAs of a pair) when these MMAs retire
            if constexpr (kCta == 2) umma_commit_2sm(&empty_bar[stage], 3); else umma_commit(&empty_bar[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          // accumulator complete -> epilogue warps (of both CTAs)
          if constexpr (kCta == 2) umma_commit_2sm(&tmem_full[buf], 3); else umma_commit(&tmem_full[buf]);
          MMD_TRACE(it, 2);
        }
      }
#ifdef MMD_STATS
      st_acc[kStMmaTotal] = clock64() - st_begin;
#endif
      MMD_ST_FLUSH();
    }
  } else {
    // ===================================================== epilogue warps (2..5)
    const int quarter = warp & 3;               // TMEM lane quarter this warp may read
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    // candidate buffers are indexed by the lane quarter (= 32-row group of the tile); with key_warps = 1 only quarter 0
    // (query rows 0..31) owns one -- the other warps' rows are all beyond Q and never touch theirs
    uint64_t* wkeys = reinterpret_cast<uint64_t*>(smem + L.keys_off) + static_cast<size_t>(quarter) * (CAP * 32);
    const uint32_t row_addr = smem_u32(wkeys) + static_cast<uint32_t>(lane) * 8u;   // slot 0 of this thread's own row
    const float kNegInf = __int_as_float(0xff800000);
    const float kPosInf = __int_as_float(0x7f800000);
    uint32_t it = 0;
    MMD_ST_DECL;
#ifdef MMD_STATS
    const long long st_begin = clock64();
    long long st_appends = 0;
#endif
    for (int u = unit0; u < p.n_units; u += unit_stride) {
      const int m_tile = u % p.n_m;
      const int strip = u / p.n_m;
      const int t0 = strip * T;
      const int t1 = min(t0 + T, p.n_n);
      const int64_t cta_row0 = static_cast<int64_t>(m_tile) * kUnitRows + static_cast<int64_t>(cta_rank) * kTileM;
      const int64_t qrow = cta_row0 + row_in_tile;
      const bool valid = qrow < p.Q;
      float thr = valid ? kNegInf : kPosInf;
      int cnt = 0;
      uint32_t seen = 0u;                       // largest shared bound of this row read or written by this thread
      MMD_ST_ADD(kStUnits, 1);
      for (int t = t0; t < t1; ++t, ++it) {
        MMD_ST_ADD(kStTiles, 1);
        MMD_ST_T0();
        const uint32_t buf = it & 1;
        // pick up what other CTAs (and earlier strips) have learnt about this row while the MMAs finish
        uint32_t shared_bound = 0u;
        if constexpr (!kDense) {
          if (valid) shared_bound = __ldcg(p.thr_global + qrow);
        }
        if (warp == 2 && lane == 0) MMD_TRACE(it, 4);
        mbar_wait(&tmem_full[buf], (it >> 1) & 1, p.status, 4);
        if (warp == 2 && lane == 0) MMD_TRACE(it, 5);
        MMD_ST_ACC(kStEpiWait);
        tc_fence_after();
        if constexpr (!kDense) {
          if (shared_bound != 0u) thr = fmaxf(thr, ordered_to_float(shared_bound));
          seen = max(seen, shared_bound);
        }
        const int64_t col_tile = static_cast<int64_t>(t) * kTileN;
        const bool edge = col_tile + kTileN > p.N;
        if constexpr (!kDense && CAP == 64) {
          // Cold start: a row about which nothing is known yet (thr = -inf) would push every score of its first
          // tile through the candidate buffer (one warp-wide sort per ~40 scores).  Instead, read the tile once
          // more: the K'-th largest of its 32 group maxima (8 columns each) is K' distinct scores of this row,
          // hence a lower bound of the row's K'-th best; it admits ~8 % of the tile.
          if (__any_sync(kFullMask, valid && thr == kNegInf)) {
            float gmx[32];
#pragma unroll
            for (int c = 0; c < kTileN / 32; ++c) {
              uint32_t raw[32];
              tmem_ld_32x32b_x32(tmem_base + lane_base + buf * kTileN + c * 32, raw);
              tmem_ld_wait();
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float mx = kNegInf;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float x = __uint_as_float(raw[8 * g + j]);
                  const bool in_range = !edge || (col_tile + c * 32 + 8 * g + j < p.N);
                  mx = in_range ? fmaxf(mx, x) : mx;
                }
                gmx[4 * c + g] = (mx == mx) ? mx : kNegInf;
              }
            }
            thread_sort_desc<32>(gmx);
            float kth = gmx[0];
#pragma unroll
            for (int i = 1; i < 32; ++i) kth = (i == p.kprime - 1) ? gmx[i] : kth;
            if (valid && thr == kNegInf && kth > kNegInf) {
              seen = publish_threshold(p, qrow, kth, seen);
              thr = ordered_to_float(float_to_ordered(kth) - 1u);     // admit scores equal to the bound
            }
          }
        }
        MMD_ST_ACC(kStEpiCold);
        if constexpr (kDense) {
#pragma unroll 1
          for (int c = 0; c < kTileN / 32; ++c) {
            uint32_t raw[32];
            tmem_ld_32x32b_x32(tmem_base + lane_base + buf * kTileN + c * 32, raw);
            tmem_ld_wait();
            const int64_t col0 = col_tile + c * 32;
            if (valid) {
              float* out = p.dense + qrow * p.ldd + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) out[j] = __uint_as_float(raw[j]) * p.out_scale;
            }
          }
        } else {
          // ---- pass 1 (filter): branch-free sweep over the tile, 32 columns per tcgen05.ld, two loads in flight: the maxima
          // of the tile's 64 groups of 4 columns against the row threshold -> a 64-bit mask per thread, OR-ed over the warp:
          // the groups that hold a candidate for SOME row.  The sweep runs at the TMEM read rate (~65 cycles per load); the
          // mask arithmetic hides under it.  (Columns beyond N are TMA zero fill; a false hit on them is sorted out in pass 2.)
          const uint32_t tile_addr = tmem_base + lane_base + buf * kTileN;
          uint32_t w0 = 0u, w1 = 0u;
          {
            uint32_t ra[32], rb[32];
            tmem_ld_32x32b_x32(tile_addr, ra);
#pragma unroll
            for (int c = 0; c < kTileN / 32; c += 2) {
              tmem_ld_wait();
              tmem_ld_32x32b_x32(tile_addr + (c + 1) * 32, rb);
              const uint32_t ma = group_mask(ra, thr) << ((c & 3) * 8);
              tmem_ld_wait();
              if (c + 2 < kTileN / 32) tmem_ld_32x32b_x32(tile_addr + (c + 2) * 32, ra);
              const uint32_t mb = group_mask(rb, thr) << (((c + 1) & 3) * 8);
              if (c < 4) w0 |= ma | mb; else w1 |= ma | mb;
            }
          }
          uint64_t wanted = (static_cast<uint64_t>(__reduce_or_sync(kFullMask, w1)) << 32) | __reduce_or_sync(kFullMask, w0);
          MMD_ST_ACC(kStEpiFilter);
          MMD_ST_ADD(kStHitChunks, wanted != 0ull ? 1 : 0);
          // ---- pass 2 (collect): only the wanted groups are appended from.  (Round 1 re-read whole 32-column chunks and
          // voted per 8-column group: ~900 cycles per marked chunk, two thirds of the epilogue's busy time in every shape.)
          const uint32_t col_lo = static_cast<uint32_t>(col_tile);                                  // (N < 2^31)
          const int cols_here = edge ? static_cast<int>(p.N - col_tile) : kTileN;
          auto visit = [&](int grp, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
            MMD_ST_ADD(kStGroups, 1);
            uint32_t need = __ballot_sync(kFullMask, cnt > CAP - 4);
            if (need) {
              // (long lists: rows that are nearly full come along, several rows share one pass of the selection)
              if constexpr (CAP > 64) need = __ballot_sync(kFullMask, cnt > p.early_level);
              MMD_ST_ACC(kStEpiCollect);
              MMD_ST_ADD(kStCompCalls, 1);
              MMD_ST_ADD(kStCompRows, __popc(need));
              const float before = thr;
              const RowState st = compact_rows<CAP>(wkeys, need, lane, p.kprime, p.keep_hi, cnt, thr);
              cnt = st.cnt;
              thr = st.thr;
              if (thr != before) seen = publish_threshold(p, qrow, thr, seen);
              MMD_ST_ACC(kStEpiCompact);
            }
            // branch-free appends: raw entry {~column, score bits} stored under a predicate
            const uint32_t ncol = ~(col_lo + static_cast<uint32_t>(grp * 4));
            const int left = cols_here - grp * 4;              // columns of this group that exist (>= 4 except at the corpus end)
            const uint32_t xs[4] = {x0, x1, x2, x3};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bool take = __uint_as_float(xs[j]) > thr && j < left;
              st_shared_v2_pred(row_addr + static_cast<uint32_t>(cnt) * 256u, ncol - j, xs[j], take);
              cnt += take ? 1 : 0;
#ifdef MMD_STATS
              st_appends += take ? 1 : 0;
#endif
            }
          };
          // 4 columns per tcgen05.ld, up to four loads in flight
#pragma unroll 1
          while (wanted != 0ull) {
            uint32_t v[4][4];
            int grp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              grp[i] = -1;
              if (wanted != 0ull) {
                grp[i] = __ffsll(static_cast<long long>(wanted)) - 1;
                wanted &= wanted - 1ull;
                tmem_ld_32x32b_x4(tile_addr + grp[i] * 4, v[i]);
              }
            }
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (grp[i] < 0) break;
              visit(grp[i], v[i][0], v[i][1], v[i][2], v[i][3]);
            }
          }
          MMD_ST_ACC(kStEpiCollect);
        }
        // every score of the tile has been looked at: hand the TMEM buffer back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCta == 2) mbar_arrive_cluster(&tmem_empty[buf], 0); else mbar_arrive(&tmem_empty[buf]);
        }
        if (warp == 2 && lane == 0) MMD_TRACE(it, 6);
      }
      if constexpr (!kDense) {
        // unit done: emit the sorted K-list of every valid row of this warp
        MMD_ST_T0();
        const int64_t row0 = cta_row0 + quarter * 32;
        uint64_t* out_rows = p.partial + (row0 * p.n_strips + strip) * p.kprime;
        const float before = thr;
        const LevelPub pub{p.lvl, p.Q, row0, strip, p.levels};
        const int64_t out_stride = static_cast<int64_t>(p.n_strips) * p.kprime;
        const RowState st = final_lists<CAP>(wkeys, lane, p.kprime, valid, cnt, thr, out_rows + lane * out_stride, out_stride, pub);
        cnt = st.cnt;
        thr = st.thr;
        if (valid && thr != before) seen = publish_threshold(p, qrow, thr, seen);
        if (p.lvl != nullptr && valid) {
          // merged bound: fold what the finished strips have left in the quantile slots (this unit's included)
          // (all loads first: they are independent, and a warp alone on its scheduler pays every L2 round trip in full)
          uint32_t slot[level_slots(kMaxLevels)];
          const int n_slots = level_slots(p.levels);
#pragma unroll
          for (int s = 0; s < level_slots(kMaxLevels); ++s)
            slot[s] = s < n_slots ? __ldcg(p.lvl + static_cast<int64_t>(s) * p.Q + qrow) : 0u;
          uint32_t best = 0u;
#pragma unroll
          for (int j = 1; j <= kMaxLevels; ++j) {
            uint32_t m = 0xffffffffu;
#pragma unroll
            for (int s = 0; s < (1 << j); ++s) m = min(m, slot[level_slot0(j) + s]);
            best = max(best, m);                       // an empty slot (0) voids its level; so does a level that is off
          }
          if (best != 0u && ordered_to_float(best) > thr) seen = publish_threshold(p, qrow, ordered_to_float(best), seen);
        }
        MMD_ST_ACC(kStEpiFinal);
      }
    }
#ifdef MMD_STATS
    st_acc[kStEpiTotal] = clock64() - st_begin;
    st_acc[kStAppends] = __reduce_add_sync(kFullMask, static_cast<int>(st_appends));
#endif
    MMD_ST_FLUSH();
  }

  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<kCta>(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------- host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

// Tensor map over prepared rows: dims {kdim, rows}, box {128 B of K, box_rows}, 128B swizzle, zero OOB fill.
int make_rows_tensor_map(CUtensorMap* tm, const void* base, int op_dtype, int64_t rows, const PreparedLayout& lay,
                         int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) {
    set_last_error("cuTensorMapEncodeTiled entry point not available from the driver");
    return MMD_ERR_CUDA;
  }
  CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  if (op_dtype == MMD_OP_F16) dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  if (op_dtype == MMD_OP_E4M3) dt = CU_TENSOR_MAP_DATA_TYPE_UINT8;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(lay.kdim), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(lay.row_bytes)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockKBytes / lay.elem_bytes), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld kdim=%lld pitch=%lld)", (int)r,
                   (long long)rows, (long long)lay.kdim, (long long)lay.row_bytes);
    return MMD_ERR_CUDA;
  }
  return MMD_OK;
}

// Candidate-buffer capacity per query row.  A row is compacted back to kprime entries when fewer than 8 free
// slots remain, so capacity - kprime is what amortises a compaction.
int cap_for_k(int kprime) {
  if (kprime <= 32) return 64;
  if (kprime <= 120) return 128;
  return 0;
}

struct Schedule {
  int n_m, n_n, T, S, n_units, grid;
};

// Pick the strip length.  Units are dealt round-robin to `sms` persistent CTAs; the cost of a schedule is
// the busiest CTA's tile count plus a per-unit restart charge (list warm-up + final sort + the extra
// strip the merge has to fold), expressed in tiles.  Fewer strips win ties.
Schedule plan_schedule(int64_t Q, int64_t N, int kprime, int sms, int kblocks, int cta) {
  Schedule s{};
  sms /= cta;                                   // schedulable entities: CTAs, or CTA pairs
  if (sms < 1) sms = 1;
  s.n_m = static_cast<int>(ceil_div(Q, kTileM * cta));
  s.n_n = static_cast<int>(ceil_div(N, kTileN));
  const int max_parts = kprime > 0 ? 4096 / kprime : 4096;
  int s_max = s.n_n < max_parts ? s.n_n : max_parts;
  if (s_max > 1024) s_max = 1024;
  if (s_max < 1) s_max = 1;
  // restart charge: ~8k cycles against kblocks * 512 cycles of MMA per tile, at least a quarter tile
  double restart = 8000.0 / (static_cast<double>(kblocks > 0 ? kblocks : 1) * 512.0);
  if (restart < 0.25) restart = 0.25;
  static const double forced_restart = [] { const char* e = getenv("MMD_RESTART_TILES"); return e ? atof(e) : 0.0; }();   // tuning knob
  if (forced_restart > 0.0) restart = forced_restart;
  double best_cost = -1.0;
  int best_S = 1, best_T = s.n_n;
  int last_T = -1;
  for (int S = 1; S <= s_max; ++S) {
    const int T = static_cast<int>(ceil_div(s.n_n, S));
    if (T == last_T) continue;
    last_T = T;
    const int S_eff = static_cast<int>(ceil_div(s.n_n, T));
    const int64_t units = static_cast<int64_t>(s.n_m) * S_eff;
    const int G = static_cast<int>(units < sms ? units : sms);
    const int T_last = s.n_n - (S_eff - 1) * T;
    const int64_t lo = units - s.n_m;   // first unit of the last (possibly shorter) strip
    double worst = 0.0;
    for (int c = 0; c < G; ++c) {
      const int64_t cnt = (units - 1 - c) / G + 1;
      const int64_t first = lo + ((c - lo % G) % G + G) % G;
      const int64_t in_last = first < units ? (units - 1 - first) / G + 1 : 0;
      const double cost = static_cast<double>((cnt - in_last) * T + in_last * T_last) + restart * static_cast<double>(cnt);
      if (cost > worst) worst = cost;
    }
    if (best_cost < 0.0 || worst < best_cost * 0.999) {
      best_cost = worst;
      best_S = S_eff;
      best_T = T;
    }
  }
  s.S = best_S;
  s.T = best_T;
  s.n_units = s.n_m * s.S;
  s.grid = (s.n_units < sms ? s.n_units : sms) * cta;
  return s;
}

// Two-CTA pairs pay off as soon as there is more than one 128-query tile; a single tile (the reference's
// one-query-per-call pattern) keeps every SM busy on its own strip instead.
int cta_mode_for(int64_t Q) {
  static const int forced = [] { const char* e = getenv("MMD_CTA_MODE"); return e ? atoi(e) : 0; }();   // tuning knob
  if (forced == 1 || forced == 2) return forced;
  return Q > kTileM ? 2 : 1;
}

}  // namespace

// The status word lives in mapped pinned host memory so that the host can still read which wait
// timed out after the watchdog trapped the context.
static DeviceStatus* g_status_host = nullptr;
DeviceStatus* device_status_word() {
  static DeviceStatus* dev_ptr = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  if (dev_ptr == nullptr) {
    DeviceStatus* h = nullptr;
    if (cudaHostAlloc(&h, sizeof(DeviceStatus), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
    h->code = 0; h->site = 0; h->block = 0; h->extra = 0;
    DeviceStatus* d = nullptr;
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) return nullptr;
    g_status_host = h;
    dev_ptr = d;
  }
  return dev_ptr;
}
const DeviceStatus* host_status_word() { return g_status_host; }

namespace {

// Optional event bracketing of the fused launches (mmd_profile_enable / mmd_profile_collect).
struct Profile {
  std::mutex mu;
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;         // eager launches: consumed by collect
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> graph_events;   // owned by captured graphs: kept, re-read every collect
} g_profile;
constexpr size_t kMaxProfiled = 512;

template <int CAP, bool kF8, bool kDense, int kCta>
int launch_fused(const CUtensorMap& tq, const CUtensorMap& tc, const FusedParams& p, uint32_t idesc, int grid,
                 cudaStream_t stream) {
  const SmemLayout L = smem_layout(p.stages, kDense ? 0 : CAP, kCta, p.a_rows, p.key_warps);
  auto kern = fused_score_topk_kernel<CAP, kF8, kDense, kCta>;
  // the attribute is per device (context): one bit per device ordinal, one mask per template instantiation
  static std::atomic<unsigned long long> attr_done{0ull};
  int dev = 0;
  MMD_CUDA_OK(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if ((attr_done.load(std::memory_order_acquire) & bit) == 0ull) {
    MMD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr_done.fetch_or(bit, std::memory_order_release);
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  bool profiled = false, capturing = false;
  {
    std::lock_guard<std::mutex> g(g_profile.mu);
    if (g_profile.on && !kDense && g_profile.events.size() + g_profile.graph_events.size() < kMaxProfiled) {
      if (cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) profiled = true;
    }
  }
  if (profiled) {
    // Under stream capture the bracket becomes two EXTERNAL event-record nodes of the graph: every replay re-records
    // them, so after a replay the pair holds that replay's fused-kernel duration.
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    capturing = cudaStreamIsCapturing(stream, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive;
    if (capturing) cudaEventRecordWithFlags(e0, stream, cudaEventRecordExternal);
    else cudaEventRecord(e0, stream);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L.total;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tq, tc, p, idesc);
  if (profiled) {
    if (capturing) cudaEventRecordWithFlags(e1, stream, cudaEventRecordExternal);
    else cudaEventRecord(e1, stream);
    std::lock_guard<std::mutex> g(g_profile.mu);
    (capturing ? g_profile.graph_events : g_profile.events).emplace_back(e0, e1);
  }
  count_launch();
  MMD_CUDA_OK(le);
  MMD_CUDA_OK(cudaGetLastError());
  return MMD_OK;
}

// op format x pair mode dispatch for one CAP
template <int CAP, bool kDense>
int launch_any(bool f8, int cta, const CUtensorMap& tq, const CUtensorMap& tc, const FusedParams& p, uint32_t idesc,
               int grid, cudaStream_t stream) {
  if (cta == 2) {
    return f8 ? launch_fused<CAP, true, kDense, 2>(tq, tc, p, idesc, grid, stream)
              : launch_fused<CAP, false, kDense, 2>(tq, tc, p, idesc, grid, stream);
  }
  return f8 ? launch_fused<CAP, true, kDense, 1>(tq, tc, p, idesc, grid, stream)
            : launch_fused<CAP, false, kDense, 1>(tq, tc, p, idesc, grid, stream);
}

int stages_for(int cap, int cta, int a_rows = kTileM, int key_warps = 4) {
  static const int forced = [] { const char* e = getenv("MMD_STAGES"); return e ? atoi(e) : 0; }();     // tuning knob
  if (forced >= 2 && forced <= kMaxStages &&
      smem_layout(forced, cap, cta, a_rows, key_warps).total <= static_cast<uint32_t>(kMaxSmem)) return forced;
  for (int st = kMaxStages; st >= 2; --st)
    if (smem_layout(st, cap, cta, a_rows, key_warps).total <= static_cast<uint32_t>(kMaxSmem)) return st;
  return 0;
}

// A batch of fewer than 128 queries (one 1-CTA tile whose upper rows do not exist) stages only the 32-row groups that
// exist and keeps candidate buffers only for those; tuning knob MMD_SMALL_BATCH=0 switches the variant off.
// Measured on a 1 M x 768 bf16 corpus, Q = 1: 4.5 -> 6.7 TB/s of corpus stream (0.341 -> 0.229 ms).
int staged_query_rows(int64_t Q, int cta) {
  static const int enabled = [] { const char* e = getenv("MMD_SMALL_BATCH"); return e ? atoi(e) : 1; }();
  if (enabled == 0 || cta != 1 || Q >= kTileM) return kTileM;
  return static_cast<int>(round_up(Q, 32));
}

int sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();   // no device (CPU-only host): plan for a B200's 148 SMs
    sms = 148;
  }
  return sms;
}

uint32_t idesc_for(int op_dtype, int cta) {
  // kind::f16: 0 = f16, 1 = bf16 ; kind::f8f6f4: 0 = e4m3.  A pair's MMA spans 256 query rows.
  const uint32_t fmt = (op_dtype == MMD_OP_BF16 || op_dtype == MMD_OP_BF16X3) ? 1u : 0u;
  return make_idesc(fmt, kTileM * cta, kTileN);
}

}  // namespace
}  // namespace mmd

extern "C" int mmd_topk_max_k(void) { return 120; }

extern "C" int mmd_profile_enable(int on) {
  std::lock_guard<std::mutex> g(mmd::g_profile.mu);
  mmd::g_profile.on = on != 0;
  // switching off forgets the brackets that live inside captured graphs (the events stay alive: the graphs own nodes
  // that reference them)
  if (!mmd::g_profile.on) mmd::g_profile.graph_events.clear();
  return MMD_OK;
}

extern "C" int mmd_profile_collect(float* ms_host, int cap) {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev, gev;
  {
    std::lock_guard<std::mutex> g(mmd::g_profile.mu);
    ev.swap(mmd::g_profile.events);
    gev = mmd::g_profile.graph_events;
  }
  int n = 0;
  for (auto& pr : ev) {
    float ms = 0.0f;
    if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess &&
        ms_host != nullptr && n < cap) {
      ms_host[n++] = ms;
    }
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  // brackets living inside captured graphs: the duration of each graph's most recent replay (if it has run)
  for (auto& pr : gev) {
    float ms = 0.0f;
    if (cudaEventQuery(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess &&
        ms_host != nullptr && n < cap) {
      ms_host[n++] = ms;
    }
  }
  cudaGetLastError();   // a never-replayed graph's events are "not recorded": not an error of the library
  return n;
}

extern "C" size_t mmd_topk_workspace_bytes(int64_t Q, int64_t N, int dim, int op_dtype, int k) {
  using namespace mmd;
  if (Q <= 0 || N <= 0 || k <= 0 || cap_for_k(k) == 0) return 0;
  // exactly what mmd_topk_scores will carve up: one K-list per (query, strip) of the planned schedule
  PreparedLayout lay;
  if (!prepared_layout(op_dtype, dim, &lay)) return 0;
  const Schedule sch = plan_schedule(Q, N, k, sm_count(), static_cast<int>(ceil_div(lay.row_bytes, kBlockKBytes)), cta_mode_for(Q));
  return static_cast<size_t>(Q) * sch.S * k * sizeof(uint64_t) + static_cast<size_t>(Q) * sizeof(uint32_t) * (1 + level_slots(kMaxLevels)) + 256;
}

namespace mmd {
namespace {
int topk_scores_impl(const void* q_prep, const void* c_prep, int op_dtype, int64_t Q, int64_t N, int dim, int k,
                     int64_t idx_offset, float* out_scores, int32_t* out_idx, void* workspace, size_t workspace_bytes,
                     uint32_t* thr_local, void* const* thr_all_host, int n_thr, const PairOut* po, void* stream,
                     const PeerArrive* arrive = nullptr, void* merge_stream = nullptr) {
  MMD_REQUIRE(Q >= 0 && N >= 0 && dim > 0 && k > 0, "mmd_topk_scores: Q=%lld N=%lld dim=%d k=%d", (long long)Q,
              (long long)N, dim, k);
  MMD_REQUIRE(N < (1ll << 31) && Q < (1ll << 31), "mmd_topk_scores: Q and N must be < 2^31");
  MMD_REQUIRE(idx_offset >= 0 && idx_offset + N < (1ll << 31), "mmd_topk_scores: idx_offset + N must be < 2^31");
  if (Q == 0) return MMD_OK;
  MMD_REQUIRE(out_scores != nullptr && out_idx != nullptr, "mmd_topk_scores: null output");
  int rc = mmd_device_check();
  if (rc != MMD_OK) return rc;
  auto st = static_cast<cudaStream_t>(stream);
  // The strip merge may run on a second stream (behind an event recorded after the contraction), so that the next
  // contraction on `stream` does not wait for it.
  auto merge_on = [&]() -> cudaStream_t {
    auto ms = static_cast<cudaStream_t>(merge_stream);
    if (ms == nullptr || ms == st) return st;
    cudaEvent_t ev = nullptr;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return st;
    cudaEventRecord(ev, st);
    cudaStreamWaitEvent(ms, ev, 0);
    cudaEventDestroy(ev);                 // released once the wait has been satisfied
    return ms;
  };
  const int cap = cap_for_k(k);
  MMD_REQUIRE(cap != 0, "mmd_topk_scores: k=%d exceeds the fused selection limit %d", k, mmd_topk_max_k());
  if (N == 0) {
    // empty corpus: every slot is (-inf, -1)
    return merge_partial_keys(nullptr, 0, Q, k, k, 1.0f, idx_offset, out_scores, out_idx, merge_on(), po, arrive);
  }
  MMD_REQUIRE(q_prep != nullptr && c_prep != nullptr, "mmd_topk_scores: null operand");
  MMD_REQUIRE(reinterpret_cast<uintptr_t>(q_prep) % 16 == 0 && reinterpret_cast<uintptr_t>(c_prep) % 16 == 0,
              "mmd_topk_scores: operands must be 16-byte aligned");
  PreparedLayout lay;
  MMD_REQUIRE(prepared_layout(op_dtype, dim, &lay), "mmd_topk_scores: bad op_dtype %d", op_dtype);

  const int sms = sm_count();
  const int cta = cta_mode_for(Q);
  const Schedule sch = plan_schedule(Q, N, k, sms, static_cast<int>(ceil_div(lay.row_bytes, kBlockKBytes)), cta);
  const size_t keys_bytes = static_cast<size_t>(Q) * sch.S * k * sizeof(uint64_t);
  const size_t need = keys_bytes + static_cast<size_t>(Q) * sizeof(uint32_t) * (1 + level_slots(kMaxLevels));
  if (workspace == nullptr || workspace_bytes < need) {
    set_last_error("mmd_topk_scores: workspace %zu < %zu bytes", workspace_bytes, need);
    return MMD_ERR_WORKSPACE;
  }
  MMD_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 8 == 0, "mmd_topk_scores: workspace must be 8-byte aligned");

  CUtensorMap tq, tc;
  const int a_rows = staged_query_rows(Q, cta);
  const int key_warps = a_rows / 32;
  rc = make_rows_tensor_map(&tq, q_prep, op_dtype, Q, lay, a_rows);
  if (rc != MMD_OK) return rc;
  rc = make_rows_tensor_map(&tc, c_prep, op_dtype, N, lay, kTileN / cta);
  if (rc != MMD_OK) return rc;

  FusedParams p{};
  p.Q = Q; p.N = N;
  p.kblocks = static_cast<int>(ceil_div(lay.row_bytes, kBlockKBytes));
  p.n_m = sch.n_m; p.n_n = sch.n_n; p.tiles_per_strip = sch.T; p.n_strips = sch.S; p.n_units = sch.n_units;
  p.kprime = k;
  p.a_rows = a_rows; p.key_warps = key_warps;
  p.stages = stages_for(cap, cta, a_rows, key_warps);
  p.partial = static_cast<uint64_t*>(workspace);
  if (thr_local == nullptr) {
    p.thr_global = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(workspace) + keys_bytes);
    MMD_CUDA_OK(cudaMemsetAsync(p.thr_global, 0, static_cast<size_t>(Q) * sizeof(uint32_t), st));
    p.n_thr_peers = 0;
  } else {
    // caller-owned threshold arrays shared by all ranks of a row-sharded corpus (zeroed by the caller before the step)
    p.thr_global = thr_local;
    p.n_thr_peers = n_thr;
    for (int i = 0; i < n_thr; ++i) p.thr_peer[i] = static_cast<uint32_t*>(thr_all_host[i]);
  }
  {
    // long lists: selection window and early inclusion (see select_batch); merged bounds: as many levels as strips fill
    const int slack = cap - 4 - k;                         // appends a row can take between two compactions (4 per visit)
    p.keep_hi = k + (slack / 8 > 1 ? slack / 8 : 1);
    if (p.keep_hi > cap - 4) p.keep_hi = cap - 4;          // (no slack at all: the selection must be exact)
    // (rows that are merely nearly full are NOT taken along: a compaction call of many rows is a burst longer than a tile's
    //  worth of MMA time, and the issuer then waits for the accumulator -- measured: 8 rows per call = 15.7 k cycles)
    static const int early = [] { const char* e = getenv("MMD_EARLY"); return e ? atoi(e) : 0; }();   // tuning knob: slack/early rows come along
    p.early_level = cap - 4 - (early > 0 ? slack / early : 0);
    if (p.early_level < p.keep_hi) p.early_level = p.keep_hi;
    static const int max_levels = [] { const char* e = getenv("MMD_LEVELS"); return e ? atoi(e) : kMaxLevels; }();   // tuning knob (0 = off)
    int levels = 0;
    while (levels < kMaxLevels && levels < max_levels && (2 << levels) <= sch.S) ++levels;
    p.levels = levels;
    p.lvl = nullptr;
    if (levels > 0) {
      p.lvl = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(workspace) + keys_bytes) + Q;
      MMD_CUDA_OK(cudaMemsetAsync(p.lvl, 0, static_cast<size_t>(Q) * sizeof(uint32_t) * level_slots(levels), st));
    }
  }
  p.dense = nullptr; p.ldd = 0; p.out_scale = 1.0f;
  p.status = device_status_word();
  p.stats = nullptr;
#ifdef MMD_STATS
  {
    static unsigned long long* dstats = nullptr;
    static int calls = 0;
    if (dstats == nullptr) { cudaMalloc(&dstats, 2048 * sizeof(unsigned long long)); cudaMemset(dstats, 0, 2048 * 8); }
    if (++calls == 4) {          // dump the timeline and the aggregates of the previous (warm) launch once
      static unsigned long long h[2048];
      cudaMemcpy(h, dstats, sizeof(h), cudaMemcpyDeviceToHost);
      {
        const unsigned long long* a = h + 1024;
        const double n_mma = static_cast<double>(sch.grid / cta), n_epi = static_cast<double>(sch.grid) * 4.0;
        fprintf(stderr, "[stats] grid %d cta %d  S %d T %d units %d levels %d  k' %d\n", sch.grid, cta, sch.S, sch.T, sch.n_units, p.levels, k);
        fprintf(stderr, "[stats] MMA issuer (avg cycles per CTA): total %.0f  waiting for accumulator %.0f (%.1f%%)  waiting for operands %.0f (%.1f%%)\n",
                a[kStMmaTotal] / n_mma, a[kStMmaEmpty] / n_mma, 100.0 * a[kStMmaEmpty] / (a[kStMmaTotal] + 1.0), a[kStMmaFull] / n_mma,
                100.0 * a[kStMmaFull] / (a[kStMmaTotal] + 1.0));
        fprintf(stderr, "[stats] epilogue warp (avg cycles): total %.0f  wait %.0f  cold %.0f  filter %.0f  collect %.0f  compact %.0f  unit-end %.0f\n",
                a[kStEpiTotal] / n_epi, a[kStEpiWait] / n_epi, a[kStEpiCold] / n_epi, a[kStEpiFilter] / n_epi, a[kStEpiCollect] / n_epi,
                a[kStEpiCompact] / n_epi, a[kStEpiFinal] / n_epi);
        const double tiles = a[kStTiles] + 1e-9;
        fprintf(stderr, "[stats] per warp-tile: busy %.0f cycles (filter %.0f collect %.0f compact %.0f)  hit chunks %.2f  group visits %.2f  appends %.2f  "
                "compaction calls %.3f rows %.3f;  cycles per compaction call %.0f, per compacted row %.0f;  unit-end %.0f cycles per warp-unit\n",
                (a[kStEpiCold] + a[kStEpiFilter] + a[kStEpiCollect] + a[kStEpiCompact]) / tiles, a[kStEpiFilter] / tiles, a[kStEpiCollect] / tiles,
                a[kStEpiCompact] / tiles, a[kStHitChunks] / tiles, a[kStGroups] / tiles, a[kStAppends] / tiles, a[kStCompCalls] / tiles,
                a[kStCompRows] / tiles, a[kStEpiCompact] / (a[kStCompCalls] + 1e-9), a[kStEpiCompact] / (a[kStCompRows] + 1e-9),
                a[kStEpiFinal] / (a[kStUnits] + 1e-9));
      }
      const unsigned long long t0 = h[0];
      for (int t = 0; t < 128 && h[t * 8 + 2] != 0; ++t)
        fprintf(stderr, "[trace] tile %3d  mma: start %8lld waited %6lld issue-done %8lld | epi: wait-start %8lld got %8lld done %8lld\n", t,
                (long long)(h[t * 8] - t0), (long long)(h[t * 8 + 1] - h[t * 8]), (long long)(h[t * 8 + 2] - t0),
                (long long)(h[t * 8 + 4] - t0), (long long)(h[t * 8 + 5] - t0), (long long)(h[t * 8 + 6] - t0));
    }
    cudaMemsetAsync(dstats + 1024, 0, 1024 * sizeof(unsigned long long), st);      // aggregates are per launch
    p.stats = dstats;
  }
#endif
  const uint32_t idesc = idesc_for(op_dtype, cta);
  const bool f8 = op_dtype == MMD_OP_E4M3;

  if (cap == 64) rc = launch_any<64, false>(f8, cta, tq, tc, p, idesc, sch.grid, st);
  else rc = launch_any<128, false>(f8, cta, tq, tc, p, idesc, sch.grid, st);
  if (rc != MMD_OK) return rc;

  const float scale = f8 ? (1.0f / 65536.0f) : 1.0f;
  return merge_partial_keys(p.partial, sch.S, Q, k, k, scale, idx_offset, out_scores, out_idx, merge_on(), po, arrive);
}
}  // namespace
}  // namespace mmd

extern "C" int mmd_topk_scores(const void* q_prep, const void* c_prep, int op_dtype, int64_t Q, int64_t N, int dim,
                               int k, int64_t idx_offset, float* out_scores, int32_t* out_idx, void* workspace,
                               size_t workspace_bytes, void* stream) {
  return mmd::topk_scores_impl(q_prep, c_prep, op_dtype, Q, N, dim, k, idx_offset, out_scores, out_idx, workspace,
                               workspace_bytes, nullptr, nullptr, 0, nullptr, stream);
}

extern "C" int mmd_topk_scores_shared(const void* q_prep, const void* c_prep, int op_dtype, int64_t Q, int64_t N, int dim,
                                      int k, int64_t idx_offset, float* out_scores, int32_t* out_idx, void* workspace,
                                      size_t workspace_bytes, uint32_t* thr_local, void* const* thr_all_host, int n_thr,
                                      void* const* pair_dst_host, int n_pair_dst, int64_t pair_offset, void* stream) {
  using namespace mmd;
  MMD_REQUIRE(n_pair_dst >= 0 && n_pair_dst <= 16 && (n_pair_dst == 0 || pair_dst_host != nullptr) && pair_offset >= 0,
              "mmd_topk_scores_shared: n_pair_dst=%d (0..16)", n_pair_dst);
  PairOut po{};
  po.n = n_pair_dst;
  po.offset = pair_offset;
  for (int i = 0; i < n_pair_dst; ++i) {
    MMD_REQUIRE(pair_dst_host[i] != nullptr && reinterpret_cast<uintptr_t>(pair_dst_host[i]) % 8 == 0,
                "mmd_topk_scores_shared: pair destination %d is null or not 8-byte aligned", i);
    po.dst[i] = static_cast<int2*>(pair_dst_host[i]);
  }
  MMD_REQUIRE(thr_local != nullptr && thr_all_host != nullptr && n_thr >= 1 && n_thr <= 8,
              "mmd_topk_scores_shared: thr_local / thr_all_host null or n_thr=%d not in 1..8", n_thr);
  for (int i = 0; i < n_thr; ++i)
    MMD_REQUIRE(thr_all_host[i] != nullptr && reinterpret_cast<uintptr_t>(thr_all_host[i]) % 4 == 0,
                "mmd_topk_scores_shared: threshold array %d is null or misaligned", i);
  return topk_scores_impl(q_prep, c_prep, op_dtype, Q, N, dim, k, idx_offset, out_scores, out_idx, workspace,
                          workspace_bytes, thr_local, thr_all_host, n_thr, n_pair_dst > 0 ? &po : nullptr, stream);
}

extern "C" int mmd_sharded_candidates(const void* q_prep, const void* c_prep, int op_dtype, int64_t Q, int64_t N, int dim, int k,
                                      int64_t idx_offset, float* out_scores, int32_t* out_idx, void* workspace,
                                      size_t workspace_bytes, uint32_t* thr_local, void* const* thr_all_host, int n_thr,
                                      int reset_thr, void* const* pair_dst_host, int n_pair_dst, int64_t pair_offset, int pair_width,
                                      void* const* arrive_flags_host, int n_arrive, uint32_t* sync_state, void* stream,
                                      void* merge_stream) {
  using namespace mmd;
  const char* who = "mmd_sharded_candidates";
  MMD_REQUIRE(n_pair_dst >= 1 && n_pair_dst <= 16 && pair_dst_host != nullptr && pair_offset >= 0 && pair_width >= k,
              "%s: n_pair_dst=%d (1..16) pair_width=%d (>= k=%d)", who, n_pair_dst, pair_width, k);
  PairOut po{};
  po.n = n_pair_dst;
  po.offset = pair_offset;
  po.width = pair_width;
  for (int i = 0; i < n_pair_dst; ++i) {
    MMD_REQUIRE(pair_dst_host[i] != nullptr && reinterpret_cast<uintptr_t>(pair_dst_host[i]) % 8 == 0,
                "%s: pair destination %d is null or not 8-byte aligned", who, i);
    po.dst[i] = static_cast<int2*>(pair_dst_host[i]);
  }
  // n_thr = 0: thresholds private to this launch (kept in the workspace); else shared with the peers as in mmd_topk_scores_shared
  MMD_REQUIRE(n_thr >= 0 && n_thr <= 8 && (n_thr == 0 || (thr_local != nullptr && thr_all_host != nullptr)),
              "%s: thr_local / thr_all_host null or n_thr=%d not in 0..8", who, n_thr);
  for (int i = 0; i < n_thr; ++i)
    MMD_REQUIRE(thr_all_host[i] != nullptr && reinterpret_cast<uintptr_t>(thr_all_host[i]) % 4 == 0,
                "%s: threshold array %d is null or misaligned", who, i);
  PeerArrive arrive{};
  int rc = fill_arrive(&arrive, arrive_flags_host, n_arrive, sync_state, who);
  if (rc != MMD_OK) return rc;
  if (n_thr > 0 && reset_thr != 0 && Q > 0)
    MMD_CUDA_OK(cudaMemsetAsync(thr_local, 0, static_cast<size_t>(Q) * sizeof(uint32_t), static_cast<cudaStream_t>(stream)));
  return topk_scores_impl(q_prep, c_prep, op_dtype, Q, N, dim, k, idx_offset, out_scores, out_idx, workspace, workspace_bytes,
                          n_thr > 0 ? thr_local : nullptr, thr_all_host, n_thr, &po, stream, n_arrive > 0 ? &arrive : nullptr,
                          merge_stream);
}

extern "C" int mmd_zero_u32(uint32_t* dst, int64_t n, void* stream) {
  MMD_REQUIRE(n >= 0 && (n == 0 || dst != nullptr), "mmd_zero_u32: n=%lld with a null pointer", (long long)n);
  if (n == 0) return MMD_OK;
  MMD_CUDA_OK(cudaMemsetAsync(dst, 0, static_cast<size_t>(n) * sizeof(uint32_t), static_cast<cudaStream_t>(stream)));
  return MMD_OK;
}

extern "C" int mmd_scores_dense(const void* q_prep, const void* c_prep, int op_dtype, int64_t Q, int64_t N, int dim,
                                float* out_scores, int64_t ld_scores, void* stream) {
  using namespace mmd;
  MMD_REQUIRE(Q >= 0 && N >= 0 && dim > 0, "mmd_scores_dense: Q=%lld N=%lld dim=%d", (long long)Q, (long long)N, dim);
  MMD_REQUIRE(N < (1ll << 31) && Q < (1ll << 31), "mmd_scores_dense: Q and N must be < 2^31");
  if (Q == 0 || N == 0) return MMD_OK;
  MMD_REQUIRE(q_prep != nullptr && c_prep != nullptr && out_scores != nullptr, "mmd_scores_dense: null buffer");
  MMD_REQUIRE(ld_scores >= N, "mmd_scores_dense: ld_scores %lld < N %lld", (long long)ld_scores, (long long)N);
  int rc = mmd_device_check();
  if (rc != MMD_OK) return rc;
  auto st = static_cast<cudaStream_t>(stream);
  PreparedLayout lay;
  MMD_REQUIRE(prepared_layout(op_dtype, dim, &lay), "mmd_scores_dense: bad op_dtype %d", op_dtype);
  const int sms = sm_count();
  const int cta = cta_mode_for(Q);
  Schedule sch{};
  sch.n_m = static_cast<int>(ceil_div(Q, kTileM * cta));
  sch.n_n = static_cast<int>(ceil_div(N, kTileN));
  sch.T = 1; sch.S = sch.n_n; sch.n_units = sch.n_m * sch.S;
  sch.grid = (sch.n_units < sms / cta ? sch.n_units : sms / cta) * cta;

  CUtensorMap tq, tc;
  rc = make_rows_tensor_map(&tq, q_prep, op_dtype, Q, lay, kTileM);
  if (rc != MMD_OK) return rc;
  rc = make_rows_tensor_map(&tc, c_prep, op_dtype, N, lay, kTileN / cta);
  if (rc != MMD_OK) return rc;

  FusedParams p{};
  p.Q = Q; p.N = N;
  p.kblocks = static_cast<int>(ceil_div(lay.row_bytes, kBlockKBytes));
  p.n_m = sch.n_m; p.n_n = sch.n_n; p.tiles_per_strip = 1; p.n_strips = sch.S; p.n_units = sch.n_units;
  p.kprime = 0;
  p.a_rows = kTileM; p.key_warps = 4;
  p.stages = stages_for(0, cta);
  if (p.stages > 6) p.stages = 6;
  p.partial = nullptr;
  p.thr_global = nullptr;
  p.stats = nullptr;
  p.dense = out_scores; p.ldd = ld_scores;
  const bool f8 = op_dtype == MMD_OP_E4M3;
  p.out_scale = f8 ? (1.0f / 65536.0f) : 1.0f;
  p.status = device_status_word();
  const uint32_t idesc = idesc_for(op_dtype, cta);
  return launch_any<64, true>(f8, cta, tq, tc, p, idesc, sch.grid, st);
}
