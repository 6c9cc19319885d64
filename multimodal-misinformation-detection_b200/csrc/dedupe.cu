// Distinct-score filter of a ranked hit list, on the device.
//
// The reference walks its descending (key, score) list and keeps the first entry of every distinct score until
// top_k entries are kept -- src/evidence/im2im_retrieval.py:94-104, src/evidence/text2text_retrieval.py:105-118 --
// and, in the evaluation scripts, always keeps the gold evidence as well (src/evidence/experiment_image.py:41-50,
// src/evidence/experiment_text.py:79-87).  On a descending list "score not seen before" is "score differs from the
// previous entry's", so the walk is a flag + prefix sum: one warp per query.
#include "common.cuh"

namespace mmd {
namespace {

__global__ void __launch_bounds__(128)
dedupe_kernel(const float* __restrict__ scores, const int32_t* __restrict__ idx, const int32_t* __restrict__ gold, int64_t Q,
              int k_in, int top_k, float* __restrict__ out_s, int32_t* __restrict__ out_i, int32_t* __restrict__ out_n) {
  const int lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (q >= Q) return;
  const float* s = scores + q * k_in;
  const int32_t* ix = idx + q * k_in;
  const int32_t g = gold != nullptr ? gold[q] : -1;
  int kept = 0;
  for (int base = 0; base < k_in && kept < top_k; base += 32) {
    const int j = base + lane;
    bool keep = false;
    float sj = 0.0f;
    int32_t ij = -1;
    if (j < k_in) {
      sj = s[j];
      ij = ix[j];
      const bool fresh = j == 0 || sj != s[j - 1];
      keep = ij >= 0 && (fresh || (g >= 0 && ij == g));
    }
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    const int pos = kept + __popc(m & ((1u << lane) - 1u));
    if (keep && pos < top_k) {
      out_s[q * top_k + pos] = sj;
      out_i[q * top_k + pos] = ij;
    }
    kept += __popc(m);
  }
  if (kept > top_k) kept = top_k;
  for (int p = kept + lane; p < top_k; p += 32) {
    out_s[q * top_k + p] = __int_as_float(0xff800000);
    out_i[q * top_k + p] = -1;
  }
  if (lane == 0 && out_n != nullptr) out_n[q] = kept;
}

}  // namespace
}  // namespace mmd

extern "C" int mmd_dedupe_scores(const float* scores, const int32_t* idx, const int32_t* gold_idx, int64_t Q, int k_in,
                                 int top_k, float* out_scores, int32_t* out_idx, int32_t* out_count, void* stream) {
  using namespace mmd;
  MMD_REQUIRE(Q >= 0 && k_in > 0 && top_k > 0, "mmd_dedupe_scores: Q=%lld k_in=%d top_k=%d", (long long)Q, k_in, top_k);
  if (Q == 0) return MMD_OK;
  MMD_REQUIRE(scores != nullptr && idx != nullptr && out_scores != nullptr && out_idx != nullptr, "mmd_dedupe_scores: null buffer");
  int rc = mmd_device_check();
  if (rc != MMD_OK) return rc;
  dedupe_kernel<<<static_cast<unsigned>(ceil_div(Q, 4)), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      scores, idx, gold_idx, Q, k_in, top_k, out_scores, out_idx, out_count);
  count_launch();
  MMD_CUDA_OK(cudaGetLastError());
  return MMD_OK;
}
