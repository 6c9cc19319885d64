// Shared host/device helpers: ordered (score, row) keys, error plumbing, launch counting.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/mmd_retrieval.h"

namespace mmd {

// ---------------------------------------------------------------- ordered candidate keys
// A candidate is one 64-bit key: high word = score mapped to an order-preserving unsigned, low word
// = ~row.  Comparing keys as unsigned integers orders by (score descending, row ascending) when the
// LARGER key wins, which is the tie rule of the whole path.  Key 0 is the empty slot (sorts last).
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f + 0.0f);   // +0.0f folds -0.0 into +0.0
#else
  union { float f; uint32_t u; } c; c.f = f + 0.0f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return (static_cast<uint64_t>(float_to_ordered(score)) << 32) | static_cast<uint64_t>(~row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) {
  return ordered_to_float(static_cast<uint32_t>(k >> 32));
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return ~static_cast<uint32_t>(k); }

// Packed copies of a result list: {score bits, row} pairs stored to up to 16 buffers (every rank's gather buffer of a
// row-sharded corpus, peer-mapped) at pair offset `offset` + position.
struct PairOut {
  int2* dst[16];
  int n;
  int width;        // pairs per query in the destinations (>= the list length; the surplus is filled with empties); 0 = list length
  int64_t offset;
};

// ---------------------------------------------------------------- host-side error plumbing
void set_last_error(const char* fmt, ...);
void count_launch(int n = 1);

#define MMD_CUDA_OK(expr)                                                                            \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      ::mmd::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MMD_ERR_CUDA;                                                                           \
    }                                                                                                \
  } while (0)

#define MMD_REQUIRE(cond, ...)              \
  do {                                      \
    if (!(cond)) {                          \
      ::mmd::set_last_error(__VA_ARGS__);   \
      return MMD_ERR_ARG;                   \
    }                                       \
  } while (0)

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Layout of prepared rows (see mmd_prepared_layout).
struct PreparedLayout {
  int64_t dpad;       // dim padded so one limb is a multiple of 16 bytes
  int64_t kdim;       // contraction length in operand elements
  int64_t row_bytes;  // pitch
  int elem_bytes;
};
inline bool prepared_layout(int op_dtype, int dim, PreparedLayout* out) {
  if (dim <= 0) return false;
  PreparedLayout l{};
  switch (op_dtype) {
    case MMD_OP_BF16:
    case MMD_OP_F16:
      l.elem_bytes = 2; l.dpad = round_up(dim, 8); l.kdim = l.dpad; break;
    case MMD_OP_E4M3:
      l.elem_bytes = 1; l.dpad = round_up(dim, 16); l.kdim = l.dpad; break;
    case MMD_OP_BF16X3:
      l.elem_bytes = 2; l.dpad = round_up(dim, 8); l.kdim = 6 * l.dpad; break;
    default:
      return false;
  }
  l.row_bytes = l.kdim * l.elem_bytes;
  *out = l;
  return true;
}

}  // namespace mmd
