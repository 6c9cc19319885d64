// Device helpers shared by the exact re-score kernels (rescore.cu) and the fused exchange stage (exchange.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace mmd {

template <typename T>
__device__ __forceinline__ float cvt(T v);
template <>
__device__ __forceinline__ float cvt<float>(float v) { return v; }
template <>
__device__ __forceinline__ float cvt<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float cvt<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
struct Vec {
  static constexpr int kElems = 16 / sizeof(T);
};

template <typename TQ, typename TC>
__device__ __forceinline__ float warp_dot(const TQ* __restrict__ a, const TC* __restrict__ b, int dim, int lane,
                                          bool vec_ok) {
  float acc = 0.0f;
  if (vec_ok) {
    // 8 elements per lane per step; both rows 16-byte aligned, dim % 8 == 0
    for (int i = lane * 8; i < dim; i += 256) {
      float x[8], y[8];
      if constexpr (sizeof(TQ) == 4) {
        const float4 u = *reinterpret_cast<const float4*>(a + i), w = *reinterpret_cast<const float4*>(a + i + 4);
        x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w; x[4] = w.x; x[5] = w.y; x[6] = w.z; x[7] = w.w;
      } else {
        const uint4 u = *reinterpret_cast<const uint4*>(a + i);
        const TQ* p = reinterpret_cast<const TQ*>(&u);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = cvt<TQ>(p[j]);
      }
      if constexpr (sizeof(TC) == 4) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(b + i)), w = __ldg(reinterpret_cast<const float4*>(b + i + 4));
        y[0] = u.x; y[1] = u.y; y[2] = u.z; y[3] = u.w; y[4] = w.x; y[5] = w.y; y[6] = w.z; y[7] = w.w;
      } else {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(b + i));
        const TC* p = reinterpret_cast<const TC*>(&u);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = cvt<TC>(p[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(x[j], y[j], acc);
    }
  } else {
    for (int i = lane; i < dim; i += 32) acc = fmaf(cvt<TQ>(a[i]), cvt<TC>(b[i]), acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}


// ---------------------------------------------------------------- multi-modality rows
// score(q, c) = sum_m weight_m * <q_m, c_m> * q_inv_m[q] * c_inv_m[c]  over up to kMaxSeg modalities, each with its own
// embeddings (own dim / dtype / stride).  Element types are switched at run time (uniform per segment).
constexpr int kMaxSeg = 4;
struct Segments {
  const void* q_src[kMaxSeg];
  const void* c_src[kMaxSeg];
  const float* q_inv[kMaxSeg];
  const float* c_inv[kMaxSeg];
  int64_t q_stride[kMaxSeg], c_stride[kMaxSeg];
  int dim[kMaxSeg], q_dtype[kMaxSeg], c_dtype[kMaxSeg];
  float weight[kMaxSeg];
  int vec_ok[kMaxSeg];   // both sides 16-byte aligned rows, dim % 8 == 0: 128-bit loads
  int n;
};

__device__ __forceinline__ float load_rt(const void* base, int dtype, int64_t i) {
  if (dtype == MMD_SRC_F32) return static_cast<const float*>(base)[i];
  if (dtype == MMD_SRC_F16) return __half2float(static_cast<const __half*>(base)[i]);
  return __bfloat162float(static_cast<const __nv_bfloat16*>(base)[i]);
}


// <a, b> over `dim` elements with run-time element types; the 128-bit vectorised path of warp_dot whenever both rows allow it.
__device__ __forceinline__ float warp_dot_rt(const void* a, int a_dtype, const void* b, int b_dtype, int64_t a_off, int64_t b_off,
                                             int dim, int lane, bool vec_ok) {
#define MMD_DOT_CASE(TA, TB) return warp_dot<TA, TB>(static_cast<const TA*>(a) + a_off, static_cast<const TB*>(b) + b_off, dim, lane, vec_ok)
  if (a_dtype == MMD_SRC_F32) {
    if (b_dtype == MMD_SRC_F32) MMD_DOT_CASE(float, float);
    if (b_dtype == MMD_SRC_F16) MMD_DOT_CASE(float, __half);
    MMD_DOT_CASE(float, __nv_bfloat16);
  }
  if (a_dtype == MMD_SRC_F16) {
    if (b_dtype == MMD_SRC_F32) MMD_DOT_CASE(__half, float);
    if (b_dtype == MMD_SRC_F16) MMD_DOT_CASE(__half, __half);
    MMD_DOT_CASE(__half, __nv_bfloat16);
  }
  if (b_dtype == MMD_SRC_F32) MMD_DOT_CASE(__nv_bfloat16, float);
  if (b_dtype == MMD_SRC_F16) MMD_DOT_CASE(__nv_bfloat16, __half);
  MMD_DOT_CASE(__nv_bfloat16, __nv_bfloat16);
#undef MMD_DOT_CASE
}

inline int elem_size_of(int src_dtype) { return src_dtype == MMD_SRC_F32 ? 4 : 2; }
inline bool segment_vec_ok(const void* q_src, int q_dtype, int64_t q_stride, const void* c_src, int c_dtype, int64_t c_stride, int dim) {
  return dim % 8 == 0 && reinterpret_cast<uintptr_t>(q_src) % 16 == 0 && reinterpret_cast<uintptr_t>(c_src) % 16 == 0 &&
         (q_stride * elem_size_of(q_dtype)) % 16 == 0 && (c_stride * elem_size_of(c_dtype)) % 16 == 0;
}

// Host side: validate and copy the per-modality tables of the C ABI (HOST arrays of n_seg entries) into a Segments.
inline int fill_segments(Segments* out, int n_seg, const void* const* q_src_host, const int* q_dtype_host, const int64_t* q_stride_host,
                         const float* const* q_inv_host, const void* const* c_src_host, const int* c_dtype_host,
                         const int64_t* c_stride_host, const float* const* c_inv_host, const int* dim_host, const float* weight_host,
                         int64_t N, const char* who) {
  MMD_REQUIRE(n_seg >= 1 && n_seg <= kMaxSeg, "%s: n_seg=%d (1..%d)", who, n_seg, kMaxSeg);
  MMD_REQUIRE(q_src_host != nullptr && c_src_host != nullptr && q_dtype_host != nullptr && c_dtype_host != nullptr &&
              q_stride_host != nullptr && c_stride_host != nullptr && dim_host != nullptr,
              "%s: null segment table", who);
  Segments sg{};
  sg.n = n_seg;
  for (int m = 0; m < n_seg; ++m) {
    MMD_REQUIRE(q_src_host[m] != nullptr && (c_src_host[m] != nullptr || N == 0) && dim_host[m] > 0,
                "%s: segment %d has a null buffer or non-positive dim", who, m);
    MMD_REQUIRE(q_dtype_host[m] >= 0 && q_dtype_host[m] <= 2 && c_dtype_host[m] >= 0 && c_dtype_host[m] <= 2,
                "%s: segment %d has an unknown dtype", who, m);
    MMD_REQUIRE(q_stride_host[m] >= dim_host[m] && (c_stride_host[m] >= dim_host[m] || N == 0),
                "%s: segment %d row stride smaller than dim", who, m);
    sg.q_src[m] = q_src_host[m]; sg.c_src[m] = c_src_host[m];
    sg.q_inv[m] = q_inv_host != nullptr ? q_inv_host[m] : nullptr;
    sg.c_inv[m] = c_inv_host != nullptr ? c_inv_host[m] : nullptr;
    sg.q_stride[m] = q_stride_host[m]; sg.c_stride[m] = c_stride_host[m];
    sg.dim[m] = dim_host[m]; sg.q_dtype[m] = q_dtype_host[m]; sg.c_dtype[m] = c_dtype_host[m];
    sg.weight[m] = weight_host != nullptr ? weight_host[m] : 1.0f;
    sg.vec_ok[m] = segment_vec_ok(q_src_host[m], q_dtype_host[m], q_stride_host[m], c_src_host[m], c_dtype_host[m], c_stride_host[m],
                                  dim_host[m]) ? 1 : 0;
  }
  *out = sg;
  return MMD_OK;
}

// Exact score of (query q, local corpus row `row`): every lane returns the full sum.
__device__ __forceinline__ float segments_score(const Segments& sg, int64_t q, int64_t row, int lane) {
  float total = 0.0f;
  for (int m = 0; m < sg.n; ++m) {
    const float acc = warp_dot_rt(sg.q_src[m], sg.q_dtype[m], sg.c_src[m], sg.c_dtype[m], q * sg.q_stride[m], row * sg.c_stride[m],
                                  sg.dim[m], lane, sg.vec_ok[m] != 0);
    const float qi = sg.q_inv[m] != nullptr ? sg.q_inv[m][q] : 1.0f;
    const float ci = sg.c_inv[m] != nullptr ? sg.c_inv[m][row] : 1.0f;
    total = fmaf(sg.weight[m], acc * qi * ci, total);      // one modality, weight 1: exactly acc * qi * ci
  }
  return total;
}

}  // namespace mmd
