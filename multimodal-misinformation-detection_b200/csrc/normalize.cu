// K1: fused row L2-normalise + cast into the operand layout the tensor-core contraction reads.
//
// Reference arithmetic this replaces (paths relative to the reference checkout):
//   sentence_transformers.util.cos_sim -> F.normalize(x, p=2, dim=1)   x / max(||x||_2, 1e-12)
//       call sites src/evidence/text2text_retrieval.py:56-63, src/evidence/experiment_text.py:25-32
//       (the reference re-normalises the WHOLE corpus on every query; here it is done once)
//   nn.CosineSimilarity(dim=1, eps=1e-6)  per-norm clamp               src/evidence/im2im_retrieval.py:38-42
//   question_embedding.to(dtype=torch.float16)  dtype cast             src/evidence/text2text_retrieval.py:53
//
// HBM-bound streaming kernel: one warp per row, 128-bit coalesced loads, the row is kept in
// registers between the sum-of-squares pass and the scale/cast pass (dim <= 2048), 128-bit stores.
// Algorithmic bytes per row: dim * sizeof(src) read + row_bytes written (+4 for inv_norm).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"

namespace mmd {
namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxCachedGroups = 8;  // 8 groups x 8 values x 32 lanes = 2048 values cached per row

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// Load the 8 consecutive source values of group g (zero beyond dim).
template <typename T, bool kVec>
__device__ __forceinline__ void load8(const T* __restrict__ row, int g, int dim, float (&v)[8]) {
  const int base = g * 8;
  if constexpr (kVec) {
    if (base + 8 <= dim) {
      if constexpr (sizeof(T) == 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(row + base));
        const float4 b = __ldg(reinterpret_cast<const float4*>(row + base + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(row + base));
        const T* p = reinterpret_cast<const T*>(&a);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = to_f32<T>(p[i]);
      }
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (base + i < dim) ? to_f32<T>(row[base + i]) : 0.0f;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Limb order along K for the 6 cross terms of the 3-limb bf16 split, smallest magnitude first:
//   segment:        0      1      2      3      4      5
//   (q limb,c limb) (3,1)  (1,3)  (2,2)  (2,1)  (1,2)  (1,1)
__device__ __constant__ int kLimbOfSegment[2][6] = {{2, 0, 1, 1, 0, 0},   // query side  (0-based limb)
                                                    {0, 2, 1, 0, 1, 0}};  // corpus side

template <int OP>
__device__ __forceinline__ void store8(uint8_t* __restrict__ dst_row, int g, int64_t dpad, int side,
                                       const float (&x)[8]) {
  if constexpr (OP == MMD_OP_BF16) {
    uint4 o;
    o.x = pack_bf16x2(x[0], x[1]); o.y = pack_bf16x2(x[2], x[3]);
    o.z = pack_bf16x2(x[4], x[5]); o.w = pack_bf16x2(x[6], x[7]);
    *reinterpret_cast<uint4*>(dst_row + static_cast<size_t>(g) * 16) = o;
  } else if constexpr (OP == MMD_OP_F16) {
    uint4 o;
    o.x = pack_f16x2(x[0], x[1]); o.y = pack_f16x2(x[2], x[3]);
    o.z = pack_f16x2(x[4], x[5]); o.w = pack_f16x2(x[6], x[7]);
    *reinterpret_cast<uint4*>(dst_row + static_cast<size_t>(g) * 16) = o;
  } else if constexpr (OP == MMD_OP_E4M3) {
    // unit-norm rows have |x| <= 1: a fixed 2^8 scale keeps them inside e4m3's range (max 448);
    // the contraction's epilogue divides by 2^16.
    uint32_t w[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const __nv_fp8x2_storage_t lo =
          __nv_cvt_float2_to_fp8x2(make_float2(x[4 * h + 0] * 256.0f, x[4 * h + 1] * 256.0f), __NV_SATFINITE, __NV_E4M3);
      const __nv_fp8x2_storage_t hi =
          __nv_cvt_float2_to_fp8x2(make_float2(x[4 * h + 2] * 256.0f, x[4 * h + 3] * 256.0f), __NV_SATFINITE, __NV_E4M3);
      w[h] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
    }
    *reinterpret_cast<uint2*>(dst_row + static_cast<size_t>(g) * 8) = make_uint2(w[0], w[1]);
  } else {  // MMD_OP_BF16X3
    float limb[3][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float l0 = __bfloat162float(__float2bfloat16_rn(x[i]));
      const float r1 = x[i] - l0;
      const float l1 = __bfloat162float(__float2bfloat16_rn(r1));
      const float r2 = r1 - l1;
      const float l2 = __bfloat162float(__float2bfloat16_rn(r2));
      limb[0][i] = l0; limb[1][i] = l1; limb[2][i] = l2;
    }
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const int l = kLimbOfSegment[side][s];
      float y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = (l == 0) ? limb[0][i] : (l == 1 ? limb[1][i] : limb[2][i]);
      uint4 o;
      o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
      o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
      *reinterpret_cast<uint4*>(dst_row + (static_cast<size_t>(s) * dpad + static_cast<size_t>(g) * 8) * 2) = o;
    }
  }
}

// G = number of 8-value groups each lane keeps in registers (row length <= 256 * G values); kTail = the row
// is longer than that and the rest is streamed twice.  Small G keeps the register count low, so more warps
// (and more 128-bit loads) are in flight per SM -- this kernel lives on memory-level parallelism.
template <typename T, int OP, bool kVec, int G, bool kTail>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
normalize_cast_kernel(const T* __restrict__ src, int64_t rows, int dim, int64_t src_stride, int normalize,
                      float eps, float scale, int side, uint8_t* __restrict__ dst, int64_t dpad, int64_t row_bytes,
                      float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t warp_stride = static_cast<int64_t>(gridDim.x) * kWarpsPerBlock;
  const int n_groups = static_cast<int>(dpad / 8);

  for (int64_t r = warp_global; r < rows; r += warp_stride) {
    const T* __restrict__ row = src + r * src_stride;
    uint8_t* __restrict__ drow = dst + r * row_bytes;

    float cache[G][8];
#pragma unroll
    for (int i = 0; i < G; ++i) {          // all loads first: G (x2 for fp32) independent 128-bit requests per lane
      const int g = lane + 32 * i;
      if (g < n_groups) {
        load8<T, kVec>(row, g, dim, cache[i]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) cache[i][j] = 0.0f;
      }
    }
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < G; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = fmaf(cache[i][j], cache[i][j], ss);
    if constexpr (kTail) {
      for (int g = lane + 32 * G; g < n_groups; g += 32) {
        float v[8];
        load8<T, kVec>(row, g, dim, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
      }
    }
    float denom = 1.0f;
    if (normalize) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      denom = fmaxf(sqrtf(ss), eps);
    }
    if (lane == 0 && inv_norm != nullptr) inv_norm[r] = 1.0f / denom;

#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int g = lane + 32 * i;
      if (g < n_groups) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = (normalize ? cache[i][j] / denom : cache[i][j]) * scale;
        store8<OP>(drow, g, dpad, side, y);
      }
    }
    if constexpr (kTail) {
      for (int g = lane + 32 * G; g < n_groups; g += 32) {
        float v[8];
        load8<T, kVec>(row, g, dim, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (normalize ? v[j] / denom : v[j]) * scale;
        store8<OP>(drow, g, dpad, side, v);
      }
    }
  }
}

template <typename T, int OP, bool kVec, int G, bool kTail>
void launch_g(unsigned grid, cudaStream_t stream, const T* s, int64_t rows, int dim, int64_t src_stride, int normalize,
              float eps, float scale, int side, uint8_t* d, const PreparedLayout& lay, float* inv_norm) {
  normalize_cast_kernel<T, OP, kVec, G, kTail><<<grid, kWarpsPerBlock * 32, 0, stream>>>(
      s, rows, dim, src_stride, normalize, eps, scale, side, d, lay.dpad, lay.row_bytes, inv_norm);
}

template <typename T, int OP, bool kVec>
void launch_v(unsigned grid, cudaStream_t stream, const T* s, int64_t rows, int dim, int64_t src_stride, int normalize,
              float eps, float scale, int side, uint8_t* d, const PreparedLayout& lay, float* inv_norm) {
  const int per_lane = static_cast<int>(ceil_div(lay.dpad / 8, 32));
  if (per_lane <= 1) launch_g<T, OP, kVec, 1, false>(grid, stream, s, rows, dim, src_stride, normalize, eps, scale, side, d, lay, inv_norm);
  else if (per_lane <= 2) launch_g<T, OP, kVec, 2, false>(grid, stream, s, rows, dim, src_stride, normalize, eps, scale, side, d, lay, inv_norm);
  else if (per_lane <= 3) launch_g<T, OP, kVec, 3, false>(grid, stream, s, rows, dim, src_stride, normalize, eps, scale, side, d, lay, inv_norm);
  else if (per_lane <= 4) launch_g<T, OP, kVec, 4, false>(grid, stream, s, rows, dim, src_stride, normalize, eps, scale, side, d, lay, inv_norm);
  else if (per_lane <= kMaxCachedGroups) launch_g<T, OP, kVec, kMaxCachedGroups, false>(grid, stream, s, rows, dim, src_stride, normalize, eps, scale, side, d, lay, inv_norm);
  else launch_g<T, OP, kVec, kMaxCachedGroups, true>(grid, stream, s, rows, dim, src_stride, normalize, eps, scale, side, d, lay, inv_norm);
}

template <typename T, int OP>
int launch(const void* src, int64_t rows, int dim, int64_t src_stride, int normalize, float eps, float scale, int side, void* dst,
           const PreparedLayout& lay, float* inv_norm, cudaStream_t stream) {
  const bool vec = (reinterpret_cast<uintptr_t>(src) % 16 == 0) && ((src_stride * sizeof(T)) % 16 == 0);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = ceil_div(rows, kWarpsPerBlock);
  const int64_t cap = static_cast<int64_t>(sms) * 8 * 8;  // several waves of resident CTAs per SM, grid-stride beyond
  const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
  auto* s = static_cast<const T*>(src);
  auto* d = static_cast<uint8_t*>(dst);
  if (vec) launch_v<T, OP, true>(grid, stream, s, rows, dim, src_stride, normalize, eps, scale, side, d, lay, inv_norm);
  else launch_v<T, OP, false>(grid, stream, s, rows, dim, src_stride, normalize, eps, scale, side, d, lay, inv_norm);
  count_launch();
  MMD_CUDA_OK(cudaGetLastError());
  return MMD_OK;
}

template <typename T>
int dispatch_op(int op, const void* src, int64_t rows, int dim, int64_t src_stride, int normalize, float eps, float scale, int side,
                void* dst, const PreparedLayout& lay, float* inv_norm, cudaStream_t stream) {
  switch (op) {
    case MMD_OP_BF16: return launch<T, MMD_OP_BF16>(src, rows, dim, src_stride, normalize, eps, scale, side, dst, lay, inv_norm, stream);
    case MMD_OP_F16: return launch<T, MMD_OP_F16>(src, rows, dim, src_stride, normalize, eps, scale, side, dst, lay, inv_norm, stream);
    case MMD_OP_E4M3: return launch<T, MMD_OP_E4M3>(src, rows, dim, src_stride, normalize, eps, scale, side, dst, lay, inv_norm, stream);
    case MMD_OP_BF16X3: return launch<T, MMD_OP_BF16X3>(src, rows, dim, src_stride, normalize, eps, scale, side, dst, lay, inv_norm, stream);
  }
  set_last_error("mmd_normalize_cast: unknown op_dtype %d", op);
  return MMD_ERR_ARG;
}

}  // namespace
}  // namespace mmd

extern "C" int mmd_prepared_layout(int op_dtype, int dim, int64_t* kdim, int64_t* row_bytes) {
  mmd::PreparedLayout lay;
  MMD_REQUIRE(mmd::prepared_layout(op_dtype, dim, &lay), "mmd_prepared_layout: bad op_dtype %d / dim %d", op_dtype, dim);
  if (kdim) *kdim = lay.kdim;
  if (row_bytes) *row_bytes = lay.row_bytes;
  return MMD_OK;
}

namespace mmd {
namespace {
int normalize_cast_any(const void* src, int src_dtype, int64_t rows, int dim, int64_t src_row_stride, int normalize,
                       float eps, float scale, int op_dtype, int side, void* dst, int64_t dst_row_bytes, float* inv_norm,
                       void* stream, const char* who) {
  MMD_REQUIRE(rows >= 0 && dim > 0, "%s: rows=%lld dim=%d", who, (long long)rows, dim);
  if (rows == 0) return MMD_OK;
  MMD_REQUIRE(src != nullptr && dst != nullptr, "%s: null buffer", who);
  MMD_REQUIRE(src_row_stride >= dim, "%s: src_row_stride %lld < dim %d", who, (long long)src_row_stride, dim);
  MMD_REQUIRE(side == MMD_SIDE_QUERY || side == MMD_SIDE_CORPUS, "%s: bad side %d", who, side);
  MMD_REQUIRE(reinterpret_cast<uintptr_t>(dst) % 16 == 0, "%s: dst must be 16-byte aligned", who);
  PreparedLayout lay;
  MMD_REQUIRE(prepared_layout(op_dtype, dim, &lay), "%s: bad op_dtype %d", who, op_dtype);
  if (dst_row_bytes != 0) {
    MMD_REQUIRE(op_dtype != MMD_OP_BF16X3, "%s: the 3-limb fp32 configuration cannot be laid out as a segment", who);
    MMD_REQUIRE(dst_row_bytes >= lay.row_bytes && dst_row_bytes % 16 == 0, "%s: dst_row_bytes %lld (segment needs %lld, 16-byte multiple)",
                who, (long long)dst_row_bytes, (long long)lay.row_bytes);
    lay.row_bytes = dst_row_bytes;
  }
  int rc = mmd_device_check();
  if (rc != MMD_OK) return rc;
  auto st = static_cast<cudaStream_t>(stream);
  switch (src_dtype) {
    case MMD_SRC_F32: return dispatch_op<float>(op_dtype, src, rows, dim, src_row_stride, normalize, eps, scale, side, dst, lay, inv_norm, st);
    case MMD_SRC_F16: return dispatch_op<__half>(op_dtype, src, rows, dim, src_row_stride, normalize, eps, scale, side, dst, lay, inv_norm, st);
    case MMD_SRC_BF16: return dispatch_op<__nv_bfloat16>(op_dtype, src, rows, dim, src_row_stride, normalize, eps, scale, side, dst, lay, inv_norm, st);
  }
  set_last_error("%s: unknown src_dtype %d", who, src_dtype);
  return MMD_ERR_ARG;
}
}  // namespace
}  // namespace mmd

extern "C" int mmd_normalize_cast(const void* src, int src_dtype, int64_t rows, int dim, int64_t src_row_stride,
                                  int normalize, float eps, int op_dtype, int side, void* dst, float* inv_norm,
                                  void* stream) {
  return mmd::normalize_cast_any(src, src_dtype, rows, dim, src_row_stride, normalize, eps, 1.0f, op_dtype, side, dst, 0,
                                 inv_norm, stream, "mmd_normalize_cast");
}

extern "C" int mmd_normalize_cast_segment(const void* src, int src_dtype, int64_t rows, int dim, int64_t src_row_stride,
                                          int normalize, float eps, float scale, int op_dtype, int side, void* dst,
                                          int64_t dst_row_bytes, float* inv_norm, void* stream) {
  MMD_REQUIRE(dst_row_bytes > 0, "mmd_normalize_cast_segment: dst_row_bytes must be positive");
  return mmd::normalize_cast_any(src, src_dtype, rows, dim, src_row_stride, normalize, eps, scale, op_dtype, side, dst,
                                 dst_row_bytes, inv_norm, stream, "mmd_normalize_cast_segment");
}
