// Warp-cooperative bitonic networks over 64-bit candidate keys: element i = e * 32 + lane, E registers per lane.
#pragma once
#include <cstdint>

namespace mmd {

constexpr uint32_t kWarpFull = 0xffffffffu;

// Full sort, descending in i.
template <int E>
__device__ __forceinline__ void warp_bitonic_desc(uint64_t (&k)[E], int lane) {
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
            const uint64_t a = k[e], b = k[e | es];
            const uint64_t mx = a > b ? a : b, mn = a > b ? b : a;
            k[e] = desc ? mx : mn;
            k[e | es] = desc ? mn : mx;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const uint64_t other = __shfl_xor_sync(kWarpFull, k[e], stride);
          const bool lower = (lane & stride) == 0;
          const bool desc = ((e * 32 + lane) & size) == 0;
          const uint64_t mx = k[e] > other ? k[e] : other, mn = k[e] > other ? other : k[e];
          k[e] = (lower == desc) ? mx : mn;
        }
      }
    }
  }
}

// Bitonic merge network: `k` holds a bitonic sequence over i = e * 32 + lane; afterwards it is descending in i.
template <int E>
__device__ __forceinline__ void warp_bitonic_merge_desc(uint64_t (&k)[E], int lane) {
#pragma unroll
  for (int stride = 16 * E; stride > 0; stride >>= 1) {
    if (stride >= 32) {
      const int es = stride >> 5;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if ((e & es) == 0) {
          const uint64_t a = k[e], b = k[e | es];
          k[e] = a > b ? a : b;
          k[e | es] = a > b ? b : a;
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const uint64_t other = __shfl_xor_sync(kWarpFull, k[e], stride);
        const bool lower = (lane & stride) == 0;
        const uint64_t mx = k[e] > other ? k[e] : other, mn = k[e] > other ? other : k[e];
        k[e] = lower ? mx : mn;
      }
    }
  }
}


}  // namespace mmd
