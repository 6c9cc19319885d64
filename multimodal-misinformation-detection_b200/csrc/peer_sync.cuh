// Cross-GPU synchronisation folded into the kernels of the row-sharded path (no barrier launches).
//
// Every stage of a sharded step (candidate scatter -> exchange+re-score -> finish) is one kernel per rank.  A stage
// "arrives" when its LAST block has finished: that block writes the stage's sequence number into one flag word on every
// peer (peer-mapped memory, NVLink).  The consuming stage on every rank spins at its start until all `world` flags of its
// own array have reached the sequence number it expects.  Sequence numbers live in device memory (they advance by one
// per launch), so the same launches can be replayed from a CUDA graph.
//
// Safe only with one process per GPU (every rank's kernels run concurrently on different devices); a producer never
// waits for its consumers, so there is no cyclic wait.  A wait that does not complete within 4 s traps (fails loudly
// instead of hanging the box).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mmd {

constexpr int kMaxPeers = 16;

// What a stage signals when it is done.  state[0] = launches of this stage completed so far (sequence number),
// state[1] = blocks of the current launch that have finished.  n == 0: nothing to signal.
struct PeerArrive {
  uint32_t* flag[kMaxPeers];   // this rank's flag word in every rank's flag array (own one included)
  int n;
  uint32_t* state;
};

// What a stage waits for at its start: flags[0..n) (local memory, written by the peers) >= state[0] + 1.
struct PeerWait {
  const uint32_t* flags;
  int n;
  const uint32_t* state;       // the waiting stage's own sequence word
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t peer_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Called by EVERY thread of EVERY block at the start of a consuming kernel (no divergent exits before it).
__device__ __forceinline__ void peer_wait_all(const PeerWait& w) {
  if (w.n == 0) return;
  if (static_cast<int>(threadIdx.x) < w.n) {
    const uint32_t want = *reinterpret_cast<const volatile uint32_t*>(w.state) + 1u;
    const uint32_t* f = w.flags + threadIdx.x;
    if (static_cast<int32_t>(ld_acquire_sys(f) - want) < 0) {
      const uint64_t t0 = peer_timer_ns();
      uint32_t spins = 0;
      while (static_cast<int32_t>(ld_acquire_sys(f) - want) < 0) {
        if ((++spins & 0xff) == 0 && peer_timer_ns() - t0 > 4000000000ull) __trap();
      }
    }
  }
  __syncthreads();
}

// Called by EVERY thread of EVERY block at the end of a producing kernel (no divergent exits before it).
// bump_only: no flags to write (n == 0) but the stage still counts its launches (a consuming-only stage).
__device__ __forceinline__ void peer_arrive_all(const PeerArrive& a) {
  if (a.state == nullptr) return;
  __threadfence_system();          // this thread's stores (to peers' buffers) are ordered before the flag
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t total = gridDim.x * gridDim.y * gridDim.z;
    const uint32_t done = atomicAdd(a.state + 1, 1u);
    if (done == total - 1u) {
      __threadfence_system();
      const uint32_t v = *reinterpret_cast<volatile uint32_t*>(a.state) + 1u;
      for (int i = 0; i < a.n; ++i) st_release_sys(a.flag[i], v);
      a.state[1] = 0u;
      *reinterpret_cast<volatile uint32_t*>(a.state) = v;
      __threadfence();
    }
  }
}

}  // namespace mmd
