// Library-level plumbing of the C ABI: error text, device gate, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace mmd {

const DeviceStatus* host_status_word();

namespace {
thread_local char g_err[768] = {0};
std::atomic<int64_t> g_launches{0};
}  // namespace

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace mmd

extern "C" int mmd_abi_version(void) { return MMD_ABI_VERSION; }

extern "C" const char* mmd_last_error(void) {
  // If a pipeline watchdog fired, say where: the status word is pinned host memory and survives the trap.
  const mmd::DeviceStatus* st = mmd::host_status_word();
  if (st != nullptr && st->code != 0) {
    static thread_local char buf[1024];
    snprintf(buf, sizeof(buf), "%s [device watchdog: wait site %d timed out in block %d thread %d]", mmd::g_err,
             st->site, st->block, st->extra);
    return buf;
  }
  return mmd::g_err;
}

extern "C" int mmd_device_check(void) {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = MMD_ERR_DEVICE;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) {
    cudaGetLastError();
    mmd::set_last_error("no CUDA device is available; this library has no CPU fallback");
    return MMD_ERR_DEVICE;
  }
  if (dev == cached_dev) return cached_rc;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cached_dev = dev;
  if (major != 10) {
    mmd::set_last_error("device %d has compute capability %d.%d; the kernels are built for sm_100a only", dev, major, minor);
    cached_rc = MMD_ERR_DEVICE;
  } else {
    cached_rc = MMD_OK;
  }
  return cached_rc;
}

extern "C" int64_t mmd_launch_count(void) { return mmd::g_launches.load(std::memory_order_relaxed); }

#ifndef MMD_BUILD_INFO
#define MMD_BUILD_INFO "src=unknown (built without build.py)"
#endif
extern "C" const char* mmd_build_info(void) { return MMD_BUILD_INFO; }
