// Host-side plumbing shared by every translation unit of liblvb200.so: error reporting for the
// C ABI, launch accounting, and TMA tensor-map encoding through the driver entry point (so the
// library does not link libcuda directly).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <atomic>

#include "../../include/lvb200.h"

namespace lv {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();
// Make the device that owns `ptr` current on the calling thread (autograd runs backward on its own
// threads, which may have no current CUDA context yet).  Returns LV_OK or LV_ECUDA.
int bind_device(const void* ptr);

// One-time per-DEVICE setup of a call site (cudaFuncSetAttribute is a per-device property and the library serves
// several devices from one process, possibly from several threads: a per-process `static bool` would leave the
// second GPU with the 48 KB default and race).  Usage: static PerDeviceOnce once; int dev; if (once.needed(&dev))
// { ...setup...; once.done(dev); }   A concurrent duplicate setup is harmless (the attributes are idempotent).
class PerDeviceOnce {
  std::atomic<int> done_[64] = {};

 public:
  bool needed(int* dev) {
    if (cudaGetDevice(dev) != cudaSuccess || *dev < 0 || *dev >= 64) {
      *dev = -1;
      return true;   // unknown device: set the attributes on every call
    }
    return done_[*dev].load(std::memory_order_acquire) == 0;
  }
  void done(int dev) {
    if (dev >= 0) done_[dev].store(1, std::memory_order_release);
  }
};

// Encode a bf16 tiled tensor map.  dims/strides innermost-first; strides in BYTES for dims 1..rank-1
// (dim 0 is contiguous).  box = tile extents, innermost-first.  swizzle128: 128-byte swizzle
// (box[0] * 2 bytes must be <= 128), otherwise no swizzle.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128);

#define LV_BIND_DEVICE(ptr)                \
  do {                                     \
    int _r = lv::bind_device(ptr);         \
    if (_r != LV_OK) return _r;            \
  } while (0)

#define LV_CHECK_ARG(cond, ...)   \
  do {                            \
    if (!(cond)) {                \
      lv::set_error(__VA_ARGS__); \
      return LV_EINVAL;           \
    }                             \
  } while (0)

#define LV_CHECK_CUDA(expr)                                                                     \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      lv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return LV_ECUDA;                                                                          \
    }                                                                                           \
  } while (0)

#define LV_CHECK_LAUNCH(name)                                                             \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      lv::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));             \
      return LV_ECUDA;                                                                    \
    }                                                                                     \
    lv::count_launch();                                                                   \
  } while (0)

}  // namespace lv
