// Shared helpers for the liblfp_sg2 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/lfp_sg2.h"

namespace lfp {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define LFP_CHECK_ARG(cond, ...)             \
  do {                                       \
    if (!(cond)) {                           \
      ::lfp::set_error(__VA_ARGS__);         \
      return LFP_EINVAL;                     \
    }                                        \
  } while (0)

#define LFP_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::lfp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                        \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)

// every launch is followed by this: launch errors surface at the call that caused them
// (the reference never checks)
#define LFP_LAUNCH_CHECK()                                                            \
  do {                                                                                \
    ::lfp::count_launch();                                                            \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      ::lfp::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),    \
                       __FILE__, __LINE__);                                           \
      return (int)_e;                                                                 \
    }                                                                                 \
  } while (0)

#define LFP_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

static inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Device staging buffers of the host-pointer ("_host") entry points: per host thread and device, grow-only, so a caller
// looping over such an entry point allocates nothing after its first iteration and no return path can leak (round 1
// allocated and freed on every call and leaked on an error between two allocations).  Freed when the thread exits.
struct HostStaging {
  static constexpr int kSlots = 4;
  int dev = -1;
  void* p[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap[kSlots] = {0, 0, 0, 0};
  ~HostStaging() { release(); }
  void release() {
    for (int i = 0; i < kSlots; ++i) { if (p[i]) cudaFree(p[i]); p[i] = nullptr; cap[i] = 0; }
  }
  // nullptr on allocation failure (the slot is left empty)
  void* get(int slot, size_t bytes) {
    int d = 0;
    cudaGetDevice(&d);
    if (d != dev) { release(); dev = d; }
    if (cap[slot] < bytes) {
      if (p[slot]) cudaFree(p[slot]);
      p[slot] = nullptr; cap[slot] = 0;
      if (cudaMalloc(&p[slot], bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      cap[slot] = bytes;
    }
    return p[slot];
  }
};
inline HostStaging& host_staging() {
  static thread_local HostStaging hs;
  return hs;
}

__device__ __forceinline__ int floor_div_i(int a, int b) {
  int q = a / b;
  return (q * b > a) ? q - 1 : q;
}

// streaming (read-once) 128-bit global load / store
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w));
}

}  // namespace lfp
