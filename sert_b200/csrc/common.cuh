// Shared host/device helpers for libsert_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/sert_b200.h"

namespace sert {

// ---- host-side error plumbing -------------------------------------------------------------
void set_error(const std::string &msg);
void count_launch(uint64_t n = 1);

#define SERT_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::sert::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                        ":" + std::to_string(__LINE__) + ")");                                \
      return -1;                                                                              \
    }                                                                                         \
  } while (0)

#define SERT_REQUIRE(cond, msg)                                                               \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      ::sert::set_error(std::string(msg) + " [" #cond "]");                                   \
      return -1;                                                                              \
    }                                                                                         \
  } while (0)

#define SERT_LAUNCH_CHECK()                                                                   \
  do {                                                                                        \
    ::sert::count_launch();                                                                   \
    SERT_CUDA(cudaGetLastError());                                                            \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// True the first time it is called for the current device with a given `mask` (one static std::atomic<uint64_t> per
// call site): per-device one-time set-up such as cudaFuncSetAttribute, safe when several host threads drive
// different devices of one process.
bool first_use_on_device(std::atomic<uint64_t> &mask);
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Reference constants (sert/models.py:200,290,900,1067-1068).  In float32, 1-1e-7 rounds to
// 0.99999988f; the literals below are evaluated in double then rounded once, like Theano's
// float32 graph constants.
#define SERT_CLIP_LO 1e-7f
#define SERT_CLIP_HI ((float)(1.0 - 1e-7))
#define SERT_TANH_LO ((float)(-1.0 + 1e-7))
#define SERT_TANH_HI ((float)(1.0 - 1e-7))

#ifdef __CUDACC__
// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Vectorised no-return float4 reduction into global memory (sm_90+: red.global.add.v4.f32).
__device__ __forceinline__ void red_add_f4(float *addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// Streaming (evict-first) 16-byte accesses for data touched once per step.
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream_f4(float4 *p, float4 v) { __stcs(p, v); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float clipf_(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
#endif  // __CUDACC__

}  // namespace sert
