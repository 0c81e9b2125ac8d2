// 64-bit candidate keys of the running top-k: (order-preserving score bits << 32) | ~row id.
// Larger key == better score; equal scores order by LOWER row id first.  Key 0 is "no candidate".
#pragma once

#include "common.cuh"

namespace sert {
#ifdef __CUDACC__
__device__ __forceinline__ unsigned int orderable(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorderable(unsigned int o) {
  const unsigned int u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}
__device__ __forceinline__ unsigned long long make_key(float score, unsigned int row) {
  return ((unsigned long long)orderable(score) << 32) | (unsigned long long)(0xffffffffu - row);
}
__device__ __forceinline__ unsigned int key_row(unsigned long long key) {
  return 0xffffffffu - (unsigned int)(key & 0xffffffffull);
}
__device__ __forceinline__ float key_score(unsigned long long key) { return unorderable((unsigned int)(key >> 32)); }
#endif
}  // namespace sert
