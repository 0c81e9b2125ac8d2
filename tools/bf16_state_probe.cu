// Which shape of the bf16-optimiser-state Adam stream (theta f32 rw, m/v bf16 rw = 16 B per parameter) reaches HBM
// speed on B200?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bf16_state_probe tools/bf16_state_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void adam(float &p, float &m, float &v, float g) {
  m = 0.9f * m + 0.1f * g; v = 0.999f * v + 0.001f * g * g; p = p - __fdividef(1e-3f * m, fast_sqrt(v) + 1e-8f);
}
template <int HASH> __device__ __forceinline__ uint32_t rnd(uint32_t e, uint32_t step) {
  if (HASH == 0) return 0x8000u | (0x8000u << 16);                 // round half up, no hash
  uint32_t h = e * 0x9E3779B1u ^ step * 0x85EBCA77u;
  if (HASH == 2) { h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15; }
  else { h ^= h >> 16; h *= 0x2C1B3C6Du; h ^= h >> 15; }
  return h;
}
__device__ __forceinline__ uint32_t sr(float x, uint32_t r16) { return (__float_as_uint(x) + r16) >> 16; }

// ELEMS per thread in {4, 8}; ADJ: the two 4-chunks of an 8-element thread are adjacent (1) or 256 chunks apart (0)
template <int ELEMS, int HASH, int WIDE>
__global__ void __launch_bounds__(256) probe(float4 *th, uint2 *s1, uint2 *s2, long long n4, uint32_t step) {
  constexpr int CH = ELEMS / 4;
  const long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * CH;
  if (base + CH > n4) return;
  float4 p[CH]; uint2 a[CH], b[CH];
  if (WIDE && CH == 2) {
    const uint4 ua = reinterpret_cast<const uint4 *>(s1)[base / 2], ub = reinterpret_cast<const uint4 *>(s2)[base / 2];
    a[0] = make_uint2(ua.x, ua.y); a[CH - 1] = make_uint2(ua.z, ua.w);
    b[0] = make_uint2(ub.x, ub.y); b[CH - 1] = make_uint2(ub.z, ub.w);
#pragma unroll
    for (int c = 0; c < CH; ++c) p[c] = th[base + c];
  } else {
#pragma unroll
    for (int c = 0; c < CH; ++c) { p[c] = th[base + c]; a[c] = s1[base + c]; b[c] = s2[base + c]; }
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    float pv[4] = {p[c].x, p[c].y, p[c].z, p[c].w};
    float m[4] = {__uint_as_float(a[c].x << 16), __uint_as_float(a[c].x & 0xffff0000u), __uint_as_float(a[c].y << 16), __uint_as_float(a[c].y & 0xffff0000u)};
    float v[4] = {__uint_as_float(b[c].x << 16), __uint_as_float(b[c].x & 0xffff0000u), __uint_as_float(b[c].y << 16), __uint_as_float(b[c].y & 0xffff0000u)};
    uint32_t om[4], ov[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      adam(pv[j], m[j], v[j], 0.01f * pv[j]);
      const uint32_t h = rnd<HASH>((uint32_t)((base + c) * 4 + j), step);
      om[j] = sr(m[j], h & 0xffffu); ov[j] = sr(v[j], h >> 16);
    }
    p[c] = make_float4(pv[0], pv[1], pv[2], pv[3]);
    a[c] = make_uint2(om[0] | (om[1] << 16), om[2] | (om[3] << 16));
    b[c] = make_uint2(ov[0] | (ov[1] << 16), ov[2] | (ov[3] << 16));
  }
  if (WIDE && CH == 2) {
#pragma unroll
    for (int c = 0; c < CH; ++c) th[base + c] = p[c];
    reinterpret_cast<uint4 *>(s1)[base / 2] = make_uint4(a[0].x, a[0].y, a[CH - 1].x, a[CH - 1].y);
    reinterpret_cast<uint4 *>(s2)[base / 2] = make_uint4(b[0].x, b[0].y, b[CH - 1].x, b[CH - 1].y);
  } else {
#pragma unroll
    for (int c = 0; c < CH; ++c) { th[base + c] = p[c]; s1[base + c] = a[c]; s2[base + c] = b[c]; }
  }
}

template <int ELEMS, int HASH, int WIDE>
void run(const char *name, float4 *th, uint2 *s1, uint2 *s2, long long n4) {
  const int ch = ELEMS / 4;
  const int blocks = (int)((n4 / ch + 255) / 256);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) probe<ELEMS, HASH, WIDE><<<blocks, 256>>>(th, s1, s2, n4, i);
  CK(cudaEventRecord(e0));
  const int reps = 20;
  for (int i = 0; i < reps; ++i) probe<ELEMS, HASH, WIDE><<<blocks, 256>>>(th, s1, s2, n4, 10 + i);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  printf("%-52s %7.1f us  %6.0f GB/s (16 B/param)\n", name, ms * 1e3, 16.0 * n4 * 4 / (ms * 1e-3) / 1e9);
}

int main() {
  const long long n = 19200000, n4 = n / 4;
  float4 *th; uint2 *s1, *s2;
  CK(cudaMalloc(&th, n * 4)); CK(cudaMalloc(&s1, n * 2)); CK(cudaMalloc(&s2, n * 2));
  CK(cudaMemset(th, 0, n * 4)); CK(cudaMemset(s1, 0, n * 2)); CK(cudaMemset(s2, 0, n * 2));
  run<4, 2, 0>("4 elems/thread, 8 B state accesses, full hash", th, s1, s2, n4);
  run<4, 1, 0>("4 elems/thread, 8 B state accesses, short hash", th, s1, s2, n4);
  run<4, 0, 0>("4 elems/thread, 8 B state accesses, no hash", th, s1, s2, n4);
  run<8, 2, 0>("8 elems/thread, 2 x 8 B state accesses, full hash", th, s1, s2, n4);
  run<8, 2, 1>("8 elems/thread, 16 B state accesses, full hash", th, s1, s2, n4);
  run<8, 1, 1>("8 elems/thread, 16 B state accesses, short hash", th, s1, s2, n4);
  run<8, 0, 1>("8 elems/thread, 16 B state accesses, no hash", th, s1, s2, n4);
  return 0;
}
