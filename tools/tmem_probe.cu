// TMEM read (tcgen05.ld) throughput probe: how many bytes per cycle can the epilogue warps of ONE SM pull out of
// tensor memory?  Decides how the top-k / softmax epilogues of gemm_tc.cu are shaped.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/tmem_probe tools/tmem_probe.cu && tools/tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// variant 0: ld x32, wait, repeat.  variant 1: two ld x32 in flight, then one wait.  variant 2: four in flight.
template <int VARIANT>
__global__ void __launch_bounds__(1024, 1) probe(int iters, long long *cycles, uint32_t *sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_ptr))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int slices = nwarps / 4;                      // warps per lane quarter
  const int cols_per = 512 / (slices > 0 ? slices : 1);
  const uint32_t col0 = (uint32_t)((warp >> 2) * cols_per);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < cols_per; c += 32 * (VARIANT == 0 ? 1 : (VARIANT == 1 ? 2 : 4))) {
      if (VARIANT == 0) {
        uint32_t v[32];
        ld32(base + lane_base + col0 + c, v);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j];
      } else if (VARIANT == 1) {
        uint32_t v[32], w[32];
        ld32(base + lane_base + col0 + c, v);
        ld32(base + lane_base + col0 + c + 32, w);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j] ^ w[j];
      } else {
        uint32_t v[32], w[32], x[32], y[32];
        ld32(base + lane_base + col0 + c, v);
        ld32(base + lane_base + col0 + c + 32, w);
        ld32(base + lane_base + col0 + c + 64, x);
        ld32(base + lane_base + col0 + c + 96, y);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j] ^ w[j] ^ x[j] ^ y[j];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

int main() {
  long long *cyc;
  uint32_t *sink;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaMalloc(&sink, 4);
  const int iters = 200;
  for (int variant = 0; variant < 3; ++variant) {
    for (int nw : {4, 8, 16, 32}) {
      if (512 / (nw / 4) < 32 * (variant == 0 ? 1 : (variant == 1 ? 2 : 4))) continue;
      for (int rep = 0; rep < 2; ++rep) {
        if (variant == 0) probe<0><<<148, nw * 32>>>(iters, cyc, sink);
        if (variant == 1) probe<1><<<148, nw * 32>>>(iters, cyc, sink);
        if (variant == 2) probe<2><<<148, nw * 32>>>(iters, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < 148; ++i) mean += (double)h[i];
      mean /= 148;
      const double bytes = (double)iters * 128.0 * 512.0 * 4.0;      // every warp quarter reads its share of all 512 columns
      printf("variant %d (loads in flight %d), %2d warps: %.0f cycles for %d x 256 KB -> %.1f B/cycle/SM, %.0f cycles per 128x256 fp32 tile\n",
             variant, variant == 0 ? 1 : (variant == 1 ? 2 : 4), nw, mean, iters, bytes / mean, 131072.0 / (bytes / mean));
    }
  }
  return 0;
}
