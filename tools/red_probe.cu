// Micro-benchmark: what bounds the scatter-add of 512-byte gradient rows (the backward of the embedding gathers,
// sert/models.py:180,990)?  86 016 row additions into a (rows x 128) float table per launch, the count of one
// BASELINE configs[1] batch (4096 instances x (10 word rows + 11 entity rows)).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/red_probe.cu -o tools/red_probe
//   mode 0  st.global.v4.f32            (plain store, no atomic: the floor)
//   mode 1  red.global.add.v4.f32       (what csrc/vs_warp.cu / vs_tile.cu issue)
//   mode 2  4 x red.global.add.f32
//   mode 3  cp.reduce.async.bulk.global.shared::cta.add.f32, 512 B per request, one thread per row
//   mode 4  red.global.add.v4.f32 after a prefetch.global.L2 of every target line (separate kernel, timed apart)
// Each mode runs on a table that misses L2 (51 MB, L2 flushed by a 512 MB memset before every launch) and on one that
// stays L2-resident (4 MB, no flush), with uniform row ids and with Zipf-like hot rows.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int kRowFloats = 128;
constexpr int kOpsPerWarp = 21;

__device__ __forceinline__ void red4(float *p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red1(float *p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256) scatter_kernel(float *table, const int *rows, int n_ops) {
  __shared__ __align__(128) float srow[8][kRowFloats];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + warp;
  const int base = gw * kOpsPerWarp;
  const float4 v = make_float4(1.f + lane, 2.f, 3.f, 4.f);
  if (MODE == 3) {
    reinterpret_cast<float4 *>(srow[warp])[lane] = v;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane < kOpsPerWarp && base + lane < n_ops) {
      const int r = rows[base + lane];
      const unsigned int s = (unsigned int)__cvta_generic_to_shared(srow[warp]);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 512;"
                   ::"l"(table + (size_t)r * kRowFloats), "r"(s) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    return;
  }
  int my = (lane < kOpsPerWarp && base + lane < n_ops) ? rows[base + lane] : -1;
#pragma unroll 7
  for (int j = 0; j < kOpsPerWarp; ++j) {
    const int r = __shfl_sync(0xffffffffu, my, j);
    if (r < 0) continue;
    float *p = table + (size_t)r * kRowFloats + lane * 4;
    if (MODE == 0) *reinterpret_cast<float4 *>(p) = v;
    if (MODE == 1 || MODE == 4) red4(p, v);
    if (MODE == 2) { red1(p, v.x); red1(p + 1, v.y); red1(p + 2, v.z); red1(p + 3, v.w); }
  }
}

__global__ void prefetch_kernel(const float *table, const int *rows, int n_ops) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per 128-byte line
  const int op = t >> 2;
  if (op < n_ops) {
    const float *p = table + (size_t)rows[op] * kRowFloats + (t & 3) * 32;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  }
}

template <int MODE>
static float run(float *table, const int *rows, int n_ops, void *flush, size_t flush_bytes, bool do_flush, float *pre_ms) {
  cudaEvent_t e0, e1, e2;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
  const int warps = (n_ops + kOpsPerWarp - 1) / kOpsPerWarp;
  const int grid = (warps + 7) / 8;
  float best = 1e9f, best_pre = 0.f;
  for (int rep = 0; rep < 6; ++rep) {
    if (do_flush) CK(cudaMemsetAsync(flush, rep, flush_bytes));
    CK(cudaEventRecord(e0));
    if (MODE == 4) prefetch_kernel<<<(n_ops * 4 + 255) / 256, 256>>>(table, rows, n_ops);
    CK(cudaEventRecord(e1));
    scatter_kernel<MODE><<<grid, 256>>>(table, rows, n_ops);
    CK(cudaEventRecord(e2));
    CK(cudaEventSynchronize(e2));
    float a, b;
    CK(cudaEventElapsedTime(&a, e0, e1));
    CK(cudaEventElapsedTime(&b, e1, e2));
    if (rep > 0 && b < best) { best = b; best_pre = a; }
  }
  *pre_ms = best_pre;
  return best;
}

int main() {
  const int n_ops = 4096 * kOpsPerWarp;
  const size_t flush_bytes = 512ull << 20;
  void *flush;
  CK(cudaMalloc(&flush, flush_bytes));
  struct Case { const char *name; int rows; bool flush; bool zipf; };
  const Case cases[] = {
      {"51 MB table, L2 flushed, uniform rows", 100000, true, false},
      {"51 MB table, L2 flushed, Zipf rows   ", 100000, true, true},
      {" 4 MB table, L2 resident, uniform rows", 8192, false, false},
      {" 4 MB table, L2 resident, Zipf rows   ", 8192, false, true},
  };
  for (const Case &c : cases) {
    float *table;
    CK(cudaMalloc(&table, (size_t)c.rows * kRowFloats * 4));
    CK(cudaMemset(table, 0, (size_t)c.rows * kRowFloats * 4));
    std::vector<int> h(n_ops);
    uint64_t s = 88172645463325252ull;
    std::vector<double> cdf;
    if (c.zipf) {
      cdf.resize(c.rows);
      double tot = 0;
      for (int r = 0; r < c.rows; ++r) { tot += 1.0 / pow(r + 2.7, 1.07); cdf[r] = tot; }
      for (double &x : cdf) x /= tot;
    }
    for (int i = 0; i < n_ops; ++i) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      if (c.zipf) {
        const double u = (double)(s >> 11) / 9007199254740992.0;
        h[i] = (int)(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
        if (h[i] >= c.rows) h[i] = c.rows - 1;
      } else {
        h[i] = (int)(s % (uint64_t)c.rows);
      }
    }
    int *rows;
    CK(cudaMalloc(&rows, n_ops * 4));
    CK(cudaMemcpy(rows, h.data(), n_ops * 4, cudaMemcpyHostToDevice));
    float pre;
    const double mb = n_ops * 512.0 / 1e6;
    printf("%s  (%d row additions = %.1f MB, %.1f M elements)\n", c.name, n_ops, mb, n_ops * 128 / 1e6);
    float t;
    t = run<0>(table, rows, n_ops, flush, flush_bytes, c.flush, &pre); printf("  st.v4            %8.2f us  %7.1f GB/s\n", t * 1e3, mb / t);
    t = run<1>(table, rows, n_ops, flush, flush_bytes, c.flush, &pre); printf("  red.v4.f32       %8.2f us  %7.1f GB/s\n", t * 1e3, mb / t);
    t = run<2>(table, rows, n_ops, flush, flush_bytes, c.flush, &pre); printf("  4 x red.f32      %8.2f us  %7.1f GB/s\n", t * 1e3, mb / t);
    t = run<3>(table, rows, n_ops, flush, flush_bytes, c.flush, &pre); printf("  bulk reduce 512B %8.2f us  %7.1f GB/s\n", t * 1e3, mb / t);
    t = run<4>(table, rows, n_ops, flush, flush_bytes, c.flush, &pre); printf("  L2 prefetch %6.2f us, then red.v4 %8.2f us\n", pre * 1e3, t * 1e3);
    CK(cudaFree(table));
    CK(cudaFree(rows));
  }
  return 0;
}
