// Micro-benchmark: which access pattern / cache policy reaches HBM peak for a 3-array in-place
// read-modify-write stream (the dense optimiser's traffic)?  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int MODE>
__device__ __forceinline__ float4 ld(const float4* p) {
  if (MODE == 0) return *p;
  if (MODE == 1) return __ldcs(p);
  if (MODE == 2) return __ldcg(p);
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
template <int MODE>
__device__ __forceinline__ void st(float4* p, float4 v) {
  if (MODE == 0) *p = v;
  else if (MODE == 1) __stcs(p, v);
  else if (MODE == 2) __stcg(p, v);
  else __stwt(p, v);
}
__device__ __forceinline__ float4 upd(float4 a, float4 b, float4 c, float s) {
  return make_float4(a.x * s + b.x + c.x, a.y * s + b.y + c.y, a.z * s + b.z + c.z, a.w * s + b.w + c.w);
}

// in-place 3-array RMW, grid-stride, UNROLL chunks per thread per iteration
template <int LD, int ST, int UNROLL>
__global__ void __launch_bounds__(256) rmw3(float4* a, float4* b, float4* c, long long n4, float s) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride * UNROLL) {
    float4 x[UNROLL], y[UNROLL], z[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long j = i + u * stride;
      if (j < n4) { x[u] = ld<LD>(a + j); y[u] = ld<LD>(b + j); z[u] = ld<LD>(c + j); }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long j = i + u * stride;
      if (j < n4) {
        st<ST>(a + j, upd(x[u], y[u], z[u], s));
        st<ST>(b + j, upd(y[u], z[u], x[u], s));
        st<ST>(c + j, upd(z[u], x[u], y[u], s));
      }
    }
  }
}
// out-of-place: read a,b,c write d,e,f
template <int LD, int ST, int UNROLL>
__global__ void __launch_bounds__(256) copy3(const float4* a, const float4* b, const float4* c, float4* d, float4* e,
                                             float4* f, long long n4, float s) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride * UNROLL) {
    float4 x[UNROLL], y[UNROLL], z[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long j = i + u * stride;
      if (j < n4) { x[u] = ld<LD>(a + j); y[u] = ld<LD>(b + j); z[u] = ld<LD>(c + j); }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long j = i + u * stride;
      if (j < n4) {
        st<ST>(d + j, upd(x[u], y[u], z[u], s));
        st<ST>(e + j, upd(y[u], z[u], x[u], s));
        st<ST>(f + j, upd(z[u], x[u], y[u], s));
      }
    }
  }
}
// block-contiguous variant: each block owns a contiguous span (better DRAM page locality per SM)
template <int LD, int ST>
__global__ void __launch_bounds__(256) rmw3_span(float4* a, float4* b, float4* c, long long n4, float s) {
  const long long per = (n4 + gridDim.x - 1) / gridDim.x;
  const long long lo = per * blockIdx.x, hi = min(n4, lo + per);
  for (long long i = lo + threadIdx.x; i < hi; i += 512) {
    const long long j = i + 256;
    float4 x0 = ld<LD>(a + i), y0 = ld<LD>(b + i), z0 = ld<LD>(c + i), x1, y1, z1;
    const bool two = j < hi;
    if (two) { x1 = ld<LD>(a + j); y1 = ld<LD>(b + j); z1 = ld<LD>(c + j); }
    st<ST>(a + i, upd(x0, y0, z0, s)); st<ST>(b + i, upd(y0, z0, x0, s)); st<ST>(c + i, upd(z0, x0, y0, s));
    if (two) { st<ST>(a + j, upd(x1, y1, z1, s)); st<ST>(b + j, upd(y1, z1, x1, s)); st<ST>(c + j, upd(z1, x1, y1, s)); }
  }
}

// 256-bit accesses (sm_100: ld/st.global.v8.b32)
struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld8(const float* p) {
  f8 r;
  asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}
__device__ __forceinline__ void st8(float* p, const f8& r) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]),
               "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
}
__global__ void __launch_bounds__(256) rmw3_v8(float* a, float* b, float* c, long long n8, float s) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    f8 x = ld8(a + i * 8), y = ld8(b + i * 8), z = ld8(c + i * 8), o1, o2, o3;
#pragma unroll
    for (int k = 0; k < 8; ++k) { o1.v[k] = x.v[k] * s + y.v[k] + z.v[k]; o2.v[k] = y.v[k] * s + z.v[k] + x.v[k]; o3.v[k] = z.v[k] * s + x.v[k] + y.v[k]; }
    st8(a + i * 8, o1); st8(b + i * 8, o2); st8(c + i * 8, o3);
  }
}

// bisecting the real dense-update kernel: MATH = Adam arithmetic, SUMSQ = block reduction + f64 RED,
// FLAGS = per-row stamp lookup + conditional gradient read/zero (row = 32 chunks)
template <int MATH, int SUMSQ, int FLAGS>
__global__ void __launch_bounds__(256) adam_probe(float4* th, float4* s1, float4* s2, float4* g4, const unsigned* flags,
                                                  long long n4, float c0, float c1, float c2, float c3, float l2,
                                                  double* acc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float sumsq = 0.f;
  if (i < n4) {
    float4 p = th[i], a = s1[i], b = s2[i];
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (FLAGS == 1) {
      const unsigned row = (unsigned)(i * 4) / 128u;
      if (__ldg(flags + row) == 7u) { g = g4[i]; g4[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
    } else if (FLAGS == 2) {          // lookup only
      const unsigned row = (unsigned)(i * 4) / 128u;
      if (__ldg(flags + row) == 7u) g.x = 1.f;
    } else if (FLAGS == 3) {          // conditional read, no zeroing
      const unsigned row = (unsigned)(i * 4) / 128u;
      if (__ldg(flags + row) == 7u) g = g4[i];
    } else if (FLAGS == 4) {          // unconditional 4th stream read + zero
      g = g4[i]; g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (FLAGS == 5) {          // conditional read; zeroing deferred to after the main stores
      const unsigned row = (unsigned)(i * 4) / 128u;
      if (__ldg(flags + row) == 7u) g = __ldcg(g4 + i);
    }
    float pv[4] = {p.x, p.y, p.z, p.w}, v1[4] = {a.x, a.y, a.z, a.w}, v2[4] = {b.x, b.y, b.z, b.w};
    const float gv[4] = {g.x, g.y, g.z, g.w};
    if (SUMSQ) sumsq = p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gv[j] + l2 * pv[j];
      if (MATH == 1) {
        const float m = c1 * v1[j] + (1.f - c1) * gj;
        const float v = c2 * v2[j] + (1.f - c2) * gj * gj;
        float sq; asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v));
        pv[j] = pv[j] - __fdividef(c0 * m, sq + c3);
        v1[j] = m; v2[j] = v;
      } else if (MATH == 2) {
        const float m = c1 * v1[j] + (1.f - c1) * gj;
        const float v = c2 * v2[j] + (1.f - c2) * gj * gj;
        pv[j] = pv[j] - c0 * m / (sqrtf(v) + c3);
        v1[j] = m; v2[j] = v;
      } else {
        pv[j] += gj; v1[j] += gj; v2[j] += gj;
      }
    }
    th[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
    s1[i] = make_float4(v1[0], v1[1], v1[2], v1[3]);
    s2[i] = make_float4(v2[0], v2[1], v2[2], v2[3]);
    if (FLAGS == 5) {
      const unsigned row = (unsigned)(i * 4) / 128u;
      if (__ldg(flags + row) == 7u) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (SUMSQ) {
    __shared__ double sp[8];
    double d = (double)sumsq;
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0; for (int w = 0; w < 8; ++w) t += sp[w];
      atomicAdd(acc + (blockIdx.x & 63), t);
    }
  }
}

template <typename F>
float time_ms(F f, int reps = 20) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / reps;
}

int main() {
  const long long n = 19216512LL;   // cfg2 parameter count
  const long long n4 = n / 4;
  float4 *buf[6];
  for (int i = 0; i < 6; ++i) { CK(cudaMalloc(&buf[i], n * 4)); CK(cudaMemset(buf[i], 0, n * 4)); }
  const double bytes = 24.0 * n;
  auto report = [&](const char* name, float ms) { printf("%-44s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, bytes / ms / 1e6); };
  int sms = 148;
  for (int bps : {2, 4, 8}) {
    const int g = sms * bps;
    char nm[96];
#define RUN(K, label) snprintf(nm, sizeof nm, "%s grid=%dxSM", label, bps); report(nm, time_ms([&] { K; }));
    RUN((rmw3<0, 0, 1><<<g, 256>>>(buf[0], buf[1], buf[2], n4, 0.5f)), "rmw3 ld/st default U1");
    RUN((rmw3<0, 0, 2><<<g, 256>>>(buf[0], buf[1], buf[2], n4, 0.5f)), "rmw3 ld/st default U2");
    RUN((rmw3<0, 0, 4><<<g, 256>>>(buf[0], buf[1], buf[2], n4, 0.5f)), "rmw3 ld/st default U4");
    RUN((rmw3<1, 1, 2><<<g, 256>>>(buf[0], buf[1], buf[2], n4, 0.5f)), "rmw3 ldcs/stcs U2");
    RUN((rmw3<1, 0, 2><<<g, 256>>>(buf[0], buf[1], buf[2], n4, 0.5f)), "rmw3 ldcs/st default U2");
    RUN((rmw3<2, 2, 2><<<g, 256>>>(buf[0], buf[1], buf[2], n4, 0.5f)), "rmw3 ldcg/stcg U2");
    RUN((rmw3<3, 3, 2><<<g, 256>>>(buf[0], buf[1], buf[2], n4, 0.5f)), "rmw3 evict_first/stwt U2");
    RUN((copy3<0, 0, 2><<<g, 256>>>(buf[0], buf[1], buf[2], buf[3], buf[4], buf[5], n4, 0.5f)), "copy3 default U2");
    RUN((copy3<1, 1, 2><<<g, 256>>>(buf[0], buf[1], buf[2], buf[3], buf[4], buf[5], n4, 0.5f)), "copy3 ldcs/stcs U2");
    RUN((copy3<0, 0, 4><<<g, 256>>>(buf[0], buf[1], buf[2], buf[3], buf[4], buf[5], n4, 0.5f)), "copy3 default U4");
    RUN((rmw3_v8<<<g, 256>>>((float*)buf[0], (float*)buf[1], (float*)buf[2], n / 8, 0.5f)), "rmw3 256-bit ld/st");
    RUN((rmw3_span<0, 0><<<g, 256>>>(buf[0], buf[1], buf[2], n4, 0.5f)), "rmw3_span default");
    RUN((rmw3_span<1, 1><<<g, 256>>>(buf[0], buf[1], buf[2], n4, 0.5f)), "rmw3_span ldcs/stcs");
  }
  {
    const int g = (int)((n4 + 255) / 256);
    report("rmw3 default U1 one-chunk-per-thread", time_ms([&] { rmw3<0, 0, 1><<<g, 256>>>(buf[0], buf[1], buf[2], n4, 0.5f); }));
    report("copy3 default U1 one-chunk-per-thread", time_ms([&] { copy3<0, 0, 1><<<g, 256>>>(buf[0], buf[1], buf[2], buf[3], buf[4], buf[5], n4, 0.5f); }));
  }
  {
    unsigned* flags; double* acc;
    const long long rows = n / 128;
    CK(cudaMalloc(&flags, rows * 4)); CK(cudaMalloc(&acc, 64 * 8)); CK(cudaMemset(acc, 0, 64 * 8));
    std::vector<unsigned> hf(rows);
    for (long long r = 0; r < rows; ++r) hf[r] = (r * 2654435761u >> 16) % 100 < 35 ? 7u : 0u;   // 35% of rows touched
    CK(cudaMemcpy(flags, hf.data(), rows * 4, cudaMemcpyHostToDevice));
    const int g = (int)((n4 + 255) / 256);
#define AP(M, S, F, label) report(label, time_ms([&] { adam_probe<M, S, F><<<g, 256>>>(buf[0], buf[1], buf[2], buf[3], flags, n4, 1e-3f, .9f, .999f, 1e-8f, 1e-6f, acc); }));
    AP(0, 0, 0, "one-shot: add only");
    AP(1, 0, 0, "one-shot: adam approx math");
    AP(2, 0, 0, "one-shot: adam IEEE math");
    AP(1, 1, 0, "one-shot: adam approx + sumsq");
    AP(1, 0, 1, "one-shot: adam approx + flags/G(35%)");
    AP(1, 1, 1, "one-shot: adam approx + sumsq + flags/G");
    AP(1, 0, 2, "one-shot: flags lookup only");
    AP(1, 0, 3, "one-shot: flags + conditional G read");
    AP(1, 0, 4, "one-shot: unconditional G read+zero");
    AP(1, 0, 5, "one-shot: cond G ldcg, zero after stores");
    AP(2, 1, 1, "one-shot: adam IEEE + sumsq + flags/G");
  }
  // plain device-to-device memcpy of the same volume for reference (3 arrays)
  report("cudaMemcpyAsync D2D x3 (12 B/elem r+w... x2)", time_ms([&] { for (int i = 0; i < 3; ++i) cudaMemcpyAsync(buf[3 + i], buf[i], n * 4, cudaMemcpyDeviceToDevice); }));
  return 0;
}
