// Row kernels of the log-linear model (sert.models.LanguageModel, sert/models.py:804-878):
// per-word softmax statistics, ProductTimestepsLayer (sum_w log clip(p) then softmax, :186-212),
// clipped categorical cross-entropy against CSR labels (:289-292) and the backward of all of it in
// the general (clipped) regime.  The (B*W,E) logits Z come from the
// word x entity projection GEMM; everything here is a streaming pass over Z / S with block-per-row
// reductions (warp shuffles + one smem hop).
#include "ll_kernels.cuh"

#include <cuda_bf16.h>

namespace sert {

__device__ __forceinline__ float block_max(float v, float *sm) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : -INFINITY;
  r = warp_max(r);
  __syncthreads();
  return r;   // valid in every thread of warp 0; broadcast below
}
__device__ __forceinline__ float block_sum(float v, float *sm) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.f;
  r = warp_sum(r);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_bcast(float v, float *sm) {
  if (threadIdx.x == 0) sm[0] = v;
  __syncthreads();
  const float r = sm[0];
  __syncthreads();
  return r;
}

// ---- per-row softmax statistics: rmax[r] = max_e Z[r,e], rsum[r] = sum_e exp(Z[r,e]-rmax[r]) ----
__global__ void __launch_bounds__(256) ll_row_stats_kernel(const float *__restrict__ Z, long long rows,
                                                           int E, long long ldz, float *__restrict__ rmax,
                                                           float *__restrict__ rsum) {
  __shared__ float sm[32];
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float *z = Z + r * ldz;
    float m = -INFINITY;
    for (int e = threadIdx.x; e < E; e += blockDim.x) m = fmaxf(m, z[e]);
    m = block_bcast(block_max(m, sm), sm);
    float s = 0.f;
    for (int e = threadIdx.x; e < E; e += blockDim.x) s += expf(z[e] - m);
    s = block_sum(s, sm);
    if (threadIdx.x == 0) { rmax[r] = m; rsum[r] = s; }
  }
}

// folds the per-slice (max, sum exp) pairs the GEMM epilogue left (gemm_tc.cuh: row_stats), one warp per row
__global__ void __launch_bounds__(256) ll_combine_slices_kernel(const float2 *__restrict__ stats, long long rows,
                                                                int slots, long long ld, float *__restrict__ rmax,
                                                                float *__restrict__ rsum) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float2 *row = stats + r * ld;
  float m = -INFINITY;
  for (int i = lane; i < slots; i += 32) m = fmaxf(m, row[i].x);
  m = warp_max(m);
  float s = 0.f;
  for (int i = lane; i < slots; i += 32) {
    const float2 v = row[i];
    s += v.y * expf(v.x - m);
  }
  s = warp_sum(s);
  if (lane == 0) { rmax[r] = m; rsum[r] = s; }
}

int launch_ll_combine_slices(const float2 *stats, int64_t rows, int slots, int64_t ld, float *rmax, float *rsum,
                             cudaStream_t st) {
  if (rows == 0) return 0;
  ll_combine_slices_kernel<<<cdiv(rows, 8), 256, 0, st>>>(stats, rows, slots, ld, rmax, rsum);
  SERT_LAUNCH_CHECK();
  return 0;
}

int launch_ll_row_stats(const float *Z, int64_t rows, int E, int64_t ldz, float *rmax, float *rsum,
                        cudaStream_t st) {
  if (rows == 0) return 0;
  const int threads = E >= 1024 ? 256 : 128;
  ll_row_stats_kernel<<<(int)std::min<int64_t>(rows, 148 * 64), threads, 0, st>>>(Z, rows, E, ldz, rmax, rsum);
  SERT_LAUNCH_CHECK();
  return 0;
}

// ---- in-place row softmax (predict_fn: unclipped per-word distributions, sert/models.py:868,880-890) --
__global__ void __launch_bounds__(256) ll_softmax_inplace_kernel(float *__restrict__ Z, long long rows, int E,
                                                                 long long ldz, const float *__restrict__ rmax,
                                                                 const float *__restrict__ rsum) {
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    float *z = Z + r * ldz;
    const float m = rmax[r], s = rsum[r];
    for (int e = threadIdx.x; e < E; e += blockDim.x) z[e] = expf(z[e] - m) / s;
  }
}

int launch_ll_softmax_inplace(float *Z, int64_t rows, int E, int64_t ldz, const float *rmax,
                              const float *rsum, cudaStream_t st) {
  if (rows == 0) return 0;
  ll_softmax_inplace_kernel<<<(int)std::min<int64_t>(rows, 148 * 64), 256, 0, st>>>(Z, rows, E, ldz, rmax, rsum);
  SERT_LAUNCH_CHECK();
  return 0;
}

// ---- joint logits S[i,e] = sum_w log clip(p[i,w,e]) ------------------------------------------------
__global__ void ll_log_kernel(const float *__restrict__ in, float *__restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = logf(in[i]);
}

__global__ void __launch_bounds__(256) ll_joint_kernel(const float *__restrict__ Z,
                                                       const float *__restrict__ rmax,
                                                       const float *__restrict__ lrsum, float *__restrict__ S,
                                                       int B, int W, int E, long long ldz, long long lds) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (e >= E) return;
  // log clip(p, lo, hi) = clip(log p, log lo, log hi) (log is monotone), and log p = z - max - log(sum): the W terms
  // of an entity cost a subtraction and two compares each instead of exp, divide and log (4.6 -> 1.6 ms at
  // BASELINE configs[4], where this pass is pure streaming of the 8.2 GB logit matrix).  The two forms differ by the
  // rounding of exp / divide / log (~1e-7 relative on log p), far inside the 1e-4 parity tolerance.
  const float log_lo = logf(SERT_CLIP_LO), log_hi = logf(SERT_CLIP_HI);
  float acc = 0.f;
  for (int w = 0; w < W; ++w) {
    const long long r = (long long)i * W + w;
    const float lp = (Z[r * ldz + e] - rmax[r]) - __ldg(lrsum + r);
    acc += fminf(fmaxf(lp, log_lo), log_hi);
  }
  S[(long long)i * lds + e] = acc;
}

// The same pass, four entities per thread (16-byte streaming loads of Z, every row of the window in flight), one CTA
// per (instance, slot of kJointSlot entities); optionally leaves the slot's (max, sum exp) of the joint logits.
__global__ void __launch_bounds__(256) ll_joint4_kernel(const float *__restrict__ Z, const float *__restrict__ rmax,
                                                        const float *__restrict__ lrsum, float *__restrict__ S,
                                                        int W, int E, long long ldz, long long lds,
                                                        float2 *__restrict__ sstats, int slots) {
  __shared__ float sm[32];
  const int e = (blockIdx.x * 256 + threadIdx.x) * 4;
  const int i = blockIdx.y;
  const float log_lo = logf(SERT_CLIP_LO), log_hi = logf(SERT_CLIP_HI);
  const bool ok = e < E;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
    const float4 *z = reinterpret_cast<const float4 *>(Z + (long long)i * W * ldz + e);
    const long long ld4 = ldz >> 2;
#pragma unroll 5
    for (int w = 0; w < W; ++w) {
      const long long r = (long long)i * W + w;
      const float4 v = __ldcs(z + w * ld4);
      const float mx = __ldg(rmax + r), ls = __ldg(lrsum + r);
      acc.x += fminf(fmaxf((v.x - mx) - ls, log_lo), log_hi);
      acc.y += fminf(fmaxf((v.y - mx) - ls, log_lo), log_hi);
      acc.z += fminf(fmaxf((v.z - mx) - ls, log_lo), log_hi);
      acc.w += fminf(fmaxf((v.w - mx) - ls, log_lo), log_hi);
    }
    *reinterpret_cast<float4 *>(S + (long long)i * lds + e) = acc;
  }
  if (sstats != nullptr) {
    float m = ok ? fmaxf(fmaxf(acc.x, acc.y), fmaxf(acc.z, acc.w)) : -INFINITY;
    m = block_bcast(block_max(m, sm), sm);
    float sum = ok ? expf(acc.x - m) + expf(acc.y - m) + expf(acc.z - m) + expf(acc.w - m) : 0.f;
    sum = block_sum(sum, sm);
    if (threadIdx.x == 0) sstats[(long long)i * slots + blockIdx.x] = make_float2(m, sum);
  }
}

int launch_ll_joint(const float *Z, const float *rmax, const float *rsum, float *S, int B, int W, int E,
                    int64_t ldz, int64_t lds, cudaStream_t st, float *lrsum_scratch, float2 *sstats) {
  if (B == 0) return 0;
  SERT_REQUIRE(lrsum_scratch != nullptr, "the joint pass needs B*W floats of scratch");
  ll_log_kernel<<<cdiv((long long)B * W, 256), 256, 0, st>>>(rsum, lrsum_scratch, (long long)B * W);
  SERT_LAUNCH_CHECK();
  if ((E & 3) == 0 && (ldz & 3) == 0 && (lds & 3) == 0 && (reinterpret_cast<uintptr_t>(Z) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(S) & 15) == 0) {
    static_assert(kJointSlot == 256 * 4, "one CTA of 256 threads x 4 entities per slot");
    const int slots = ll_joint_slots(E);
    dim3 grid4(slots, B);
    ll_joint4_kernel<<<grid4, 256, 0, st>>>(Z, rmax, lrsum_scratch, S, W, E, ldz, lds, sstats, slots);
    SERT_LAUNCH_CHECK();
    return sstats != nullptr ? 1 : 0;
  }
  dim3 grid(cdiv(E, 256), B);
  ll_joint_kernel<<<grid, 256, 0, st>>>(Z, rmax, lrsum_scratch, S, B, W, E, ldz, lds);
  SERT_LAUNCH_CHECK();
  return 0;
}

// ---- per-instance softmax over S, CSR cross-entropy and (train) ds = dL/dS ------------------------
__global__ void __launch_bounds__(256) ll_instance_kernel(LlInstanceArgs a) {
  __shared__ float sm[32];
  const int i = blockIdx.x;
  const float *s = a.S + (long long)i * a.lds;
  float m = -INFINITY, sum = 0.f;
  if (a.sstats != nullptr) {
    // the joint pass left (max, sum exp) per slot of the row: fold them instead of reading the row twice
    const float2 *st = a.sstats + (long long)i * a.slots;
    for (int b = threadIdx.x; b < a.slots; b += blockDim.x) m = fmaxf(m, st[b].x);
    m = block_bcast(block_max(m, sm), sm);
    for (int b = threadIdx.x; b < a.slots; b += blockDim.x) {
      const float2 v = st[b];
      sum += v.y * expf(v.x - m);
    }
    sum = block_bcast(block_sum(sum, sm), sm);
  } else {
    for (int e = threadIdx.x; e < a.E; e += blockDim.x) m = fmaxf(m, s[e]);
    m = block_bcast(block_max(m, sm), sm);
    for (int e = threadIdx.x; e < a.E; e += blockDim.x) sum += expf(s[e] - m);
    sum = block_bcast(block_sum(sum, sm), sm);
  }

  const long long p0 = a.indptr[i] - a.nnz_base, p1 = a.indptr[i + 1] - a.nnz_base;
  const float wi = a.w ? a.w[i] : 1.0f;
  const float cw = a.train ? wi * a.inv_B : 0.f;
  float ell = 0.f, adot = 0.f;
  for (long long p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
    const int e = a.indices[p];
    const float y = a.data[p];
    const float o = expf(s[e] - m) / sum;
    const float c = clipf_(o, SERT_CLIP_LO, SERT_CLIP_HI);
    ell -= y * logf(c);
    if (o >= SERT_CLIP_LO && o <= SERT_CLIP_HI) adot += (-cw * y / c) * o;   // sum_e do*o
  }
  ell = block_sum(ell, sm);
  adot = block_bcast(block_sum(adot, sm), sm);
  if (threadIdx.x == 0) {
    if (a.ell_out) a.ell_out[i] = ell;
    atomicAdd(a.loss_acc, (double)(a.train ? wi * ell : ell));
  }
  if (!a.train) return;
  // ds = o * (do - sum_e do*o); do is non-zero only on the label columns
  float *ds = a.DS + (long long)i * a.lds;
  if ((a.E & 3) == 0 && (a.lds & 3) == 0 && (reinterpret_cast<uintptr_t>(a.S) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(a.DS) & 15) == 0) {
    const float4 *s4 = reinterpret_cast<const float4 *>(s);
    float4 *d4 = reinterpret_cast<float4 *>(ds);
    for (int e = threadIdx.x; e < (a.E >> 2); e += blockDim.x) {
      const float4 v = __ldcs(s4 + e);
      d4[e] = make_float4(-(expf(v.x - m) / sum) * adot, -(expf(v.y - m) / sum) * adot, -(expf(v.z - m) / sum) * adot,
                          -(expf(v.w - m) / sum) * adot);
    }
  } else {
    for (int e = threadIdx.x; e < a.E; e += blockDim.x) ds[e] = -(expf(s[e] - m) / sum) * adot;
  }
  __syncthreads();
  for (long long p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
    const int e = a.indices[p];
    const float y = a.data[p];
    const float o = expf(s[e] - m) / sum;
    const float c = clipf_(o, SERT_CLIP_LO, SERT_CLIP_HI);
    if (o >= SERT_CLIP_LO && o <= SERT_CLIP_HI) atomicAdd(ds + e, o * (-cw * y / c));
  }
}

int launch_ll_instance(const LlInstanceArgs &a, cudaStream_t st) {
  if (a.B == 0) return 0;
  SERT_REQUIRE(!a.train || a.DS != nullptr, "training needs a dS buffer");
  ll_instance_kernel<<<a.B, a.E >= 1024 ? 256 : 128, 0, st>>>(a);
  SERT_LAUNCH_CHECK();
  return 0;
}

// ---- dZ in place: one block per (i,w) row ---------------------------------------------------------
// MODE 0: both halves (single device).  Entity-sharded step: MODE 1 writes the shard's partial row sum
// acc[r] = sum_{e local, p unclipped} ds[e] (all-reduced by the caller), MODE 2 applies dz = dp*p - p*acc[r].
template <int MODE>
__global__ void __launch_bounds__(256) ll_dz_kernel(float *__restrict__ Z, const float *__restrict__ rmax,
                                                    const float *__restrict__ rsum,
                                                    const float *__restrict__ DS, long long rows, int W, int E,
                                                    long long ldz, long long lds, float *__restrict__ racc) {
  __shared__ float sm[32];
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    float *z = Z + r * ldz;
    const float *ds = DS + (r / W) * lds;
    const float m = rmax[r], s = rsum[r];
    float acc = 0.f;
    if (MODE != 2) {
      for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const float p = expf(z[e] - m) / s;
        if (p >= SERT_CLIP_LO && p <= SERT_CLIP_HI) acc += ds[e];   // dp*p = ds/clip(p)*p = ds when unclipped
      }
      acc = block_bcast(block_sum(acc, sm), sm);
      if (MODE == 1) {
        if (threadIdx.x == 0) racc[r] = acc;
        continue;
      }
    } else {
      acc = racc[r];
    }
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      const float p = expf(z[e] - m) / s;
      const float dpp = (p >= SERT_CLIP_LO && p <= SERT_CLIP_HI) ? ds[e] : 0.f;
      z[e] = dpp - p * acc;
    }
    __syncthreads();
  }
}

int launch_ll_dz(float *Z, const float *rmax, const float *rsum, const float *DS, int B, int W, int E,
                 int64_t ldz, int64_t lds, cudaStream_t st, int mode, float *racc) {
  const long long rows = (long long)B * W;
  if (rows == 0) return 0;
  SERT_REQUIRE(mode == 0 || racc != nullptr, "sharded dZ needs the row-sum buffer");
  const int grid = (int)std::min<long long>(rows, 148 * 64), threads = E >= 1024 ? 256 : 128;
  if (mode == 0) ll_dz_kernel<0><<<grid, threads, 0, st>>>(Z, rmax, rsum, DS, rows, W, E, ldz, lds, racc);
  else if (mode == 1) ll_dz_kernel<1><<<grid, threads, 0, st>>>(Z, rmax, rsum, DS, rows, W, E, ldz, lds, racc);
  else ll_dz_kernel<2><<<grid, threads, 0, st>>>(Z, rmax, rsum, DS, rows, W, E, ldz, lds, racc);
  SERT_LAUNCH_CHECK();
  return 0;
}

// ---- fused backward tail of the tensor-core path ---------------------------------------------------------------
// dZ = dpp - p * acc_r (acc_r = sum over the row's unclipped entries of ds) never exists in float32: pass A streams
// Z once for acc_r, pass B streams it once more and writes dZ directly as the two bf16x3 split operands the gradient
// GEMMs consume -- dZs (B*W, 3*E64), K-major A of dX = dZ . Wd^T, and dZT_s (E, 3*BW64), K-major B of gWd = X^T . dZ
// (transposed through shared memory).  Replaces ll_dz (2 reads + 1 write of Z) + split_bf16 + split_bf16_t + colsum
// (3 more reads of the float32 dZ): 30 ms -> 9 ms at BASELINE configs[4].  The clip mask is evaluated in the log
// domain, log p = z - max - log(sum) against log(1e-7) / log(1 - 1e-7), like ll_joint_kernel.
__global__ void __launch_bounds__(256) ll_racc_log_kernel(const float *__restrict__ Z, const float *__restrict__ rmax,
                                                          const float *__restrict__ lrsum,
                                                          const float *__restrict__ DS, long long rows, int W, int E,
                                                          long long ldz, long long lds, float *__restrict__ racc) {
  __shared__ float sm[32];
  const float log_lo = logf(SERT_CLIP_LO), log_hi = logf(SERT_CLIP_HI);
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float *z = Z + r * ldz;
    const float *ds = DS + (r / W) * lds;
    const float off = rmax[r] + lrsum[r];
    float acc = 0.f;
    if ((E & 3) == 0 && (ldz & 3) == 0 && (lds & 3) == 0) {
      const float4 *z4 = reinterpret_cast<const float4 *>(z);
      const float4 *d4 = reinterpret_cast<const float4 *>(ds);
      for (int e = threadIdx.x; e < (E >> 2); e += blockDim.x) {
        const float4 zv = __ldcs(z4 + e);
        const float4 dv = __ldg(d4 + e);
        const float l0 = zv.x - off, l1 = zv.y - off, l2 = zv.z - off, l3 = zv.w - off;
        acc += (l0 >= log_lo && l0 <= log_hi) ? dv.x : 0.f;
        acc += (l1 >= log_lo && l1 <= log_hi) ? dv.y : 0.f;
        acc += (l2 >= log_lo && l2 <= log_hi) ? dv.z : 0.f;
        acc += (l3 >= log_lo && l3 <= log_hi) ? dv.w : 0.f;
      }
    } else {
      for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const float lp = z[e] - off;
        acc += (lp >= log_lo && lp <= log_hi) ? ds[e] : 0.f;
      }
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) racc[r] = acc;
  }
}

int launch_ll_racc_log(const float *Z, const float *rmax, const float *lrsum, const float *DS, int B, int W, int E,
                       int64_t ldz, int64_t lds, float *racc, cudaStream_t st) {
  const long long rows = (long long)B * W;
  if (rows == 0) return 0;
  ll_racc_log_kernel<<<(int)std::min<long long>(rows, 148 * 64), 256, 0, st>>>(Z, rmax, lrsum, DS, rows, W, E, ldz, lds,
                                                                               racc);
  SERT_LAUNCH_CHECK();
  return 0;
}

constexpr int kDzTile = 64;      // rows x columns of Z per CTA in pass B

__device__ __forceinline__ unsigned short bf16_bits(float x) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}

// grid (ceil(E64 / 64), ceil(BW64 / 64)); dZs row stride terms*E64, dZT_s row stride terms*BW64 (elements);
// terms 3: A rows [hi | hi | mid], B rows [hi | mid | hi]; terms 2 (pair operands, gemm_tc.cuh): [hi | mid] for both
__global__ void __launch_bounds__(256) ll_dz_split_kernel(const float *__restrict__ Z, const float *__restrict__ rmax,
                                                          const float *__restrict__ lrsum,
                                                          const float *__restrict__ racc,
                                                          const float *__restrict__ DS, long long rows, int W, int E,
                                                          long long ldz, long long lds, long long E64, long long BW64,
                                                          int terms, __nv_bfloat16 *__restrict__ dZs,
                                                          __nv_bfloat16 *__restrict__ dZT_s) {
  __shared__ unsigned short t_hi[kDzTile][kDzTile + 2], t_mid[kDzTile][kDzTile + 2];   // [column][row]
  const float log_lo = logf(SERT_CLIP_LO), log_hi = logf(SERT_CLIP_HI);
  const long long c0 = (long long)blockIdx.x * kDzTile, r0 = (long long)blockIdx.y * kDzTile;
  const int tc = (threadIdx.x & 15) * 4;          // 16 threads x 4 columns per row
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int tr = (threadIdx.x >> 4) + 16 * k;
    const long long r = r0 + tr, c = c0 + tc;
    float dz[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < rows && c < E) {
      const float off = rmax[r] + lrsum[r], acc = racc[r];
      const float *z = Z + r * ldz + c;
      const float *ds = DS + (r / W) * lds + c;
      float zv[4], dv[4];
      if (c + 3 < E && (ldz & 3) == 0 && (lds & 3) == 0) {
        const float4 a = __ldcs(reinterpret_cast<const float4 *>(z));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(ds));
        zv[0] = a.x; zv[1] = a.y; zv[2] = a.z; zv[3] = a.w;
        dv[0] = b.x; dv[1] = b.y; dv[2] = b.z; dv[3] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          zv[j] = c + j < E ? z[j] : 0.f;
          dv[j] = c + j < E ? ds[j] : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lp = zv[j] - off;
        const float pj = __expf(lp);
        const float dpp = (lp >= log_lo && lp <= log_hi) ? dv[j] : 0.f;
        dz[j] = c + j < E ? dpp - pj * acc : 0.f;
      }
    }
    unsigned short hi[4], mid[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat16 h = __float2bfloat16_rn(dz[j]);
      hi[j] = __bfloat16_as_ushort(h);
      mid[j] = bf16_bits(dz[j] - __bfloat162float(h));
      if (dZT_s != nullptr) {
        t_hi[tc + j][tr] = hi[j];
        t_mid[tc + j][tr] = mid[j];
      }
    }
    if (r < BW64 && c < E64) {
      // A operand rows [hi | hi | mid]
      const uint2 h2 = make_uint2((unsigned)hi[0] | ((unsigned)hi[1] << 16), (unsigned)hi[2] | ((unsigned)hi[3] << 16));
      const uint2 m2 = make_uint2((unsigned)mid[0] | ((unsigned)mid[1] << 16), (unsigned)mid[2] | ((unsigned)mid[3] << 16));
      __nv_bfloat16 *row = dZs + r * terms * E64 + c;
      if (r < rows) {
        *reinterpret_cast<uint2 *>(row) = h2;
        if (terms == 2) {
          *reinterpret_cast<uint2 *>(row + E64) = m2;
        } else {
          *reinterpret_cast<uint2 *>(row + E64) = h2;
          *reinterpret_cast<uint2 *>(row + 2 * E64) = m2;
        }
      }
    }
  }
  if (dZT_s == nullptr) return;      // the gradient GEMM reads dZs N-major (launch_gemm_tc_pair_bn): no transposed copy
  __syncthreads();
  // B operand rows [hi | mid | hi] of the transpose: thread = (column, 16-row chunk)
  const int col = threadIdx.x >> 2, rc = (threadIdx.x & 3) * 16;
  const long long c = c0 + col;
  if (c < E && r0 + rc < BW64) {
    unsigned int hw[8], mw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hw[j] = (unsigned)t_hi[col][rc + 2 * j] | ((unsigned)t_hi[col][rc + 2 * j + 1] << 16);
      mw[j] = (unsigned)t_mid[col][rc + 2 * j] | ((unsigned)t_mid[col][rc + 2 * j + 1] << 16);
    }
    __nv_bfloat16 *dst = dZT_s + c * terms * BW64 + r0 + rc;
    uint4 *d0 = reinterpret_cast<uint4 *>(dst), *d1 = reinterpret_cast<uint4 *>(dst + BW64);
    d0[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); d0[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
    d1[0] = make_uint4(mw[0], mw[1], mw[2], mw[3]); d1[1] = make_uint4(mw[4], mw[5], mw[6], mw[7]);
    if (terms == 3) {
      uint4 *d2 = reinterpret_cast<uint4 *>(dst + 2 * BW64);
      d2[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); d2[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
    }
  }
}

int launch_ll_dz_split(const float *Z, const float *rmax, const float *lrsum, const float *racc, const float *DS,
                       int B, int W, int E, int64_t ldz, int64_t lds, int terms, __nv_bfloat16 *dZs, __nv_bfloat16 *dZT_s,
                       cudaStream_t st) {
  SERT_REQUIRE(terms == 2 || terms == 3, "dZ operands: 2 (pair) or 3 blocks");
  const long long rows = (long long)B * W;
  if (rows == 0) return 0;
  const long long E64 = (long long)align_up((size_t)E, 64), BW64 = (long long)align_up((size_t)rows, 64);
  dim3 grid((unsigned)(E64 / kDzTile), (unsigned)(BW64 / kDzTile));
  SERT_REQUIRE(grid.y < 65536, "too many rows for the dZ tile grid");
  ll_dz_split_kernel<<<grid, 256, 0, st>>>(Z, rmax, lrsum, racc, DS, rows, W, E, ldz, lds, E64, BW64, terms, dZs, dZT_s);
  SERT_LAUNCH_CHECK();
  return 0;
}

// =================================================================================================
// Entity-sharded (column-parallel) softmax pieces, SURVEY.md 8(e): every rank holds the columns
// [e_begin, e_begin+E) of Wd / bd and therefore of Z and S.  Row statistics are computed per shard,
// gathered, and combined here; label terms are computed by the rank that owns the label's column.
// =================================================================================================

// parts: [shard][2][rows] = per-shard (row max, row sum of exp(z - shard max)) -> global max / sum
__global__ void ll_combine_stats_kernel(const float *__restrict__ parts, int shards, long long rows,
                                        float *__restrict__ rmax, float *__restrict__ rsum) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float m = -INFINITY;
  for (int s = 0; s < shards; ++s) m = fmaxf(m, parts[(2ll * s) * rows + r]);
  float sum = 0.f;
  for (int s = 0; s < shards; ++s)
    sum += parts[(2ll * s + 1) * rows + r] * expf(parts[(2ll * s) * rows + r] - m);
  rmax[r] = m;
  rsum[r] = sum;
}

int launch_ll_combine_stats(const float *parts, int shards, int64_t rows, float *rmax, float *rsum,
                            cudaStream_t st) {
  if (rows == 0) return 0;
  ll_combine_stats_kernel<<<cdiv(rows, 256), 256, 0, st>>>(parts, shards, rows, rmax, rsum);
  SERT_LAUNCH_CHECK();
  return 0;
}

// One warp per instance: the label terms of the rows' cross-entropy that fall into this shard.
// adot_out[i] = sum_{labels e in shard, o unclipped} (-cw*y/clip(o)) * o  (partial of sum_e do*o)
__global__ void __launch_bounds__(256) ll_shard_labels_kernel(LlInstanceArgs a, const float *__restrict__ smax,
                                                              const float *__restrict__ ssum, int e_begin,
                                                              float *__restrict__ adot_out) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= a.B) return;
  const float *s = a.S + (long long)i * a.lds;
  const float m = smax[i], sum = ssum[i];
  const long long p0 = a.indptr[i] - a.nnz_base, p1 = a.indptr[i + 1] - a.nnz_base;
  const float wi = a.w ? a.w[i] : 1.0f;
  const float cw = a.train ? wi * a.inv_B : 0.f;
  float ell = 0.f, adot = 0.f;
  for (long long p = p0 + lane; p < p1; p += 32) {
    const int e = a.indices[p] - e_begin;
    if (e < 0 || e >= a.E) continue;
    const float y = a.data[p];
    const float o = expf(s[e] - m) / sum;
    const float c = clipf_(o, SERT_CLIP_LO, SERT_CLIP_HI);
    ell -= y * logf(c);
    if (o >= SERT_CLIP_LO && o <= SERT_CLIP_HI) adot += (-cw * y / c) * o;
  }
  ell = warp_sum(ell);
  adot = warp_sum(adot);
  if (lane == 0) {
    if (a.ell_out) a.ell_out[i] = ell;
    if (adot_out) adot_out[i] = adot;
    if (ell != 0.f) atomicAdd(a.loss_acc, (double)(a.train ? wi * ell : ell));
  }
}

int launch_ll_shard_labels(const LlInstanceArgs &a, const float *smax, const float *ssum, int e_begin,
                           float *adot_out, cudaStream_t st) {
  if (a.B == 0) return 0;
  ll_shard_labels_kernel<<<cdiv(a.B, 8), 256, 0, st>>>(a, smax, ssum, e_begin, adot_out);
  SERT_LAUNCH_CHECK();
  return 0;
}

// ds[i,e] = o*(do - adot[i]) over the shard's columns, adot all-reduced over the shards.
__global__ void __launch_bounds__(256) ll_shard_ds_kernel(LlInstanceArgs a, const float *__restrict__ smax,
                                                          const float *__restrict__ ssum, int e_begin,
                                                          const float *__restrict__ adot_all) {
  const int i = blockIdx.x;
  const float *s = a.S + (long long)i * a.lds;
  float *ds = a.DS + (long long)i * a.lds;
  const float m = smax[i], sum = ssum[i], adot = adot_all[i];
  for (int e = threadIdx.x; e < a.E; e += blockDim.x) ds[e] = -(expf(s[e] - m) / sum) * adot;
  __syncthreads();
  const long long p0 = a.indptr[i] - a.nnz_base, p1 = a.indptr[i + 1] - a.nnz_base;
  const float wi = a.w ? a.w[i] : 1.0f;
  const float cw = wi * a.inv_B;
  for (long long p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
    const int e = a.indices[p] - e_begin;
    if (e < 0 || e >= a.E) continue;
    const float y = a.data[p];
    const float o = expf(s[e] - m) / sum;
    const float c = clipf_(o, SERT_CLIP_LO, SERT_CLIP_HI);
    if (o >= SERT_CLIP_LO && o <= SERT_CLIP_HI) atomicAdd(ds + e, o * (-cw * y / c));
  }
}

int launch_ll_shard_ds(const LlInstanceArgs &a, const float *smax, const float *ssum, int e_begin,
                       const float *adot_all, cudaStream_t st) {
  if (a.B == 0) return 0;
  SERT_REQUIRE(a.DS != nullptr, "training needs a dS buffer");
  ll_shard_ds_kernel<<<a.B, a.E >= 1024 ? 256 : 128, 0, st>>>(a, smax, ssum, e_begin, adot_all);
  SERT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sert
