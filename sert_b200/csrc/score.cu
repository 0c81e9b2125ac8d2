// Entity scoring: query x all-entities inner products with a fused running top-k.
// Replaces VectorSpaceCallback's sklearn k-NN / cdist + argsort (bin/query.py:241-318): on
// L2-normalised vectors, Euclidean k-NN order == inner-product order, so the device returns the
// top-k rows by inner product and the host callback re-scores them exactly like the reference.
//
// The (Q,E) score matrix is never materialised.  Entities are swept in chunks; the score-tile
// kernel compares each score with its query's running threshold tau_q (the k-th best so far) and
// appends survivors to a per-query candidate list; a prune kernel re-selects the best k whenever
// a list is more than half full and raises tau_q.  Chunk size == half the list capacity, so a list
// can never overflow (worst case: every score of a chunk survives) and the result is exact.
// Keys are 64-bit (order-preserving score bits << 32 | ~row id): larger key == better score, ties
// broken by lower row id, so selection and the final order are deterministic.
#include "score.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "comm.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "topk_keys.cuh"

namespace sert {

// ---- row L2 normalisation (bin/query.py:270-274, 333-336) --------------------------------------
__global__ void __launch_bounds__(256) normalise_rows_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                             long long rows, int d) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float *x = in + r * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s += x[c] * x[c];
  s = sqrtf(warp_sum(s));
  for (int c = lane; c < d; c += 32) out[r * d + c] = x[c] / s;
}

int launch_normalise_rows(const float *in, float *out, int64_t rows, int d, cudaStream_t st) {
  if (rows == 0) return 0;
  normalise_rows_kernel<<<cdiv(rows, 8), 256, 0, st>>>(in, out, rows, d);
  SERT_LAUNCH_CHECK();
  return 0;
}

// ---- score tile + threshold filter (fp32 FMA tile; the tcgen05 variant shares the epilogue) ------
constexpr int TQ = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) score_filter_kernel(const float *__restrict__ Qm,   // (Q,d)
                                                           const float *__restrict__ En,   // (rows,d)
                                                           int Q, long long n_begin, long long n_end, int d,
                                                           long long row_offset,
                                                           const unsigned long long *__restrict__ tau,
                                                           int *__restrict__ count,
                                                           unsigned long long *__restrict__ cand, int cap,
                                                           int *__restrict__ overflow) {
  __shared__ float Qs[TK][TQ + 4];
  __shared__ float Es[TK][TN + 4];
  const int tid = threadIdx.x;
  const int q0 = blockIdx.y * TQ;
  const long long n0 = n_begin + (long long)blockIdx.x * TN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < d; k0 += TK) {
#pragma unroll
    for (int l = 0; l < (TQ * TK) / 256; ++l) {
      const int e = tid + l * 256;
      const int k = e % TK, m = e / TK;
      const int gq = q0 + m, gk = k0 + k;
      Qs[k][m] = (gq < Q && gk < d) ? Qm[(size_t)gq * d + gk] : 0.f;
    }
#pragma unroll
    for (int l = 0; l < (TN * TK) / 256; ++l) {
      const int e = tid + l * 256;
      const int k = e % TK, n = e / TK;
      const long long gn = n0 + n;
      const int gk = k0 + k;
      Es[k][n] = (gn < n_end && gk < d) ? En[(size_t)gn * d + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 av = *reinterpret_cast<const float4 *>(&Qs[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4 *>(&Es[k][tx * 4]);
      const float a_[4] = {av.x, av.y, av.z, av.w};
      const float b_[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a_[i], b_[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gq = q0 + ty * 4 + i;
    if (gq >= Q) continue;
    // rows are swept in increasing id order, so `score > tau_score` is the exact key test (see gemm_tc.cu)
    const unsigned long long t = tau[gq];
    const float tau_score = t == 0ull ? -INFINITY : key_score(t);
    int n_pass = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) n_pass += (n0 + tx * 4 + j < n_end && acc[i][j] > tau_score) ? 1 : 0;
    if (n_pass == 0) continue;
    int pos = atomicAdd(count + gq, n_pass);               // one reservation per (thread, query row)
    if (pos + n_pass > cap) *overflow = 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long gn = n0 + tx * 4 + j;
      if (gn < n_end && acc[i][j] > tau_score) {
        if (pos < cap) cand[(size_t)gq * cap + pos] = make_key(acc[i][j], (unsigned int)(gn + row_offset));
        ++pos;
      }
    }
  }
}

// ---- block-wide bitonic sort (descending) of n_pow2 keys in shared memory -------------------------
__device__ void bitonic_sort_desc(unsigned long long *keys, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
}

__global__ void reset_topk_state_kernel(unsigned long long *tau, int *count, int Q) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < Q) { tau[q] = 0ull; count[q] = 0; }
}

// fp32 inner product of a query held in shared memory with one entity row, by a whole warp.  ONE summation order for
// every caller (finalize_kernel, rescore_kernel), so a (query, row) pair scores bit-identically whichever path or
// shard layout produced the candidate: lane l accumulates elements 4l..4l+3, 4(l+32).. with fmaf, then a butterfly.
__device__ __forceinline__ float warp_dot(const float *__restrict__ qs, const float *__restrict__ e, int d, int lane) {
  float s = 0.f;
  if ((d & 3) == 0 && (reinterpret_cast<uintptr_t>(e) & 15) == 0) {
    const float4 *e4 = reinterpret_cast<const float4 *>(e);
    const float4 *q4 = reinterpret_cast<const float4 *>(qs);
    for (int c = lane; c < (d >> 2); c += 32) {
      const float4 ev = __ldg(e4 + c);
      const float4 qv = q4[c];
      s = fmaf(ev.x, qv.x, s); s = fmaf(ev.y, qv.y, s); s = fmaf(ev.z, qv.z, s); s = fmaf(ev.w, qv.w, s);
    }
  } else {
    for (int c0 = lane * 4; c0 < d; c0 += 128) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + j < d) s = fmaf(__ldg(e + c0 + j), qs[c0 + j], s);
    }
  }
  return warp_sum(s);
}

// Exact float32 re-scoring of the surviving candidates (tensor-core mode): the bf16 scores picked the
// candidates, the returned scores and the final order come from fp32 dot products of the fp32 vectors.
__global__ void __launch_bounds__(256) rescore_kernel(const float *__restrict__ Qm, const float *__restrict__ En,
                                                      int d, long long row_begin, unsigned long long *__restrict__ cand,
                                                      const int *__restrict__ count, int cap) {
  extern __shared__ __align__(16) float qs[];
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int n = min(count[q], cap);
  for (int c = threadIdx.x; c < d; c += blockDim.x) qs[c] = Qm[(size_t)q * d + c];
  __syncthreads();
  unsigned long long *list = cand + (size_t)q * cap;
  for (int j = warp; j < n; j += nwarps) {
    const unsigned int r0 = key_row(list[j]);
    const float s0 = warp_dot(qs, En + (size_t)((long long)r0 - row_begin) * d, d, lane);
    if (lane == 0) list[j] = make_key(s0, r0);
  }
}

// Queries of a tensor-core sweep, one warp per query: the 3-term bf16 split row [hi | hi | mid] (the A operand of
// gemm_tc.cuh), the reset of the query's list state, and the rigorous error margin of the COARSE scores (hi.hi term
// alone):  q.e - qh.eh = (q - qh).e + qh.(e - eh)  =>  |coarse - exact| <= |q - qh| max|e| + |qh| max|e - eh|
// (Cauchy-Schwarz, exact in real arithmetic; bf16 round-to-nearest has unit roundoff 2^-8, and measuring the two
// residual norms instead of assuming the worst case 2^-8 |q| roughly halves the band), plus d 2^-22 |qh| max|eh| for
// the fp32 accumulation of the exact bf16 products inside the tensor core (truncating adds: one ulp each).  Two rows
// whose exact scores order one way can order the other way in the coarse scores only within 2 eps_q: margin = 2 eps_q.
__global__ void __launch_bounds__(256) prep_queries_kernel(const float *__restrict__ Qm, int Q, int products, int d, int Kp,
                                                           __nv_bfloat16 *__restrict__ split, float ent_norm_max,
                                                           float ent_err_max, float *__restrict__ margin,
                                                           unsigned long long *__restrict__ tau, int *__restrict__ count,
                                                           int *__restrict__ flag) {
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) *flag = 0;        // the seeded sweep's verdict word (no memset node of its own)
  if (q >= Q) return;
  __nv_bfloat16 *row = split + (size_t)q * 3 * Kp;
  float e2 = 0.f, h2 = 0.f;
  for (int c = lane; c < Kp; c += 32) {
    const float x = c < d ? Qm[(size_t)q * d + c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const float fh = __bfloat162float(hi);
    const float r = x - fh;                                  // exact (Sterbenz-like: fh is x rounded to 8 bits)
    row[c] = hi;
    row[Kp + c] = hi;
    row[2 * Kp + c] = __float2bfloat16_rn(r);
    e2 = fmaf(r, r, e2);
    h2 = fmaf(fh, fh, h2);
  }
  e2 = warp_sum(e2);
  h2 = warp_sum(h2);
  if (lane == 0) {
    const float qh = sqrtf(h2);
    // ent_err_max is the residual of what the coarse GEMM keeps of an entity row (hi alone, or hi + mid); `d` counts the
    // accumulated products (twice the dimension with two blocks)
    const float eps = sqrtf(e2) * ent_norm_max + qh * ent_err_max + (float)products * 2.3841858e-7f * qh * ent_norm_max * 1.004f;
    margin[q] = 2.0f * eps * 1.02f + 1e-30f;
    tau[q] = 0ull;
    count[q] = 0;
  }
}

// largest row norm and largest residual norms after one / two bf16 terms, as ordered uint bits (norms are >= 0): out[0..2]
__global__ void __launch_bounds__(256) row_norm_max_kernel(const float *__restrict__ E, long long rows, int d,
                                                           unsigned int *__restrict__ out) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float ss = 0.f, rr = 0.f, r2 = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = E[(size_t)r * d + c];
    const float res = v - __bfloat162float(__float2bfloat16_rn(v));
    const float res2 = res - __bfloat162float(__float2bfloat16_rn(res));
    ss = fmaf(v, v, ss);
    rr = fmaf(res, res, rr);
    r2 = fmaf(res2, res2, r2);
  }
  ss = warp_sum(ss);
  rr = warp_sum(rr);
  r2 = warp_sum(r2);
  if (lane == 0) {
    atomicMax(out, __float_as_uint(sqrtf(ss)));
    atomicMax(out + 1, __float_as_uint(sqrtf(rr)));
    atomicMax(out + 2, __float_as_uint(sqrtf(r2)));
  }
}

// ---- seeded sweep: threshold seed from a strided row sample --------------------------------------------------
// The sample GEMM (TC_EPI_GROUPMAX) left, per query, the maxima of G groups of g sampled rows.  The j-th largest
// group maximum is a LOWER bound of the j-th largest sampled score, and with j/G ~ 1 - exp(-g T/rows) about T rows of
// the shard score above it (topk_plan picks g, G, j for T ~ 5k).  tau = that value - margin; the one-launch sweep
// then appends every row scoring above tau.  Nothing here has to be right for the result to be exact: finalize_kernel
// verifies that the list holds >= k rows and that (k-th best - margin) clears tau, else the flag sends the call to
// the multi-chunk sweep.  One warp per query; bitwise bisection on the order-preserving bits.
__global__ void __launch_bounds__(256) seed_tau_kernel(const float *__restrict__ gmax, int ld, int G, int j,
                                                       const float *__restrict__ margin,
                                                       unsigned long long *__restrict__ tau, int Q) {
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= Q) return;
  unsigned int v[kSeedGroups / 32];
#pragma unroll
  for (int i = 0; i < kSeedGroups / 32; ++i) {
    const int g = lane + 32 * i;
    v[i] = g < G ? orderable(gmax[(size_t)q * ld + g]) : 0u;
  }
  unsigned int t = 0u;
#pragma unroll 1
  // 20 of the 32 bits: the result (low bits zero) is a lower bound of the exact j-th largest value, 2^-11 relative
  // below it at most -- any lower bound seeds a valid threshold, and this one lets through well under 1 % more rows
  for (int bit = 31; bit >= 12; --bit) {
    const unsigned int cand = t | (1u << bit);
    unsigned int c = 0u;
#pragma unroll
    for (int i = 0; i < kSeedGroups / 32; ++i) c += v[i] >= cand ? 1u : 0u;
    // shuffle butterfly, not __reduce_add_sync: REDUX issues at a fraction of the shuffle rate, and 64 resident warps
    // doing 32 of them each made this kernel 28 us for 10 k queries
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (c >= (unsigned int)j) t = cand;
  }
  if (lane == 0) tau[q] = make_key(unorderable(t) - margin[q], 0xffffffffu);
}

// ---- seeded sweep: selection + exact re-scoring + final order in ONE kernel, one warp per query ---------------
// The list holds every row whose coarse score beat tau (a few hundred to a thousand keys, straight from the GEMM
// epilogue, in L2).  (1) A 256-bin histogram between the list's smallest and largest score locates the bin of the
// k-th best; Lf = the smallest score in that bin or above is a lower bound of the k-th best coarse score.  (2) Every
// row with coarse score >= Lf - margin is kept: by the margin's construction this includes every row of the exact
// top k (prep_queries_kernel).  (3) The survivors (k + the margin band) are re-scored with fp32 dot products, sorted
// in shared memory, and the best k are written.  Sets *flag (=> multi-chunk sweep) when the list is short of k rows,
// when Lf - margin does not clear the seeded tau (rows the guarantee needs may have been filtered), or when more
// survivors than the shared-memory slots remain (masses of near-ties).
struct FinalizeArgs {
  const float *Qm;
  const float *En;
  int Q, d;
  long long row_begin, rows;
  const unsigned long long *cand;
  const int *count;
  int cap;
  const unsigned long long *tau;
  const float *margin;
  int k, cmax;
  int32_t *out_idx;
  float *out_score;
  int *flag;
  int per_warp_bytes;
};

constexpr int kFinRegs = 32;      // keys per lane held in registers: lists of up to 1024 keys are read from L2 once

__global__ void __launch_bounds__(256, 3) finalize_kernel(const FinalizeArgs a) {
  extern __shared__ __align__(16) unsigned char fin_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= a.Q) return;
  unsigned char *mine = fin_smem + (size_t)warp * a.per_warp_bytes;
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(mine);
  int *hist = reinterpret_cast<int *>(mine + (size_t)a.cmax * 8);
  float *qs = reinterpret_cast<float *>(mine + (size_t)a.cmax * 8 + 1024);
  const unsigned long long *list = a.cand + (size_t)q * a.cap;
  const int n = min(a.count[q], a.cap);
  const int need = (int)min((long long)a.k, a.rows);
  const unsigned long long tk = a.tau[q];
  const float tau_s = tk != 0ull ? key_score(tk) : -INFINITY;
  int32_t *oi = a.out_idx + (size_t)q * a.k;
  float *os = a.out_score + (size_t)q * a.k;
  if (n < need || a.count[q] > a.cap) {
    if (lane == 0) *a.flag = 1;
    return;
  }
  if (need == 0) {
    for (int i = lane; i < a.k; i += 32) { oi[i] = -1; os[i] = -INFINITY; }
    return;
  }
  // the query vector, for step (3); issued first so that it travels with the list
  for (int cc = lane; cc < a.d; cc += 32) qs[cc] = a.Qm[(size_t)q * a.d + cc];
  // The list: in registers when it fits (one coalesced sweep, every load in flight at once), else re-read per pass.
  const bool inreg = n <= 32 * kFinRegs;
  unsigned long long kreg[kFinRegs];
  if (inreg) {
#pragma unroll
    for (int i = 0; i < kFinRegs; ++i) kreg[i] = (i * 32 + lane < n) ? list[i * 32 + lane] : 0ull;
  }
#define SERT_FOR_EACH_KEY(BODY)                                                   \
  if (inreg) {                                                                    \
    _Pragma("unroll") for (int i_ = 0; i_ < kFinRegs; ++i_) {                     \
      if (i_ * 32 < n && i_ * 32 + lane < n) { const unsigned long long key = kreg[i_]; BODY }   \
    }                                                                             \
  } else {                                                                        \
    for (int i_ = lane; i_ < n; i_ += 32) { const unsigned long long key = list[i_]; BODY }      \
  }
  // (1) score range of the list, histogram, bin of the need-th best
  float hi = -INFINITY, lo = INFINITY;
  SERT_FOR_EACH_KEY({ const float s = key_score(key); hi = fmaxf(hi, s); lo = fminf(lo, s); })
  hi = warp_max(hi);
  lo = -warp_max(-lo);
  const float scale = hi > lo ? 255.9f / (hi - lo) : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) hist[lane * 8 + i] = 0;
  __syncwarp();
  SERT_FOR_EACH_KEY({ atomicAdd(&hist[min(255, (int)((key_score(key) - lo) * scale))], 1); })
  __syncwarp();
  int local[8], lsum = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { local[i] = hist[lane * 8 + i]; lsum += local[i]; }
  int incl = lsum;                                           // suffix sums over lanes: lanes above hold higher bins
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += v;
  }
  const int above = incl - lsum;                             // keys in bins of higher lanes
  int bstar = -1;
  if (above < need && need <= above + lsum) {
    int acc = above;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      if (bstar < 0 && acc + local[i] >= need) bstar = lane * 8 + i;
      acc += local[i];
    }
  }
  bstar = __reduce_max_sync(0xffffffffu, bstar);
  // (2) Lf = smallest score at or above that bin; survivors = everything within the margin below it
  float lf = INFINITY;
  SERT_FOR_EACH_KEY({
    const float s = key_score(key);
    if (min(255, (int)((s - lo) * scale)) >= bstar) lf = fminf(lf, s);
  })
  lf = -warp_max(-lf);
  const float thr = lf - a.margin[q];
  if (tk != 0ull && !(thr > tau_s)) {
    if (lane == 0) *a.flag = 1;
    return;
  }
  int c = 0;
  if (inreg) {
#pragma unroll
    for (int i = 0; i < kFinRegs; ++i) {
      if (i * 32 < n) {                                      // warp-uniform
        const bool keep = (i * 32 + lane < n) && key_score(kreg[i]) >= thr;
        const unsigned int m = __ballot_sync(0xffffffffu, keep);
        const int pos = c + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < a.cmax) keys[pos] = kreg[i];
        c += __popc(m);
      }
    }
  } else {
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      unsigned long long key = 0ull;
      bool keep = false;
      if (i < n) {
        key = list[i];
        keep = key_score(key) >= thr;
      }
      const unsigned int m = __ballot_sync(0xffffffffu, keep);
      const int pos = c + __popc(m & ((1u << lane) - 1u));
      if (keep && pos < a.cmax) keys[pos] = key;
      c += __popc(m);
    }
  }
#undef SERT_FOR_EACH_KEY
  if (c > a.cmax) {
    if (lane == 0) *a.flag = 1;
    return;
  }
  __syncwarp();
  // (3) exact fp32 scores of the survivors, eight rows in flight per warp (a row is one or two 16-byte requests per
  // lane: latency-bound on its own)
  constexpr int RF = 8;
  for (int j0 = 0; j0 < c; j0 += RF) {
    unsigned int r[RF];
    float sc[RF];
#pragma unroll
    for (int i = 0; i < RF; ++i) r[i] = key_row(keys[min(j0 + i, c - 1)]);
    if ((a.d & 3) == 0) {
      const float4 *q4 = reinterpret_cast<const float4 *>(qs);
      const float4 *e4[RF];
#pragma unroll
      for (int i = 0; i < RF; ++i)
        e4[i] = reinterpret_cast<const float4 *>(a.En + (size_t)((long long)r[i] - a.row_begin) * a.d);
#pragma unroll
      for (int i = 0; i < RF; ++i) sc[i] = 0.f;
      for (int cc = lane; cc < (a.d >> 2); cc += 32) {
        const float4 qv = q4[cc];
        float4 ev[RF];
#pragma unroll
        for (int i = 0; i < RF; ++i) ev[i] = __ldg(e4[i] + cc);
#pragma unroll
        for (int i = 0; i < RF; ++i) {
          sc[i] = fmaf(ev[i].x, qv.x, sc[i]); sc[i] = fmaf(ev[i].y, qv.y, sc[i]);
          sc[i] = fmaf(ev[i].z, qv.z, sc[i]); sc[i] = fmaf(ev[i].w, qv.w, sc[i]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < RF; ++i) {
        const float *e = a.En + (size_t)((long long)r[i] - a.row_begin) * a.d;
        float s = 0.f;
        for (int c0 = lane * 4; c0 < a.d; c0 += 128) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c0 + j < a.d) s = fmaf(__ldg(e + c0 + j), qs[c0 + j], s);
        }
        sc[i] = s;
      }
    }
    // Eight butterfly reductions at once, transposed: at the xor-16 / 8 / 4 steps a lane hands over the partial sums
    // of the rows its partner keeps (4, 2, 1 exchanges instead of 8 each), the last two steps are plain.  Every row's
    // sum is built by the SAME tree as warp_sum (pairs 16 apart, then 8, 4, 2, 1; fp addition commutes), so the
    // result is bit-identical to warp_dot's.  Row i ends up in the lanes with (lane >> 2) & 7 == bitrev3(i).
    {
      const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4;
      float t4[4], t2[2], t1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {          // keep rows i (lower half) or i + 4 (upper half)
        const float give = up16 ? sc[i] : sc[i + 4];
        const float keep = up16 ? sc[i + 4] : sc[i];
        t4[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float give = up8 ? t4[i] : t4[i + 2];
        const float keep = up8 ? t4[i + 2] : t4[i];
        t2[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
      }
      {
        const float give = up4 ? t2[0] : t2[1];
        const float keep = up4 ? t2[1] : t2[0];
        t1 = keep + __shfl_xor_sync(0xffffffffu, give, 4);
      }
      t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
      t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
      // this lane's row: bit 2 of i from lane bit 4, bit 1 from lane bit 3, bit 0 from lane bit 2
      const int mine_i = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
      unsigned int mine_r = r[0];
#pragma unroll
      for (int i = 1; i < RF; ++i)
        if (mine_i == i) mine_r = r[i];
      __syncwarp();
      if ((lane & 3) == 0 && j0 + mine_i < c) keys[j0 + mine_i] = make_key(t1, mine_r);
    }
  }
  // (4) order the survivors (bitonic, descending, in the warp's shared-memory slots) and write the best `need`
  int n_pow2 = 32;
  while (n_pow2 < c) n_pow2 <<= 1;
  for (int i = c + lane; i < n_pow2; i += 32) keys[i] = 0ull;
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncwarp();
      for (int t = lane; t < (n_pow2 >> 1); t += 32) {
        const int l = 2 * t - (t & (stride - 1));
        const int h = l + stride;
        const bool desc = ((l & size) == 0);
        const unsigned long long x = keys[l], y = keys[h];
        if ((x < y) == desc) { keys[l] = y; keys[h] = x; }
      }
    }
  }
  __syncwarp();
  for (int i = lane; i < a.k; i += 32) {
    if (i < need) {
      oi[i] = (int32_t)key_row(keys[i]);
      os[i] = key_score(keys[i]);
    } else {
      oi[i] = -1;
      os[i] = -INFINITY;
    }
  }
}

// ---- prune by radix selection ---------------------------------------------------------------------------
// Re-selects the best k candidates of a query and raises tau.  mode 0: only lists longer than cap/4; mode 2: every
// list; mode 1 ("final"): every list, and the sorted (row id, score) outputs are written.  The k-th largest key is FOUND (11-bit MSB radix passes over the 64-bit keys
// read from L2, then a direct ranking once <= 256 keys share the prefix) instead of sorting the whole list
// (a full bitonic sort measured ~6 us per list of 1-2 k candidates and needed 64 KB of shared memory for the
// longest lists; this is ~1 us with 8 KB + k keys for any list length).  Keys are unique (score bits | ~row id), so exactly k keys are >= the selected threshold.  Only
// the final call sorts, and only the k survivors.
constexpr int kSelBins = 2048;

__global__ void __launch_bounds__(256) prune_select_kernel(unsigned long long *__restrict__ cand, int *__restrict__ count,
                                                           unsigned long long *__restrict__ tau, int cap, int k,
                                                           int mode, int32_t *__restrict__ out_idx,
                                                           float *__restrict__ out_score, int k_pow2,
                                                           const float *__restrict__ margin, int *__restrict__ overflow) {
  extern __shared__ unsigned long long sel_out[];            // k_pow2 keys
  __shared__ unsigned int hist[kSelBins];
  __shared__ unsigned int warp_tot[8];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_krem, s_bucket, s_fill;
  __shared__ unsigned long long small[256];
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(count[q], cap);
  if (mode == 0 && n <= cap / 4) return;
  unsigned long long *mine = cand + (size_t)q * cap;
  int keep = min(n, k);
  unsigned long long T = 0ull;                               // threshold key: keep keys >= T

  if (n > k) {
    if (tid == 0) { s_prefix = 0ull; s_krem = k; }
    int shift = 64;
    unsigned long long mask = 0ull;                          // bits already decided
    __syncthreads();
    while (n > 256) {                                        // short lists go straight to the direct ranking
      const int bits = shift >= 11 ? 11 : shift;
      shift -= bits;
      for (int b = tid; b < kSelBins; b += blockDim.x) hist[b] = 0u;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int i = tid; i < n; i += blockDim.x) {
        const unsigned long long key = mine[i];
        if ((key & mask) == prefix) atomicAdd(&hist[(unsigned int)(key >> shift) & ((1u << bits) - 1u)], 1u);
      }
      __syncthreads();
      // suffix scan over the bins: each thread owns 8 consecutive bins (descending digit order)
      const int nb = 1 << bits, per = kSelBins / 256;
      unsigned int local[8], tsum = 0u;
#pragma unroll
      for (int j = 0; j < per; ++j) {
        const int bin = nb - 1 - (tid * per + j);           // thread 0 holds the highest digits
        local[j] = bin >= 0 ? hist[bin] : 0u;
        tsum += local[j];
      }
      unsigned int incl = tsum;                              // inclusive scan over threads (descending digits)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) warp_tot[warp] = incl;
      __syncthreads();
      unsigned int base = 0u;
      for (int w = 0; w < warp; ++w) base += warp_tot[w];
      const unsigned int before = base + incl - tsum;        // keys with a larger digit than this thread's bins
      const int krem = s_krem;
      __syncthreads();
      if ((unsigned int)krem > before && (unsigned int)krem <= before + tsum) {
        unsigned int acc = before;
#pragma unroll
        for (int j = 0; j < per; ++j) {
          if ((unsigned int)krem <= acc + local[j]) {
            const int bin = nb - 1 - (tid * per + j);
            s_prefix = prefix | ((unsigned long long)bin << shift);
            s_krem = krem - (int)acc;
            s_bucket = (int)local[j];
            break;
          }
          acc += local[j];
        }
      }
      __syncthreads();
      mask |= (((1ull << bits) - 1ull) << shift);
      if (shift == 0 || s_bucket <= 256) break;
    }
    // direct ranking inside the bucket (<= 256 keys share the decided prefix, or every bit is decided)
    const unsigned long long prefix = s_prefix;
    if (tid == 0) s_fill = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned long long key = mine[i];
      if ((key & mask) == prefix) small[atomicAdd(&s_fill, 1)] = key;
    }
    __syncthreads();
    const int nsmall = s_fill, krem = s_krem;
    if (tid < nsmall) {
      const unsigned long long mykey = small[tid];
      int larger = 0;
      for (int j = 0; j < nsmall; ++j) larger += small[j] > mykey ? 1 : 0;
      if (larger == krem - 1) s_prefix = mykey;              // the k-th largest key overall
    }
    __syncthreads();
    T = s_prefix;
  }

  if (margin != nullptr) {
    // ---- margin mode (coarse scores): keep EVERY candidate whose score is within margin[q] of the k-th best, so that
    // no row whose exact score belongs to the top k is dropped (topk_sweep).  k_pow2 = cap/2 slots of shared memory.
    if (n < k) return;                                       // nothing to drop, no threshold yet
    if (tid == 0) s_fill = 0;
    __syncthreads();
    if (n == k) {                                            // the k-th best is the minimum
      unsigned long long mn = ~0ull;
      for (int i = tid; i < n; i += blockDim.x) mn = min(mn, mine[i]);
      for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      if (lane == 0) small[warp] = mn;
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < 8; ++w) mn = min(mn, small[w]);
        tau[q] = make_key(key_score(mn) - margin[q], 0xffffffffu);
      }
      return;
    }
    const unsigned long long Tm = make_key(key_score(T) - margin[q], 0xffffffffu);   // smallest key of that score
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned long long key = mine[i];
      if (key >= Tm) {
        const int slot = atomicAdd(&s_fill, 1);
        if (slot < k_pow2) sel_out[slot] = key;
      }
    }
    __syncthreads();
    const int c = s_fill;
    if (c > k_pow2) {                                        // too many near-ties for the coarse scores: exact sweep instead
      if (tid == 0) *overflow = 1;
      return;
    }
    for (int i = tid; i < c; i += blockDim.x) mine[i] = sel_out[i];
    if (tid == 0) { count[q] = c; tau[q] = Tm; }
    return;
  }

  // ---- compaction: the survivors, in arbitrary order, go to shared memory and then to the list head ----
  if (tid == 0) s_fill = 0;
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    const unsigned long long key = mine[i];
    if (key >= T) sel_out[atomicAdd(&s_fill, 1)] = key;
  }
  __syncthreads();
  unsigned long long kth = ~0ull;
  if (mode == 1) {
    for (int i = keep + tid; i < k_pow2; i += blockDim.x) sel_out[i] = 0ull;
    bitonic_sort_desc(sel_out, k_pow2);
    if (keep > 0) kth = sel_out[keep - 1];
  } else {
    kth = T;
  }
  for (int i = tid; i < keep; i += blockDim.x) mine[i] = sel_out[i];
  if (tid == 0) {
    count[q] = keep;
    if (keep == k) {
      if (n > k) tau[q] = T;
      else if (mode == 1) tau[q] = kth;
      else {                                                  // exactly k candidates, unsorted: tau = their minimum
        unsigned long long mn = ~0ull;
        for (int i = 0; i < keep; ++i) mn = sel_out[i] < mn ? sel_out[i] : mn;
        tau[q] = mn;
      }
    }
  }
  if (mode == 1) {
    for (int i = tid; i < k; i += blockDim.x) {
      if (i < keep) {
        out_idx[(size_t)q * k + i] = (int32_t)key_row(sel_out[i]);
        out_score[(size_t)q * k + i] = key_score(sel_out[i]);
      } else {                                               // fewer than k rows in the shard
        out_idx[(size_t)q * k + i] = -1;
        out_score[(size_t)q * k + i] = -INFINITY;
      }
    }
  }
}

static int launch_prune(const TopkState &s, int Q, int k, int mode, int32_t *out_idx, float *out_score,
                        cudaStream_t st, const float *margin = nullptr) {
  int k_pow2 = 2;
  while (k_pow2 < k) k_pow2 <<= 1;
  // margin mode keeps k + (rows inside the margin band) survivors: 1024 slots (8 KB, 8 CTAs per SM) hold them on
  // every workload where the coarse sweep pays; more near-ties than that raise the overflow flag (bf16x3 sweep)
  if (margin != nullptr) k_pow2 = std::min(s.cap / 2, std::max(1024, 2 * k_pow2));
  prune_select_kernel<<<Q, 256, (size_t)k_pow2 * sizeof(unsigned long long), st>>>(
      s.cand, s.count, s.tau, s.cap, k, mode, out_idx, out_score, k_pow2, margin, s.overflow);
  SERT_LAUNCH_CHECK();
  return 0;
}

// One pass over the shard.  `optimistic`: geometrically growing chunks (1024, 4096, ... rows) with a forced
// prune after each, so tau tightens early and later chunks append only ~k*chunk/seen rows per query; a list
// that would overflow sets *overflow and the caller falls back to the conservative pass (chunk = cap/2 rows,
// which cannot overflow even if every score of a chunk survives).
// rows per launch once tau has warmed up: two 256-row n-tiles per SM (the B-stationary schedule of gemm_tc.cu
// hands whole n-tiles to CTAs)
constexpr long long kBigChunk = 2ll * kNumSMs * 256;

static int topk_pass(const TopkState &s, const float *queries_dev, int Q, int k_sel, bool optimistic,
                     cudaStream_t st, bool coarse = false) {
  const bool tensor = s.mode == SCORE_TENSOR;
  reset_topk_state_kernel<<<cdiv(Q, 256), 256, 0, st>>>(s.tau, s.count, Q);
  SERT_LAUNCH_CHECK();
  SERT_CUDA(cudaMemsetAsync(s.overflow, 0, sizeof(int), st));
  long long chunk = optimistic ? std::min<long long>(1024, s.cap / 2) : s.cap / 2;
  for (long long n0 = 0; n0 < s.rows;) {
    const long long n1 = std::min<long long>(s.rows, n0 + chunk);
    if (tensor) {
      TcEpilogue ep;
      ep.mode = TC_EPI_TOPK;
      ep.tau = s.tau; ep.count = s.count; ep.cand = s.cand; ep.cap = s.cap; ep.row_offset = s.row_begin;
      ep.overflow = s.overflow;
      // coarse: the first block of both operands is the hi term ([hi|hi|mid] x [hi|mid|hi]); same buffers, depth kt/3
      const int depth = coarse ? s.coarse_blocks * (s.kt / s.terms) : s.kt;
      if (launch_gemm_tc_ld(s.q_split, s.kt, Q, s.ent_split, s.kt, s.rows, n0, n1, depth, ep, st)) return -1;
    } else {
      dim3 grid(cdiv(n1 - n0, TN), cdiv(Q, TQ));
      score_filter_kernel<<<grid, 256, 0, st>>>(queries_dev, s.entities, Q, n0, n1, s.d, s.row_begin, s.tau, s.count,
                                                s.cand, s.cap, s.overflow);
      SERT_LAUNCH_CHECK();
    }
    // forced prune (mode 2) after every optimistic chunk: tau only moves in the prune, and a stale tau lets ~k more
    // rows per query through the next chunk's epilogue (measured: 13.2 ms vs 10.6 ms at BASELINE configs[3]).  The
    // conservative pass prunes the lists more than a quarter full, and everything at the end.
    const bool force = optimistic || n1 == s.rows;
    if (launch_prune(s, Q, k_sel, force ? 2 : 0, nullptr, nullptr, st, coarse ? s.margin : nullptr)) return -1;
    n0 = n1;
    // chunks grow x4 up to 64k rows, then double (rounded to two n-tiles per SM): the expected number of new
    // candidates per query and chunk stays ~k ln 2, and the number of launches ~log2(rows)
    if (optimistic) {
      chunk = chunk < (1 << 16) ? chunk * 4 : std::max<long long>(kBigChunk, n0 / kBigChunk * kBigChunk);
    }
  }
  return 0;
}

// ---- optional phase timing (SERT_SCORE_TRACE=1): CUDA events between the phases of a call, printed to stderr ------
struct PhaseTrace {
  static constexpr int kMax = 12;
  cudaEvent_t ev[kMax];
  const char *name[kMax];
  int n = 0;
  bool on = false;
  cudaStream_t st = nullptr;
  explicit PhaseTrace(cudaStream_t s) : st(s) {
    static const bool enabled = getenv("SERT_SCORE_TRACE") != nullptr;
    on = enabled;
  }
  void mark(const char *what) {
    if (!on || n >= kMax) return;
    cudaEventCreate(&ev[n]);
    cudaEventRecord(ev[n], st);
    name[n++] = what;
  }
  void report(const char *title) {
    if (!on || n < 2) return;
    cudaEventSynchronize(ev[n - 1]);
    fprintf(stderr, "[sert trace] %s:", title);
    for (int i = 1; i < n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      fprintf(stderr, " %s %.1f us |", name[i], ms * 1e3f);
    }
    float tot = 0.f;
    cudaEventElapsedTime(&tot, ev[0], ev[n - 1]);
    fprintf(stderr, " total %.1f us\n", tot * 1e3f);
    for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
    n = 0;
  }
};

// ---- seeded sweep: plan ------------------------------------------------------------------------------------------
// Picks the row sample (G groups of g rows = G*g/256 strided n-tiles) and the rank j of the group maximum that seeds
// tau so that about T rows per query survive the one-launch sweep.  With x = g T / rows a group holds a row above the
// T-th best score with probability 1 - exp(-x); j = that fraction of G.  The survivor count concentrates like
// Gamma(j) / j: P(fewer than k survive) ~ P(Gamma(j) < j k / T) = P(Poisson(j k / T) >= j), so a smaller T needs a
// larger j, i.e. a larger sample.  The epilogue and the finalize kernel cost in proportion to T (measured: the sweep
// of 10 k queries over 50 k rows takes ~0.7 us per unit of T), the sample in proportion to G*g rows: the plan takes the
// smallest T = r k, r in {3, 4, 5, 6, 8, 10, 16}, whose sample stays below a sixth of the shard and 512 groups, at a
// failure probability below 1e-7 per query.  Shards of at most cap/2 rows need no seed (every row fits the list).
struct SweepPlan {
  bool seeded = false;
  int g = 0, G = 0, j = 0;
  long long stride = 1;       // n-tiles between sample tiles
  double T = 0;               // expected survivors per query
};

// smallest j with P(Poisson(j / r) >= j) <= eps
static int seed_rank_for_ratio(double r, double eps) {
  for (int j = 4; j <= 400; ++j) {
    const double lam = j / r;
    double term = exp(-lam), cdf = term;             // P(X <= j-1)
    for (int i = 1; i < j; ++i) { term *= lam / i; cdf += term; }
    if (1.0 - cdf <= eps) return j;
  }
  return 400;
}

static SweepPlan topk_plan(long long rows, int k, int cap) {
  SweepPlan best;
  if (rows <= cap / 2) return best;
  const double ratios[] = {3, 4, 5, 6, 8, 10, 16};
  for (double r : ratios) {
    const double T = std::max(r * k, 96.0);
    if (T > cap / 4) break;
    const int j = std::max(seed_rank_for_ratio(T / k, 1e-7), 8);
    for (int g : {8, 16, 32, 64}) {
      const int unit = 256 / g;                               // groups per n-tile
      const double x = T * g / (double)rows;
      if (x > 1.2) continue;
      const double frac = 1.0 - exp(-x);
      long long G = (long long)ceil(j / frac);
      G = (G + unit - 1) / unit * unit;
      if (G > kSeedGroups) continue;
      const long long S = G * g;
      if (S * 6 > rows) continue;
      const int jj = std::max(j, (int)floor(frac * (double)G));   // rounding G up may only raise the rank, never T
      if (!best.seeded || S < (long long)best.G * best.g) {
        best.seeded = true;
        best.g = g; best.G = (int)G; best.j = jj;
        best.stride = ((rows + 255) / 256) / (S / 256);
        best.T = (double)rows * -log(1.0 - (double)jj / (double)G) / g;
      }
    }
    if (best.seeded) return best;
  }
  return best;
}

static int finalize_per_warp_bytes(int cmax, int d) { return (int)align_up((size_t)cmax * 8 + 1024 + (size_t)d * 4, 16); }

// `deferred`: non-null = the caller checks the seeded attempt itself.  If the seeded one-launch sweep applies, it is
// only ENQUEUED (no host synchronisation): *deferred = 1 and the verdict is left in s.overflow on the device (non-zero
// = the result is not valid; call again with deferred == nullptr).  Otherwise *deferred = 0 and the call completes as
// usual.  `allow_seeded` = false skips the seeded attempt (the retry after a failed one).
int topk_sweep(const TopkState &s, const float *queries_dev, int Q, int k, int32_t *out_idx, float *out_score,
               cudaStream_t st, int *deferred, bool allow_seeded) {
  if (deferred) *deferred = 0;
  SERT_REQUIRE(k >= 1 && k <= s.cap / 2, "k exceeds the scorer's max_k");
  SERT_REQUIRE(Q >= 0 && Q <= s.max_queries, "more queries than the scorer's max_queries");
  if (Q == 0) return 0;
  const bool tensor = s.mode == SCORE_TENSOR;
  // tensor-core mode keeps a margin of candidates beyond k so that bf16x3 rounding at the k-th place
  // cannot drop a true top-k row before the exact re-scoring
  const int k_sel = tensor ? std::min(s.cap / 2, k + 16) : k;
  PhaseTrace trace(st);
  trace.mark("start");
  if (tensor) {
    // split rows [hi|hi|mid] of the queries, their coarse-score margins, list state reset
    prep_queries_kernel<<<cdiv(Q, 8), 256, 0, st>>>(queries_dev, Q, s.d * s.coarse_blocks, s.d, s.kt / s.terms, s.q_split,
                                                    s.ent_norm_max, s.coarse_blocks == 2 ? s.ent_err2_max : s.ent_err_max,
                                                    s.margin, s.tau, s.count, s.overflow);
    SERT_LAUNCH_CHECK();
  }
  int overflow = 1;
  if (tensor && s.coarse && s.terms == 3 && s.margin != nullptr) {
    // Coarse-then-exact: ONE bf16 GEMM (the hi.hi term, a third of the tensor work) scores every row with an error of
    // at most eps_q (prep_queries_kernel).  Every row of the exact top k then scores within 2 eps_q of the coarse k-th
    // best, so everything above (k-th - 2 eps_q) is kept and re-scored in fp32.
    int cmax = 512;
    while (cmax < 2 * k + 256) cmax <<= 1;
    const int per_warp = finalize_per_warp_bytes(cmax, s.d);
    const int warps = std::min(8, (200 * 1024) / per_warp);
    const SweepPlan plan = topk_plan(s.rows, k, s.cap);
    if (allow_seeded && s.seeded && warps >= 1 && s.rows > 0 && (plan.seeded || s.rows <= s.cap / 2)) {
      // Seeded one-launch sweep: sample GEMM -> tau -> ONE GEMM over the shard -> finalize (select, re-score, sort).
      trace.mark("prep");                   // prep_queries_kernel also cleared the verdict word
      TcEpilogue ep;
      const int depth = s.coarse_blocks * (s.kt / s.terms);     // leading block(s) of both split operands
      if (plan.seeded) {
        ep.mode = TC_EPI_GROUPMAX;
        ep.gmax = s.gmax; ep.gmax_ld = kSeedGroups; ep.group = plan.g; ep.tile_stride = (int)plan.stride;
        const long long sample_rows = (long long)plan.G * plan.g;
        if (launch_gemm_tc_ld(s.q_split, s.kt, Q, s.ent_split, s.kt, s.rows, 0, sample_rows, depth, ep, st)) return -1;
        trace.mark("sample");
        seed_tau_kernel<<<cdiv(Q, 8), 256, 0, st>>>(s.gmax, kSeedGroups, plan.G, plan.j, s.margin, s.tau, Q);
        SERT_LAUNCH_CHECK();
        trace.mark("seed");
      }
      ep = TcEpilogue();
      ep.mode = TC_EPI_TOPK;
      ep.tau = s.tau; ep.count = s.count; ep.cand = s.cand; ep.cap = s.cap; ep.row_offset = s.row_begin;
      ep.overflow = s.overflow;
      if (launch_gemm_tc_ld(s.q_split, s.kt, Q, s.ent_split, s.kt, s.rows, 0, s.rows, depth, ep, st)) return -1;
      trace.mark("sweep");
      FinalizeArgs fa;
      fa.Qm = queries_dev; fa.En = s.entities; fa.Q = Q; fa.d = s.d; fa.row_begin = s.row_begin; fa.rows = s.rows;
      fa.cand = s.cand; fa.count = s.count; fa.cap = s.cap; fa.tau = s.tau; fa.margin = s.margin; fa.k = k;
      fa.cmax = cmax; fa.out_idx = out_idx; fa.out_score = out_score; fa.flag = s.overflow; fa.per_warp_bytes = per_warp;
      finalize_kernel<<<cdiv(Q, warps), warps * 32, (size_t)warps * per_warp, st>>>(fa);
      SERT_LAUNCH_CHECK();
      trace.mark("finalize");
      if (deferred) {
        *deferred = 1;
        trace.report("seeded sweep (check deferred)");
        return 0;
      }
      SERT_CUDA(cudaMemcpyAsync(&overflow, s.overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
      SERT_CUDA(cudaStreamSynchronize(st));
      trace.report("seeded sweep");
      if (s.stats) ++s.stats[overflow ? 1 : 0];
      if (!overflow) return 0;
      // a list came up short, overflowed, or held too many near-ties: the multi-chunk sweeps below answer
    }
    if (topk_pass(s, queries_dev, Q, k, true, st, true)) return -1;
    SERT_CUDA(cudaMemcpyAsync(&overflow, s.overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    SERT_CUDA(cudaStreamSynchronize(st));
  }
  if (overflow) {
    if (topk_pass(s, queries_dev, Q, k_sel, true, st)) return -1;
    SERT_CUDA(cudaMemcpyAsync(&overflow, s.overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    SERT_CUDA(cudaStreamSynchronize(st));
    if (overflow && topk_pass(s, queries_dev, Q, k_sel, false, st)) return -1;
  }
  if (tensor && s.rows > 0) {
    // the query's shared-memory copy is padded to whole 16-byte pieces: the compiler reads it in 16-byte loads even on
    // the ragged path (compute-sanitizer flagged the last, partly unused piece when d is no multiple of 4)
    rescore_kernel<<<Q, 256, align_up((size_t)s.d * sizeof(float), 16), st>>>(queries_dev, s.entities, s.d, s.row_begin, s.cand,
                                                                s.count, s.cap);
    SERT_LAUNCH_CHECK();
  }
  return launch_prune(s, Q, k, 1, out_idx, out_score, st);
}

// ---- merge of gathered per-shard lists: [parts][q][k] -> (q,k) ------------------------------------
__global__ void __launch_bounds__(256) merge_kernel(const int32_t *__restrict__ idx, const float *__restrict__ score,
                                                    size_t part_stride, int parts, int Q, int k,
                                                    int32_t *__restrict__ out_idx, float *__restrict__ out_score,
                                                    int n_pow2) {
  extern __shared__ unsigned long long keys[];
  const int q = blockIdx.x;
  const int n = parts * k;
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
    unsigned long long key = 0ull;
    if (i < n) {
      const int p = i / k, j = i % k;
      const size_t src = (size_t)p * part_stride + (size_t)q * k + j;
      const int32_t id = idx[src];
      if (id >= 0) key = make_key(score[src], (unsigned int)id);
    }
    keys[i] = key;
  }
  bitonic_sort_desc(keys, n_pow2);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const unsigned long long key = keys[i];
    if (key != 0ull) {
      out_idx[(size_t)q * k + i] = (int32_t)(0xffffffffu - (unsigned int)(key & 0xffffffffull));
      out_score[(size_t)q * k + i] = unorderable((unsigned int)(key >> 32));
    } else {
      out_idx[(size_t)q * k + i] = -1;
      out_score[(size_t)q * k + i] = -INFINITY;
    }
  }
}

int topk_prepare(int cap) {
  // the prune keeps k <= cap/2 survivors (rounded up to a power of two) in dynamic shared memory
  // once per device: room for the largest list the ABI admits (max_k <= 8192 -> cap <= 32768 -> 128 KB)
  static std::atomic<uint64_t> configured{0};
  (void)cap;
  if (first_use_on_device(configured)) {
    SERT_CUDA(cudaFuncSetAttribute(prune_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    SERT_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SERT_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  return 0;
}

int launch_topk_merge(const int32_t *idx, const float *score, int parts, int Q, int k, int32_t *out_idx,
                      float *out_score, cudaStream_t st, size_t part_stride) {
  if (part_stride == 0) part_stride = (size_t)Q * k;
  SERT_REQUIRE(parts >= 1 && k >= 1, "bad merge shape");
  if (Q == 0) return 0;
  int n_pow2 = 2;
  while (n_pow2 < parts * k) n_pow2 <<= 1;
  const size_t smem = (size_t)n_pow2 * sizeof(unsigned long long);
  SERT_REQUIRE(smem <= 200 * 1024, "merge fan-in too large");
  if (topk_prepare(0)) return -1;             // per-device shared-memory opt-in (was a process-wide static)
  merge_kernel<<<Q, 256, smem, st>>>(idx, score, part_stride, parts, Q, k, out_idx, out_score, n_pow2);
  SERT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sert

// =================================================================================================
using namespace sert;

struct sert_scorer {
  TopkState s;
  cudaStream_t st = nullptr;
  float *queries = nullptr;     // (max_queries, d) staging / normalised queries
  int32_t *out_idx = nullptr;   // (max_queries, max_k)
  float *out_score = nullptr;
  int max_k = 0;
  long long stats[2] = {0, 0};
  long long retries = 0;        // sharded calls repeated with full-length lists
  // row-sharded scoring (sert_scorer_set_comm): gathered per-shard lists, world blocks of [ids (Q,k) | scores (Q,k)]
  sert_comm *comm = nullptr;
  int32_t *gathered = nullptr;
};

// ---- merge of the gathered per-shard lists (row-sharded scoring), one warp per query -----------------------------
// Every part's list is sorted (best first) and keys are unique, so an entry's place in the merged order is its place
// in its own list plus, for every other part, the number of that part's keys above it -- one binary search each; no
// sort.  Entries whose place is below k are written straight to their slot.  When the parts sent fewer than k
// entries each (k_loc < k, scorer_topk), a part ALL of whose entries made the global top k may hold further rows that
// belong there: that sets *flag and the call is repeated with full-length lists.
namespace sert {
__global__ void __launch_bounds__(256) merge_sorted_kernel(const int32_t *__restrict__ gathered, size_t part_stride,
                                                           int parts, int Q, int k_loc, int k,
                                                           int32_t *__restrict__ out_idx, float *__restrict__ out_score,
                                                           int *__restrict__ flag) {
  extern __shared__ unsigned long long mkeys[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= Q) return;
  const int n = parts * k_loc;
  unsigned long long *keys = mkeys + (size_t)warp * n;
  int valid = 0;
  for (int e = lane; e < n; e += 32) {
    const int p = e / k_loc, i = e - p * k_loc;
    const int32_t *blk = gathered + (size_t)p * part_stride;
    const int32_t id = blk[(size_t)q * k_loc + i];
    const float sc = reinterpret_cast<const float *>(blk + (size_t)Q * k_loc)[(size_t)q * k_loc + i];
    keys[e] = id >= 0 ? make_key(sc, (unsigned int)id) : 0ull;
    valid += id >= 0 ? 1 : 0;
  }
  valid = __reduce_add_sync(0xffffffffu, valid);
  __syncwarp();
  bool more = false;
  for (int e = lane; e < n; e += 32) {
    const unsigned long long key = keys[e];
    if (key == 0ull) continue;
    const int p = e / k_loc, i = e - p * k_loc;
    int place = i;
    for (int o = 0; o < parts; ++o) {
      if (o == p) continue;
      const unsigned long long *other = keys + o * k_loc;
      int lo = 0, hi = k_loc;                                // first position of `other` holding a smaller key
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (other[mid] > key) lo = mid + 1; else hi = mid;
      }
      place += lo;
    }
    if (place < k) {
      out_idx[(size_t)q * k + place] = (int32_t)key_row(key);
      out_score[(size_t)q * k + place] = key_score(key);
      if (i == k_loc - 1 && k_loc < k) more = true;         // this part's whole list is inside the global top k
    }
  }
  for (int i = valid + lane; i < k; i += 32) {               // fewer than k rows in all shards together
    out_idx[(size_t)q * k + i] = -1;
    out_score[(size_t)q * k + i] = -INFINITY;
  }
  if (__any_sync(0xffffffffu, more) && lane == 0) *flag = 1;
}
}  // namespace sert

static int parts_bytes(int world, int k_loc) { return (int)sert::align_up((size_t)world * k_loc * 8, 16); }

// Entries a shard contributes to a global top k over `world` equal shards: mean k / world plus six standard
// deviations of the binomial count, so that on rows spread evenly over the shards one attempt in ~1e8 (per query and
// shard) is repeated; never more than k.
static int shard_list_length(int k, int world) {
  const double mean = (double)k / world;
  const int k_loc = (int)ceil(mean + 6.0 * sqrt(mean * (1.0 - 1.0 / world)) + 2.0);
  return std::min(k, std::max(k_loc, 8));
}

// Top k of the GLOBAL entity matrix on every rank: local sweep into this rank's block of the gather buffer, ONE
// all-gather of the packed (row id, score)[Q, k_loc] lists, merge.  Keys carry global row ids and ties order by row
// id, so the merged list equals the single-device list.  First attempt: every shard sends its best k_loc <= k rows
// (shard_list_length) and the seeded local sweep is not checked on the host; the blocks carry each rank's verdict,
// the merge adds its own, and ONE synchronisation at the end reads them.  All ranks see the same gathered data,
// hence take the same decision: on any doubt the call is repeated with k rows per shard and checked local sweeps.
static int scorer_topk(sert_scorer *s, const float *q_dev, int q, int k, int32_t *out_idx, float *out_score) {
  if (s->comm == nullptr || s->comm->world == 1)
    return sert::topk_sweep(s->s, q_dev, q, k, out_idx, out_score, s->st, nullptr, true);
  if (q == 0) return 0;
  const int world = s->comm->world;
  bool seeded_ok = true;          // false once some rank's seeded sweep reported a failure: the retry uses chunked sweeps
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int k_loc = attempt == 0 ? shard_list_length(k, world) : k;
    const size_t block = (size_t)2 * q * k_loc + 4;           // 32-bit words per rank: ids, scores, verdict + padding
    int32_t *mine = s->gathered + (size_t)s->comm->rank * block;
    sert::PhaseTrace trace(s->st);
    trace.mark("start");
    int deferred = 0;
    if (sert::topk_sweep(s->s, q_dev, q, k_loc, mine, reinterpret_cast<float *>(mine + (size_t)q * k_loc), s->st,
                         attempt == 0 ? &deferred : nullptr, seeded_ok))
      return -1;
    if (deferred) {
      SERT_CUDA(cudaMemcpyAsync(mine + (size_t)2 * q * k_loc, s->s.overflow, sizeof(int), cudaMemcpyDeviceToDevice, s->st));
    } else {
      SERT_CUDA(cudaMemsetAsync(mine + (size_t)2 * q * k_loc, 0, sizeof(int), s->st));
    }
    SERT_CUDA(cudaMemsetAsync(s->s.overflow + 3, 0, sizeof(int), s->st));      // the merge's verdict
    trace.mark("local sweep");
    if (sert::comm_all_gather(s->comm, mine, s->gathered, block * 4, s->st)) return -1;
    trace.mark("all-gather");
    const int per_warp = parts_bytes(world, k_loc);
    const int warps = std::max(1, std::min(8, (96 * 1024) / per_warp));
    sert::merge_sorted_kernel<<<sert::cdiv(q, warps), warps * 32, (size_t)warps * per_warp, s->st>>>(
        s->gathered, block, world, q, k_loc, k, out_idx, out_score, s->s.overflow + 3);
    SERT_LAUNCH_CHECK();
    trace.mark("merge");
    // verdicts: one word per rank inside the gathered blocks + the merge's
    int merge_flag = 0;
    std::vector<int> local_flags((size_t)world, 0);
    SERT_CUDA(cudaMemcpyAsync(&merge_flag, s->s.overflow + 3, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    SERT_CUDA(cudaMemcpy2DAsync(local_flags.data(), sizeof(int), s->gathered + (size_t)2 * q * k_loc, block * 4,
                                sizeof(int), world, cudaMemcpyDeviceToHost, s->st));
    SERT_CUDA(cudaStreamSynchronize(s->st));
    trace.report(attempt == 0 ? "sharded top-k" : "sharded top-k (full-length retry)");
    bool redo = merge_flag != 0;
    for (int r = 0; r < world; ++r)
      if (local_flags[r] != 0) { redo = true; seeded_ok = false; }
    if (deferred) ++s->stats[local_flags[s->comm->rank] ? 1 : 0];
    if (!redo) return 0;
    ++s->retries;
  }
  sert::set_error("sharded top-k: the full-length attempt reported a failure");
  return -1;
}

namespace sert {
static size_t carve_scorer(sert_scorer &sc, void *base, int64_t rows, int d, int max_queries, int max_k) {
  char *b = static_cast<char *>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = align_up(off, 256);
    char *p = b ? b + off : nullptr;
    off += bytes;
    return p;
  };
  int cap = 8192;                       // chunk = cap/2 entity rows per score-tile launch
  while (cap / 2 < max_k + 16) cap <<= 1;
  sc.s.cap = cap;
  sc.s.rows = rows;
  sc.s.d = d;
  sc.s.max_queries = max_queries;
  sc.max_k = max_k;
  sc.s.entities = reinterpret_cast<float *>(take((size_t)rows * d * sizeof(float)));
  sc.s.cand = reinterpret_cast<unsigned long long *>(take((size_t)max_queries * cap * sizeof(unsigned long long)));
  sc.s.tau = reinterpret_cast<unsigned long long *>(take((size_t)max_queries * sizeof(unsigned long long)));
  sc.s.count = reinterpret_cast<int *>(take((size_t)max_queries * sizeof(int)));
  sc.s.overflow = reinterpret_cast<int *>(take(8 * sizeof(int)));      // [3]: merge verdict; [4..6]: scratch of row_norm_max_kernel
  sc.s.margin = reinterpret_cast<float *>(take((size_t)max_queries * sizeof(float)));
  sc.s.gmax = reinterpret_cast<float *>(take((size_t)max_queries * kSeedGroups * sizeof(float)));
  sc.queries = reinterpret_cast<float *>(take((size_t)max_queries * d * sizeof(float)));
  sc.s.terms = 3;
  sc.s.kt = sc.s.terms * tc_padded_k(d);
  sc.s.ent_split = reinterpret_cast<__nv_bfloat16 *>(take((size_t)rows * sc.s.kt * sizeof(__nv_bfloat16)));
  sc.s.q_split = reinterpret_cast<__nv_bfloat16 *>(take((size_t)max_queries * sc.s.kt * sizeof(__nv_bfloat16)));
  sc.out_idx = reinterpret_cast<int32_t *>(take((size_t)max_queries * max_k * sizeof(int32_t)));
  sc.out_score = reinterpret_cast<float *>(take((size_t)max_queries * max_k * sizeof(float)));
  return align_up(off, 256);
}
}  // namespace sert

extern "C" {

int sert_scorer_arena_bytes(int64_t rows, int32_t d, int32_t max_queries, int32_t max_k, size_t *bytes) {
  SERT_REQUIRE(bytes, "null argument");
  SERT_REQUIRE(rows >= 0 && d > 0 && max_queries > 0 && max_k > 0, "bad scorer shape");
  SERT_REQUIRE(max_k <= 8192, "max_k above 8192 is not supported");
  sert_scorer tmp;
  *bytes = carve_scorer(tmp, nullptr, rows, d, max_queries, max_k);
  return 0;
}

int sert_scorer_create(const float *entities_host, int64_t rows, int32_t d, int64_t row_begin, int32_t normalise,
                       int32_t max_queries, int32_t max_k, void *arena_dev, size_t arena_bytes, void *stream,
                       sert_scorer **out) {
  SERT_REQUIRE(out && arena_dev && (entities_host || rows == 0), "null argument");
  SERT_REQUIRE(rows >= 0 && d > 0 && max_queries > 0 && max_k > 0 && max_k <= 8192, "bad scorer shape");
  SERT_REQUIRE(row_begin >= 0 && row_begin + rows < (1ll << 31), "row ids must fit in int32");
  int ndev = 0;
  SERT_CUDA(cudaGetDeviceCount(&ndev));
  SERT_REQUIRE(ndev > 0, "no CUDA device: libsert_b200 has no CPU fallback");
  sert_scorer *sc = new sert_scorer();
  const size_t need = carve_scorer(*sc, arena_dev, rows, d, max_queries, max_k);
  if (need > arena_bytes) {
    delete sc;
    set_error("scorer arena too small: need " + std::to_string(need) + " bytes");
    return -1;
  }
  sc->st = static_cast<cudaStream_t>(stream);
  sc->s.row_begin = row_begin;
  sc->s.stats = sc->stats;
  if (topk_prepare(sc->s.cap)) { delete sc; return -1; }
  if (rows > 0) {
    cudaError_t e = cudaMemcpyAsync(sc->s.entities, entities_host, (size_t)rows * d * sizeof(float),
                                    cudaMemcpyHostToDevice, sc->st);
    if (e != cudaSuccess) { delete sc; set_error(cudaGetErrorString(e)); return -1; }
    if (normalise && launch_normalise_rows(sc->s.entities, sc->s.entities, rows, d, sc->st)) { delete sc; return -1; }
    // bf16x3 split of the (normalised) entity rows: the B operand of the tcgen05 scoring GEMM
    if (launch_split_bf16(sc->s.entities, rows, d, d, sc->s.terms, SPLIT_B, sc->s.ent_split, sc->st)) {
      delete sc;
      return -1;
    }
  }
  sc->s.mode = SCORE_TENSOR;
  unsigned int norm_bits[3] = {0u, 0u, 0u};
  if (rows > 0) {
    unsigned int *scratch = reinterpret_cast<unsigned int *>(sc->s.overflow + 4);
    cudaMemsetAsync(scratch, 0, 3 * sizeof(unsigned int), sc->st);
    row_norm_max_kernel<<<cdiv(rows, 8), 256, 0, sc->st>>>(sc->s.entities, rows, d, scratch);
    count_launch();
    cudaMemcpyAsync(norm_bits, scratch, 3 * sizeof(unsigned int), cudaMemcpyDeviceToHost, sc->st);
  }
  cudaError_t e = cudaStreamSynchronize(sc->st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { delete sc; set_error(cudaGetErrorString(e)); return -1; }
  memcpy(&sc->s.ent_norm_max, &norm_bits[0], sizeof(float));
  memcpy(&sc->s.ent_err_max, &norm_bits[1], sizeof(float));
  memcpy(&sc->s.ent_err2_max, &norm_bits[2], sizeof(float));
  sc->s.coarse_blocks = (sc->s.kt / sc->s.terms <= 128 && getenv("SERT_COARSE_BLOCKS1") == nullptr) ? 2 : 1;
  *out = sc;
  return 0;
}

int sert_scorer_set_mode(sert_scorer *s, int32_t mode) {
  SERT_REQUIRE(s, "null scorer");
  SERT_REQUIRE(mode >= 0 && mode <= 3, "unknown scoring mode");
  s->s.coarse = mode == 2 ? 0 : 1;      // 2: tensor cores without the coarse first sweep (bf16x3 scores throughout)
  s->s.seeded = mode == 3 ? 0 : 1;      // 3: coarse sweep in growing chunks with a prune after each (no threshold seed)
  s->s.mode = mode == SCORE_FMA ? SCORE_FMA : SCORE_TENSOR;
  return 0;
}

int sert_scorer_plan(sert_scorer *s, int32_t k, int32_t *group_rows, int32_t *groups, int32_t *rank,
                     int64_t *tile_stride, double *expected_survivors) {
  SERT_REQUIRE(s && group_rows && groups && rank && tile_stride && expected_survivors, "null argument");
  const sert::SweepPlan p = sert::topk_plan(s->s.rows, k, s->s.cap);
  *group_rows = p.seeded ? p.g : 0;
  *groups = p.G; *rank = p.j; *tile_stride = p.stride; *expected_survivors = p.T;
  return 0;
}

int sert_scorer_stats(sert_scorer *s, int64_t *seeded_sweeps, int64_t *fallback_sweeps) {
  SERT_REQUIRE(s && seeded_sweeps && fallback_sweeps, "null argument");
  *seeded_sweeps = s->stats[0];
  *fallback_sweeps = s->stats[1];
  return 0;
}

int sert_scorer_set_comm(sert_scorer *s, sert_comm *comm) {
  SERT_REQUIRE(s, "null scorer");
  if (s->gathered) { cudaStreamSynchronize(s->st); cudaFree(s->gathered); s->gathered = nullptr; }
  s->comm = comm;
  if (comm != nullptr && comm->world > 1)
    SERT_CUDA(cudaMalloc(&s->gathered, (size_t)comm->world * ((size_t)2 * s->s.max_queries * s->max_k + 4) * sizeof(int32_t)));
  return 0;
}

int sert_scorer_destroy(sert_scorer *s) {
  if (s) {
    cudaStreamSynchronize(s->st);
    if (s->gathered) cudaFree(s->gathered);
    delete s;
  }
  return 0;
}

int sert_scorer_topk_dev(sert_scorer *s, const float *queries_dev, int32_t q, int32_t normalise_q, int32_t k,
                         int32_t *out_idx_dev, float *out_score_dev) {
  SERT_REQUIRE(s && (queries_dev || q == 0) && out_idx_dev && out_score_dev, "null argument");
  SERT_REQUIRE(q <= s->s.max_queries, "more queries than the scorer's max_queries");
  SERT_REQUIRE(k >= 1 && k <= s->max_k, "k exceeds the scorer's max_k");
  const float *qq = queries_dev;
  if (normalise_q) {
    if (launch_normalise_rows(queries_dev, s->queries, q, s->s.d, s->st)) return -1;
    qq = s->queries;
  }
  return scorer_topk(s, qq, q, k, out_idx_dev, out_score_dev);
}

int sert_scorer_topk_host(sert_scorer *s, const float *queries_host, int32_t q, int32_t normalise_q, int32_t k,
                          int32_t *out_idx_host, float *out_score_host) {
  SERT_REQUIRE(s && (queries_host || q == 0) && out_idx_host && out_score_host, "null argument");
  SERT_REQUIRE(q <= s->s.max_queries, "more queries than the scorer's max_queries");
  SERT_REQUIRE(k >= 1 && k <= s->max_k, "k exceeds the scorer's max_k");
  if (q == 0) return 0;
  SERT_CUDA(cudaMemcpyAsync(s->queries, queries_host, (size_t)q * s->s.d * sizeof(float), cudaMemcpyHostToDevice,
                            s->st));
  if (normalise_q && launch_normalise_rows(s->queries, s->queries, q, s->s.d, s->st)) return -1;
  if (scorer_topk(s, s->queries, q, k, s->out_idx, s->out_score)) return -1;
  SERT_CUDA(cudaMemcpyAsync(out_idx_host, s->out_idx, (size_t)q * k * sizeof(int32_t), cudaMemcpyDeviceToHost, s->st));
  SERT_CUDA(cudaMemcpyAsync(out_score_host, s->out_score, (size_t)q * k * sizeof(float), cudaMemcpyDeviceToHost,
                            s->st));
  SERT_CUDA(cudaStreamSynchronize(s->st));
  return 0;
}

int sert_scorer_scores_host(sert_scorer *s, const float *queries_host, int32_t q, int32_t normalise_q,
                            float *out_host) {
  SERT_REQUIRE(s && (queries_host || q == 0) && out_host, "null argument");
  SERT_REQUIRE(s->s.rows < (1ll << 31), "too many rows");
  const long long rows = s->s.rows;
  if (q == 0 || rows == 0) return 0;
  // the candidate buffer doubles as scratch for the dense (q_chunk, rows) score block
  float *scratch = reinterpret_cast<float *>(s->s.cand);
  const long long cap_floats = (long long)s->s.max_queries * s->s.cap * 2;
  long long q_chunk = std::min<long long>(std::min<long long>(q, s->s.max_queries), cap_floats / rows);
  SERT_REQUIRE(q_chunk >= 1, "scorer arena too small for a full score row; raise max_queries");
  for (long long q0 = 0; q0 < q; q0 += q_chunk) {
    const int n = (int)std::min<long long>(q_chunk, q - q0);
    SERT_CUDA(cudaMemcpyAsync(s->queries, queries_host + q0 * s->s.d, (size_t)n * s->s.d * sizeof(float),
                              cudaMemcpyHostToDevice, s->st));
    if (normalise_q && launch_normalise_rows(s->queries, s->queries, n, s->s.d, s->st)) return -1;
    if (launch_gemm_f32(s->queries, s->s.entities, scratch, n, (int)rows, s->s.d, false, true, s->s.d, s->s.d,
                        (int)rows, EPI_STORE, nullptr, 1, s->st))
      return -1;
    SERT_CUDA(cudaMemcpyAsync(out_host + q0 * rows, scratch, (size_t)n * rows * sizeof(float),
                              cudaMemcpyDeviceToHost, s->st));
    SERT_CUDA(cudaStreamSynchronize(s->st));
  }
  return 0;
}

int sert_topk_merge_dev(const int32_t *idx_dev, const float *score_dev, int32_t parts, int32_t q, int32_t k,
                        int32_t *out_idx_dev, float *out_score_dev, void *stream) {
  SERT_REQUIRE(idx_dev && score_dev && out_idx_dev && out_score_dev, "null argument");
  return launch_topk_merge(idx_dev, score_dev, parts, q, k, out_idx_dev, out_score_dev,
                           static_cast<cudaStream_t>(stream));
}

}  // extern "C"
