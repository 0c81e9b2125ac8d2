// Embedding gather / pool / scatter and the negative-sampling loss kernels (HBM-bound row work).
//
// Reference math: sert/models.py:180 (gather), :226,1051 (window mean), :893-902 (sigmoid distance),
// :981-1009 (entity gather), :1072-1098 (loss), :1065-1068 (tanh clip).  Backward forms follow
// SURVEY.md Appendix A.2 in the general (clipped) regime; T.clip's gradient is 1 on the closed
// interval and 0 outside.
#include "kernels.cuh"

namespace sert {

// ------------------------------------------------------------------------------------------------
// gather + pool: one warp per instance, lanes stride over 16-byte chunks of the row so every
// row read is a fully coalesced 512-byte (d=128) request; W independent loads in flight per lane.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_pool_kernel(const int32_t *__restrict__ x,
                                                          const float4 *__restrict__ R,
                                                          float4 *__restrict__ out, int B, int W, int d4,
                                                          float denom) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int32_t *xi = x + (size_t)warp * W;
  // warp-uniform trip counts: every lane takes part in the index shuffles, loads are predicated
  for (int c0 = 0; c0 < d4; c0 += 32) {
    const int c = c0 + lane;
    const bool active = c < d4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w0 = 0; w0 < W; w0 += 32) {
      const int nw = min(32, W - w0);
      const int idx = (lane < nw) ? __ldg(xi + w0 + lane) : 0;
#pragma unroll 5
      for (int w = 0; w < nw; ++w) {
        const int r = __shfl_sync(0xffffffffu, idx, w);
        if (active) {
          const float4 v = __ldg(R + (size_t)r * d4 + c);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
    }
    if (active) {
      acc.x /= denom; acc.y /= denom; acc.z /= denom; acc.w /= denom;
      out[(size_t)warp * d4 + c] = acc;
    }
  }
}

int launch_gather_pool(const int32_t *x, const float *R, float *out, int B, int W, int d, float denom,
                       cudaStream_t st) {
  SERT_REQUIRE(d % 4 == 0, "representation size must be a multiple of 4");
  if (B == 0) return 0;
  const int warps_per_block = 8;
  gather_pool_kernel<<<cdiv(B, warps_per_block), warps_per_block * 32, 0, st>>>(
      x, reinterpret_cast<const float4 *>(R), reinterpret_cast<float4 *>(out), B, W, d / 4, denom);
  SERT_LAUNCH_CHECK();
  return 0;
}

// unpooled gather: one warp per output row
__global__ void __launch_bounds__(256) gather_rows_kernel(const int32_t *__restrict__ x,
                                                          const float4 *__restrict__ R,
                                                          float4 *__restrict__ out, long long rows, int d4) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int r = __ldg(x + warp);
  for (int c = lane; c < d4; c += 32) out[warp * d4 + c] = __ldg(R + (size_t)r * d4 + c);
}

int launch_gather_rows(const int32_t *x, const float *R, float *out, int64_t rows, int d, cudaStream_t st) {
  SERT_REQUIRE(d % 4 == 0, "representation size must be a multiple of 4");
  if (rows == 0) return 0;
  gather_rows_kernel<<<cdiv(rows, 8), 256, 0, st>>>(x, reinterpret_cast<const float4 *>(R),
                                                    reinterpret_cast<float4 *>(out), rows, d / 4);
  SERT_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// scatter-add of per-instance row gradients into the word table gradient (duplicates accumulate).
// red.global.add.v4.f32: the L2 atomic units absorb collisions; touched rows are stamped so the
// dense optimiser only reads gradient rows that exist.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scatter_rows_kernel(const int32_t *__restrict__ x,
                                                           const float4 *__restrict__ dh,
                                                           float *__restrict__ gR,
                                                           uint32_t *__restrict__ flagR, uint32_t stamp,
                                                           int B, int W, int d4, float denom, int row_lo, int row_hi) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int32_t *xi = x + (size_t)warp * W;
  for (int w0 = 0; w0 < W; w0 += 32) {
    const int nw = min(32, W - w0);
    const int idx = (lane < nw) ? __ldg(xi + w0 + lane) : 0;
    if (lane < nw && idx >= row_lo && idx < row_hi) flagR[idx] = stamp;
    for (int c0 = 0; c0 < d4; c0 += 32) {
      const int c = c0 + lane;
      const bool active = c < d4;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active) {
        g = dh[(size_t)warp * d4 + c];
        g.x /= denom; g.y /= denom; g.z /= denom; g.w /= denom;
      }
      for (int w = 0; w < nw; ++w) {
        const int r = __shfl_sync(0xffffffffu, idx, w);
        if (active && r >= row_lo && r < row_hi) red_add_f4(gR + ((size_t)r * d4 + c) * 4, g);
      }
    }
  }
}

// table shards with look-ahead: stamps the rows the next batch reads (kernels.cuh: launch_mark_needed)
__global__ void __launch_bounds__(256) mark_needed_kernel(const int32_t *__restrict__ x, const int32_t *__restrict__ y,
                                                          const int32_t *__restrict__ neg, long long nx, long long ny,
                                                          long long nn, uint32_t *__restrict__ need_r,
                                                          uint32_t *__restrict__ need_e, uint32_t stamp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nx) need_r[__ldg(x + i)] = stamp;
  else if (i < nx + ny) need_e[__ldg(y + (i - nx))] = stamp;
  else if (i < nx + ny + nn) need_e[__ldg(neg + (i - nx - ny))] = stamp;
}

int launch_mark_needed(const int32_t *x, const int32_t *y, const int32_t *neg, int B, int W, int k, uint32_t *need_r,
                       uint32_t *need_e, uint32_t stamp, cudaStream_t st) {
  const long long nx = (long long)B * W, ny = B, nn = (long long)B * k;
  if (nx + ny + nn == 0) return 0;
  mark_needed_kernel<<<cdiv(nx + ny + nn, 256), 256, 0, st>>>(x, y, neg, nx, ny, nn, need_r, need_e, stamp);
  SERT_LAUNCH_CHECK();
  return 0;
}

// instance shards: which ranks' instances of the next batch read a row (kernels.cuh: launch_mark_needed_by)
struct InstanceBounds { int b[kMaxOwners + 1]; };
__global__ void __launch_bounds__(256) mark_needed_by_kernel(const int32_t *__restrict__ x, const int32_t *__restrict__ y,
                                                             const int32_t *__restrict__ neg, long long nx, long long ny,
                                                             long long nn, int W, int k, uint32_t *__restrict__ need_r,
                                                             uint32_t *__restrict__ need_e, InstanceBounds ib, int n_ranks) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int inst, row;
  uint32_t *dst;
  if (i < nx) { inst = (int)(i / W); row = __ldg(x + i); dst = need_r; }
  else if (i < nx + ny) { inst = (int)(i - nx); row = __ldg(y + inst); dst = need_e; }
  else if (i < nx + ny + nn) { inst = (int)((i - nx - ny) / k); row = __ldg(neg + (i - nx - ny)); dst = need_e; }
  else return;
  int r = 0;
  for (int q = 1; q < n_ranks; ++q) r += inst >= ib.b[q] ? 1 : 0;
  const uint32_t bit = 1u << r;
  if ((__ldcg(dst + row) & bit) == 0u) atomicOr(dst + row, bit);
}

int launch_mark_needed_by(const int32_t *x, const int32_t *y, const int32_t *neg, int B, int W, int k, uint32_t *need_r,
                          uint32_t *need_e, const int *i_bound, int n_ranks, cudaStream_t st) {
  SERT_REQUIRE(n_ranks >= 1 && n_ranks <= kMaxOwners, "bad rank count");
  const long long nx = (long long)B * W, ny = B, nn = (long long)B * k;
  if (nx + ny + nn == 0) return 0;
  InstanceBounds ib;
  for (int r = 0; r <= n_ranks; ++r) ib.b[r] = i_bound[r];
  mark_needed_by_kernel<<<cdiv(nx + ny + nn, 256), 256, 0, st>>>(x, y, neg, nx, ny, nn, W, k, need_r, need_e, ib, n_ranks);
  SERT_LAUNCH_CHECK();
  return 0;
}

int launch_scatter_rows(const int32_t *x, const float *dh, float *gR, uint32_t *flagR, uint32_t stamp,
                        int B, int W, int d, float denom, cudaStream_t st, int row_lo, int row_hi) {
  SERT_REQUIRE(d % 4 == 0, "representation size must be a multiple of 4");
  if (B == 0) return 0;
  scatter_rows_kernel<<<cdiv(B, 8), 256, 0, st>>>(x, reinterpret_cast<const float4 *>(dh), gR, flagR,
                                                  stamp, B, W, d / 4, denom, row_lo, row_hi);
  SERT_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// negative-sampling loss forward + backward.  One warp per instance; lanes own 16-byte chunks of
// the de-vector.  Rows are fetched GROUP at a time before any reduction so GROUP*MAXC independent
// 16-byte loads are in flight per lane.
// ------------------------------------------------------------------------------------------------
template <int MAXC, int GROUP, bool TRAIN>
__global__ void __launch_bounds__(256) vs_nce_kernel(VsNceArgs a) {
  const int warp_in_block = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + warp_in_block;
  const int d4 = a.de >> 2;
  __shared__ double s_loss[8];
  double my_loss = 0.0;

  if (i < a.B) {
    const float4 *t4 = reinterpret_cast<const float4 *>(a.t) + (size_t)i * d4;
    const float4 *E4 = reinterpret_cast<const float4 *>(a.Eemb);
    float4 tt[MAXC], u[MAXC], du[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + 32 * c;
      tt[c] = (ch < d4) ? t4[ch] : make_float4(0.f, 0.f, 0.f, 0.f);
      u[c].x = clipf_(tt[c].x, SERT_TANH_LO, SERT_TANH_HI);
      u[c].y = clipf_(tt[c].y, SERT_TANH_LO, SERT_TANH_HI);
      u[c].z = clipf_(tt[c].z, SERT_TANH_LO, SERT_TANH_HI);
      u[c].w = clipf_(tt[c].w, SERT_TANH_LO, SERT_TANH_HI);
      du[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.dbg_u != nullptr && ch < d4) reinterpret_cast<float4 *>(a.dbg_u)[(size_t)i * d4 + ch] = u[c];
    }
    const float coef_scale = TRAIN ? ((a.w ? __ldg(a.w + i) : 1.0f) * a.inv_B) : 0.0f;
    const int yi = __ldg(a.y + i);
    const int32_t *negi = a.neg + (size_t)i * a.k;
    float ell = 0.0f;

    for (int j0 = 0; j0 <= a.k; j0 += GROUP) {
      int rows[GROUP];
      float4 e[GROUP][MAXC];
      float dots[GROUP];
#pragma unroll
      for (int g = 0; g < GROUP; ++g) {
        const int j = j0 + g;
        rows[g] = (j > a.k) ? -1 : (j == 0 ? yi : __ldg(negi + j - 1));
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          const int ch = lane + 32 * c;
          e[g][c] = (rows[g] >= 0 && ch < d4) ? __ldg(E4 + (size_t)rows[g] * d4 + ch)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int g = 0; g < GROUP; ++g) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          s += e[g][c].x * u[c].x + e[g][c].y * u[c].y + e[g][c].z * u[c].z + e[g][c].w * u[c].w;
        dots[g] = s;
      }
#pragma unroll
      for (int g = 0; g < GROUP; ++g) dots[g] = warp_sum(dots[g]);
#pragma unroll
      for (int g = 0; g < GROUP; ++g) {
        if (rows[g] < 0) continue;
        const int j = j0 + g;
        const float score = dots[g];
        const float sg = sigmoidf_(score);
        const float cl = clipf_(sg, SERT_CLIP_LO, SERT_CLIP_HI);
        const bool inside = (sg >= SERT_CLIP_LO) && (sg <= SERT_CLIP_HI);
        float coef;
        if (j == 0) {
          ell -= logf(cl);
          coef = inside ? (-coef_scale / cl) * sg * (1.0f - sg) : 0.0f;
        } else {
          ell -= logf(1.0f - cl);
          coef = inside ? (coef_scale / (1.0f - cl)) * sg * (1.0f - sg) : 0.0f;
        }
        if (a.dbg_scores != nullptr && lane == 0) a.dbg_scores[(size_t)i * (a.k + 1) + j] = score;
        if (TRAIN) {
          const bool mine = a.own.entity(rows[g]);     // table shards: another rank forms this row's gradient
          if (lane == 0 && mine) a.flagE[rows[g]] = a.stamp;
#pragma unroll
          for (int c = 0; c < MAXC; ++c) {
            const int ch = lane + 32 * c;
            du[c].x += coef * e[g][c].x; du[c].y += coef * e[g][c].y;
            du[c].z += coef * e[g][c].z; du[c].w += coef * e[g][c].w;
            if (ch < d4 && mine)
              red_add_f4(a.gE + ((size_t)rows[g] * d4 + ch) * 4,
                         make_float4(coef * u[c].x, coef * u[c].y, coef * u[c].z, coef * u[c].w));
          }
        }
      }
    }
    if (a.dbg_ell != nullptr && lane == 0) a.dbg_ell[i] = ell;
    if (TRAIN) {
      float4 *da4 = reinterpret_cast<float4 *>(a.da) + (size_t)i * d4;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch >= d4) continue;
        float4 o;
        o.x = (tt[c].x >= SERT_TANH_LO && tt[c].x <= SERT_TANH_HI) ? du[c].x * (1.0f - tt[c].x * tt[c].x) : 0.f;
        o.y = (tt[c].y >= SERT_TANH_LO && tt[c].y <= SERT_TANH_HI) ? du[c].y * (1.0f - tt[c].y * tt[c].y) : 0.f;
        o.z = (tt[c].z >= SERT_TANH_LO && tt[c].z <= SERT_TANH_HI) ? du[c].z * (1.0f - tt[c].z * tt[c].z) : 0.f;
        o.w = (tt[c].w >= SERT_TANH_LO && tt[c].w <= SERT_TANH_HI) ? du[c].w * (1.0f - tt[c].w * tt[c].w) : 0.f;
        da4[ch] = o;
      }
      my_loss = (double)((a.w ? __ldg(a.w + i) : 1.0f) * ell);
    } else {
      my_loss = (double)ell;
    }
  }
  if (lane == 0) s_loss[warp_in_block] = my_loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) s += s_loss[wv];
    atomicAdd(a.loss_acc, s);
  }
}

template <int MAXC, int GROUP>
static int launch_vs_nce_t(const VsNceArgs &a, cudaStream_t st) {
  const int blocks = cdiv(a.B, 8);
  if (a.train)
    vs_nce_kernel<MAXC, GROUP, true><<<blocks, 256, 0, st>>>(a);
  else
    vs_nce_kernel<MAXC, GROUP, false><<<blocks, 256, 0, st>>>(a);
  SERT_LAUNCH_CHECK();
  return 0;
}

int launch_vs_nce(const VsNceArgs &a, cudaStream_t st) {
  SERT_REQUIRE(a.de % 4 == 0, "entity representation size must be a multiple of 4");
  SERT_REQUIRE(a.de <= 1024, "entity representation size above 1024 is not supported");
  if (a.B == 0) return 0;
  const int d4 = a.de / 4;
  if (d4 <= 32) return launch_vs_nce_t<1, 4>(a, st);
  if (d4 <= 64) return launch_vs_nce_t<2, 4>(a, st);
  if (d4 <= 96) return launch_vs_nce_t<3, 2>(a, st);
  if (d4 <= 128) return launch_vs_nce_t<4, 2>(a, st);
  return launch_vs_nce_t<8, 1>(a, st);
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 negatives: uniform over [0,E) with replacement, not excluding the positive
// (sert/models.py:927-931,956-973).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
                                             uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__global__ void sample_negatives_kernel(int32_t *__restrict__ out, long long n, unsigned long long E,
                                        unsigned long long seed, unsigned long long step) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // 4 outputs per thread
  if (q * 4 >= n) return;
  uint32_t c0 = (uint32_t)q, c1 = (uint32_t)(q >> 32), c2 = (uint32_t)step, c3 = (uint32_t)(step >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const uint32_t r[4] = {c0, c1, c2, c3};
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (q * 4 + j < n) out[q * 4 + j] = (int32_t)(((unsigned long long)r[j] * E) >> 32);
}

int launch_sample_negatives(int32_t *out, int64_t n, int64_t E, uint64_t seed, uint64_t step,
                            cudaStream_t st) {
  if (n == 0) return 0;
  SERT_REQUIRE(E > 0 && E < (1ll << 31), "entity count out of range");
  const long long threads = (n + 3) / 4;
  sample_negatives_kernel<<<cdiv(threads, 256), 256, 0, st>>>(out, n, (unsigned long long)E, seed, step);
  SERT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sert
