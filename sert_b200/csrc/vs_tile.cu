// Vector-space training step, forward + backward, ONE CTA PER TILE OF 8 INSTANCES, d_e = 128 and d_w a multiple of 4 up to
// 384 in NW = ceil(d_w / 128) chunks of 128 floats per lane (BASELINE configs[1]: 128 / 128, NW = 1;
// product-search.sh:102-147: 300 / 128, NW = 3; sert/models.py:1044-1098).  Same stages and arithmetic as csrc/vs_warp.cu, re-cut around
// what the profiles of that kernel and of the earlier versions of this one showed (profiles/ncu_vs_tile_r1*.txt,
// tools/red_probe.cu):
//   * warp g owns instance g for the gather, the loss and the two scatters (one 512-byte row per request);
//   * the two 128 x 128 matrix-vector products are computed for the 8 instances together, split along K: warp w
//     reads rows 16w..16w+15 of the matrix straight from global memory (coalesced 512-byte requests, each row
//     fetched once per CTA), multiplies them into 8 x 128 partial products held in registers, and the 8 partials
//     are summed through shared memory so that warp g ends up with the full row of instance g.  Earlier cuts read
//     64-byte slices of the matrix per warp, first through L1/L2 (bound by bytes in flight), then from a shared-memory
//     copy (bound by the 4 shared-memory wavefronts every LDS.128 costs whatever it broadcasts);
//   * the 1+k entity rows of an instance are prefetched into L2 when its indices arrive and copied to shared
//     memory with cp.async (LDGSTS) behind the projection, into the space the partial products occupied;
//   * the 1+k score partials of an instance are reduced with a 16-value transposing butterfly (16 shuffles
//     instead of 5 per score), after which lane 2j holds score j and the sigmoid / log / clip arithmetic of all
//     rows runs once, in parallel over the lanes, instead of once per row;
//   * additions to the gradient rows of very frequent words go to per-CTA-group copies (kernels.cuh: hot_slot);
//   * one extra CTA writes the PREVIOUS step's loss, so that no finalisation kernel sits between two steps.
// 32 resident warps per SM (vs_warp: 14), all 512 CTAs of a 4096-instance batch resident at once.
#include <stdlib.h>

#include <algorithm>

#include "kernels.cuh"

namespace sert {

namespace {

constexpr int kT = 8;                 // instances per CTA == warps per CTA
constexpr int kThreads = kT * 32;
constexpr int kD = 128;               // entity representation size served by this kernel; also the column block of the products
constexpr int kD4 = kD / 4;
constexpr int kLd = kD + 4;           // padded staging rows (entity-sized vectors)
constexpr int kMaxNW = 3;             // word rows of up to 3 x 128 floats
constexpr int kMaxRows = 16;          // 1 + k scores per instance handled by the butterfly
constexpr int kMaxWindow = 32;
constexpr int kSkipRow = (int)0x80000000;   // gradient destination of a row another rank updates (table shards)
constexpr int kOwnerShift = 28;             // instance shards: destination = row | owner << 28 (VsFusedArgs::n_owner)
constexpr int kRowMask = (1 << kOwnerShift) - 1;
constexpr int kRowsPerWarp = kD / kT; // matrix rows per warp in the K-split products
constexpr int kPartFloats = kT * kT * kD;   // 8 warps x 8 instances x 128 partial products (32 KB)

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_fma(float4 &acc, float s, const float4 &v) {
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}
__device__ __forceinline__ void f4_add(float4 &acc, const float4 &v) {
  acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
}
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Row `warp` of  vec (8 x K, shared, row stride ldv) . M[:, c0 .. c0 + 128) (K x ldm4*4, global, row-major), columns
// c0 + 4*lane .. + 3; only the first ncol4 float4 columns of the block exist (lanes beyond return zeros).  K is split
// over the 8 warps, rows_per_warp rows each (a multiple of 4; rows >= K are skipped).  Every thread of the CTA must
// call.  `part` (kPartFloats floats of shared memory) is scratch; the barrier at the start also publishes `vec`, the
// one at the end releases `part` for other uses.
// SQUARE: the 128 x 128 case of BASELINE configs[1] (K = 128, 16 rows per warp, full columns) with every bound a
// compile-time constant, as measured in round 1; the general form guards rows and columns.
template <bool SQUARE>
__device__ __forceinline__ float4 tile_matvec(const float *__restrict__ vec, int ldv, float *part,
                                              const float *__restrict__ M, int K, int rows_per_warp, int ldm4, int c0_4,
                                              int ncol4, int warp, int lane) {
  float4 acc[kT];
#pragma unroll
  for (int g = 0; g < kT; ++g) acc[g] = f4_zero();
  if (SQUARE) {
    const float4 *m = reinterpret_cast<const float4 *>(M) + (size_t)warp * kRowsPerWarp * kD4 + lane;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kRowsPerWarp; q += 4) {
      const float4 r0 = __ldg(m + (q + 0) * kD4);
      const float4 r1 = __ldg(m + (q + 1) * kD4);
      const float4 r2 = __ldg(m + (q + 2) * kD4);
      const float4 r3 = __ldg(m + (q + 3) * kD4);
#pragma unroll
      for (int g = 0; g < kT; ++g) {
        const float4 s = *reinterpret_cast<const float4 *>(vec + g * ldv + warp * kRowsPerWarp + q);   // broadcast
        f4_fma(acc[g], s.x, r0);
        f4_fma(acc[g], s.y, r1);
        f4_fma(acc[g], s.z, r2);
        f4_fma(acc[g], s.w, r3);
      }
    }
  } else {
    const bool col_ok = lane < ncol4;
    const float4 *m = reinterpret_cast<const float4 *>(M) + (size_t)warp * rows_per_warp * ldm4 + c0_4 + lane;
    __syncthreads();
    const int row0 = warp * rows_per_warp;
    for (int q = 0; q < rows_per_warp; q += 4) {
      if (row0 + q >= K) break;                           // warp-uniform
      float4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = (col_ok && row0 + q + u < K) ? __ldg(m + (size_t)(q + u) * ldm4) : f4_zero();
#pragma unroll
      for (int g = 0; g < kT; ++g) {
        const float4 s = *reinterpret_cast<const float4 *>(vec + g * ldv + row0 + q);   // broadcast
        f4_fma(acc[g], s.x, r[0]);
        f4_fma(acc[g], s.y, r[1]);
        f4_fma(acc[g], s.z, r[2]);
        f4_fma(acc[g], s.w, r[3]);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < kT; ++g) reinterpret_cast<float4 *>(part + (size_t)(warp * kT + g) * kD)[lane] = acc[g];
  __syncthreads();
  float4 out = f4_zero();
#pragma unroll
  for (int w = 0; w < kT; ++w) f4_add(out, reinterpret_cast<const float4 *>(part + (size_t)(w * kT + warp) * kD)[lane]);
  __syncthreads();
  return out;
}

// One step of the transposing butterfly: 2n values per lane -> n values per lane; lanes whose bit `o` is set keep
// the upper half.  After the steps o = 16, 8, 4, 2 and a plain exchange with o = 1, lanes 2j and 2j+1 hold the
// warp-wide sum of value j.
template <int n>
__device__ __forceinline__ void fold(float (&v)[kMaxRows], int lane, int o) {
  const bool upper = (lane & o) != 0;
#pragma unroll
  for (int i = 0; i < n; ++i) {
    const float keep = upper ? v[i + n] : v[i];
    const float send = upper ? v[i] : v[i + n];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
  }
}

// NW = chunks of 128 floats in a word row; SQ = the word rows are exactly 128 floats (every bound constant)
template <int NW, bool SQ>
__global__ void __launch_bounds__(kThreads, NW == 1 ? 4 : 3) vs_tile_kernel(VsFusedArgs a, const float *__restrict__ WpT) {
  extern __shared__ __align__(16) float scratch[];      // partial products, then [kT][K1][kD] entity rows of the tile
  constexpr int kLdH = NW * kD + 4;                     // padded rows of the word-sized staging buffer
  __shared__ __align__(16) float hbuf[kT * kLdH];       // h (word-sized rows), later da (entity-sized rows, stride kLd)
  const int dw = SQ ? kD : a.dw, dw4 = SQ ? kD4 : (a.dw >> 2);
  __shared__ int xs[kT * kMaxWindow];
  __shared__ double s_loss[kT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == gridDim.x - 1) {
    // The extra CTA: the previous step's loss (kernels.cuh: fin_acc).  loss = mean data loss + lambda/(2B) *
    // sum(theta^2) (sert/models.py:745-755,773-793), accumulators reset for the step after this one.
    if (a.fin_acc != nullptr && warp < 2) {
      const int s = threadIdx.x;        // kSumsqSlots == 64 slots, two warps
      double v = __ldcg(a.fin_acc + 1 + s);
      a.fin_acc[1 + s] = 0.0;
      v = warp_sum_d(v);
      if (lane == 0) s_loss[warp] = v;
      asm volatile("bar.sync 1, 64;" ::: "memory");
      if (s == 0) {
        const double ss = s_loss[0] + s_loss[1];
        *a.fin_loss = (float)((float)(__ldcg(a.fin_acc) * (double)a.fin_inv_B) + (float)((double)a.fin_reg_coeff * ss));
        a.fin_acc[0] = 0.0;
      }
    }
    return;
  }
  const int W = a.W, K1 = a.k + 1;
  const int i = a.i0 + blockIdx.x * kT + warp;  // warp == instance, everywhere but inside tile_matvec
  const bool ok = i < a.B;
  const size_t il = (size_t)(i - a.i0);         // row of the instance in h / da (instance shards start at i0)
  const float4 *R4 = reinterpret_cast<const float4 *>(a.R);
  const float4 *E4 = reinterpret_cast<const float4 *>(a.Eemb);
  float *my_ent = scratch + (size_t)warp * K1 * kD;

  // ---- A: the instance's indices, touched stamps, entity rows on their way into L2 -------------------------
  // xdst: gradient row id, -1 - slot for a hot row (kernels.cuh: hot_slot), kSkipRow for a row another rank updates
  // (table shards: its gradient is formed there); rdst: the same for the entity rows
  int xi = 0, ri = 0, xdst = kSkipRow, rdst = kSkipRow;
  if (ok && lane < W) {
    xi = __ldg(a.x + (size_t)i * W + lane);
    if (a.n_owner > 0) {
      // instance shards: the row's gradient is formed in the arena of the rank that updates it
      const int slot = a.hot_slot != nullptr ? (int)__ldg(a.hot_slot + xi) : -1;
      if (slot >= 0) {
        xdst = -1 - slot;
      } else {
        int o = 0;
        for (int q = 1; q < a.n_owner; ++q) o += xi >= a.r_bound[q] ? 1 : 0;
        a.flagR_peer[o][xi] = a.stamp;
        xdst = xi | (o << kOwnerShift);
      }
    } else if (a.own.word(xi)) {
      const int slot = a.hot_slot != nullptr ? (int)__ldg(a.hot_slot + xi) : -1;
      xdst = slot < 0 ? xi : -1 - slot;
      if (slot < 0) a.flagR[xi] = a.stamp;   // hot rows keep their kHotRowMark
    }
  }
  if (ok && lane < K1) {
    ri = lane == 0 ? __ldg(a.y + i) : __ldg(a.neg + (size_t)i * a.k + lane - 1);
    if (a.n_owner > 0) {
      int o = 0;
      for (int q = 1; q < a.n_owner; ++q) o += ri >= a.e_bound[q] ? 1 : 0;
      a.flagE_peer[o][ri] = a.stamp;
      rdst = ri | (o << kOwnerShift);
    } else if (a.own.entity(ri)) {
      a.flagE[ri] = a.stamp;
      rdst = ri;
    }
  }
  xs[warp * kMaxWindow + lane] = xdst;   // read back by this warp only
  for (int j = 0; j < K1; ++j) {
    const int r = __shfl_sync(0xffffffffu, ri, j);
    if (lane < 4) prefetch_l2(E4 + (size_t)r * kD4 + lane * 8);      // 4 lines of 128 B per row
  }

  // ---- B: gather + window mean (sert/models.py:180,226,1051) ------------------------------------------------
  float4 h[NW];
#pragma unroll
  for (int u = 0; u < NW; ++u) h[u] = f4_zero();
#pragma unroll 5
  for (int w = 0; w < W; ++w) {
    const int r = __shfl_sync(0xffffffffu, xi, w);
#pragma unroll
    for (int u = 0; u < NW; ++u)
      if (lane + 32 * u < dw4) f4_add(h[u], __ldg(R4 + (size_t)r * dw4 + lane + 32 * u));
  }
  const float den = (float)W;
#pragma unroll
  for (int u = 0; u < NW; ++u) {
    h[u].x = __fdiv_rn(h[u].x, den); h[u].y = __fdiv_rn(h[u].y, den);
    h[u].z = __fdiv_rn(h[u].z, den); h[u].w = __fdiv_rn(h[u].w, den);
    reinterpret_cast<float4 *>(hbuf + warp * kLdH)[lane + 32 * u] = h[u];          // chunks beyond dw hold zeros
    if (ok && lane + 32 * u < dw4) reinterpret_cast<float4 *>(a.h)[il * dw4 + lane + 32 * u] = h[u];
  }

  // ---- C: t = tanh(h . Wp + bp) (sert/models.py:1055-1061) --------------------------------------------------
  // K = dw rows of Wp (dw x 128) over the 8 warps, 4-row steps
  const int rpw_c = ((dw + kT - 1) / kT + 3) & ~3;
  float4 t = tile_matvec<SQ>(hbuf, kLdH, scratch, a.Wp, dw, rpw_c, kD4, 0, kD4, warp, lane);
  {
    const float4 b = __ldg(reinterpret_cast<const float4 *>(a.bp) + lane);
    t.x = tanhf(t.x + b.x); t.y = tanhf(t.y + b.y); t.z = tanhf(t.z + b.z); t.w = tanhf(t.w + b.w);
  }
  // the partial products are dead (barrier at the end of tile_matvec): this warp's entity rows take their place
  for (int j = 0; j < K1; ++j) {
    const int r = __shfl_sync(0xffffffffu, ri, j);
    cp_async_16(my_ent + j * kD + lane * 4, E4 + (size_t)r * kD4 + lane);
  }
  cp_async_wait_all();
  __syncwarp();

  // ---- D: negative-sampling loss of the instance, forward and backward (sert/models.py:893-902,1072-1098) ----
  float ell_lane = 0.f;
  float wi = 1.0f;
  {
    float4 u;
    u.x = clipf_(t.x, SERT_TANH_LO, SERT_TANH_HI); u.y = clipf_(t.y, SERT_TANH_LO, SERT_TANH_HI);
    u.z = clipf_(t.z, SERT_TANH_LO, SERT_TANH_HI); u.w = clipf_(t.w, SERT_TANH_LO, SERT_TANH_HI);
    float part[kMaxRows];
#pragma unroll
    for (int j = 0; j < kMaxRows; ++j) {
      part[j] = 0.f;
      if (j < K1) {
        const float4 e = reinterpret_cast<const float4 *>(my_ent + j * kD)[lane];
        part[j] = e.x * u.x + e.y * u.y + e.z * u.z + e.w * u.w;
      }
    }
    fold<8>(part, lane, 16);
    fold<4>(part, lane, 8);
    fold<2>(part, lane, 4);
    fold<1>(part, lane, 2);
    const float score = part[0] + __shfl_xor_sync(0xffffffffu, part[0], 1);
    const int j = lane >> 1;            // lanes 2j, 2j+1 hold score j
    if (ok && a.w != nullptr) wi = __ldg(a.w + i);
    const float coef_scale = ok ? wi * a.inv_B : 0.f;
    float coef = 0.f;
    if (j < K1) {
      const float sg = sigmoidf_(score);
      const float cl = clipf_(sg, SERT_CLIP_LO, SERT_CLIP_HI);
      const bool inside = (sg >= SERT_CLIP_LO) && (sg <= SERT_CLIP_HI);
      const float q = j == 0 ? cl : 1.0f - cl;          // probability the loss takes the log of
      const float ellj = -logf(q);
      const float mag = (coef_scale / q) * sg * (1.0f - sg);
      coef = inside ? (j == 0 ? -mag : mag) : 0.0f;
      if ((lane & 1) == 0) ell_lane = ellj;
    }
    float4 du = f4_zero();
    for (int jj = 0; jj < K1; ++jj) {
      const float c = __shfl_sync(0xffffffffu, coef, 2 * jj);
      const int r = __shfl_sync(0xffffffffu, rdst, jj);
      const float4 e = reinterpret_cast<const float4 *>(my_ent + jj * kD)[lane];
      f4_fma(du, c, e);
      if (ok && r != kSkipRow)
        red_add_f4(a.gE_peer[r >> kOwnerShift] + ((size_t)(r & kRowMask) * kD4 + lane) * 4,
                   make_float4(c * u.x, c * u.y, c * u.z, c * u.w));
    }
    float4 da;
    da.x = (t.x >= SERT_TANH_LO && t.x <= SERT_TANH_HI) ? du.x * (1.0f - t.x * t.x) : 0.f;
    da.y = (t.y >= SERT_TANH_LO && t.y <= SERT_TANH_HI) ? du.y * (1.0f - t.y * t.y) : 0.f;
    da.z = (t.z >= SERT_TANH_LO && t.z <= SERT_TANH_HI) ? du.z * (1.0f - t.z * t.z) : 0.f;
    da.w = (t.w >= SERT_TANH_LO && t.w <= SERT_TANH_HI) ? du.w * (1.0f - t.w * t.w) : 0.f;
    reinterpret_cast<float4 *>(hbuf + warp * kLd)[lane] = da;     // h is dead since the barriers of C
    if (ok) reinterpret_cast<float4 *>(a.da)[il * kD4 + lane] = da;
  }

  // ---- E: dh = da . Wp^T, scatter-add of dh / W into the word-gradient rows (the entity rows are dead behind the
  //         first barrier of tile_matvec) ------------------------------------------------------------------------
  {
    // WpT is (128 x dw): K = 128 rows, 16 per warp; the dw output columns in NW blocks of 128
    float4 dh[NW];
#pragma unroll
    for (int u = 0; u < NW; ++u) {
      const int ncol4 = min(kD4, dw4 - 32 * u);
      dh[u] = tile_matvec<SQ>(hbuf, kLd, scratch, WpT, kD, kRowsPerWarp, dw4, 32 * u, ncol4, warp, lane);
      dh[u].x = __fdiv_rn(dh[u].x, den); dh[u].y = __fdiv_rn(dh[u].y, den);
      dh[u].z = __fdiv_rn(dh[u].z, den); dh[u].w = __fdiv_rn(dh[u].w, den);
    }
    if (ok) {
      float *hot = a.hot_slot == nullptr
                       ? nullptr
                       : a.hot_acc + (size_t)(blockIdx.x & (a.hot_replicas - 1)) * kMaxHotRows * dw;
      for (int w = 0; w < W; ++w) {
        const int r = xs[warp * kMaxWindow + w];
        if (r == kSkipRow) continue;
        float *row = r >= 0 ? a.gR_peer[r >> kOwnerShift] + (size_t)(r & kRowMask) * dw : hot + (size_t)(-1 - r) * dw;
#pragma unroll
        for (int u = 0; u < NW; ++u)
          if (lane + 32 * u < dw4) red_add_f4(row + (lane + 32 * u) * 4, dh[u]);
      }
    }
  }

  // ---- tile loss: sum_i w_i * ell_i ------------------------------------------------------------------------
  const float ell = warp_sum(ell_lane);
  if (lane == 0) s_loss[warp] = ok ? (double)(wi * ell) : 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int wv = 0; wv < kT; ++wv) tot += s_loss[wv];
    if (tot != 0.0) atomicAdd(a.loss_acc, tot);
  }
}

}  // namespace

bool vs_tile_supported(int dw, int de, int W, int k) {
  return dw >= 4 && dw % 4 == 0 && dw <= kMaxNW * kD && de == kD && W >= 1 && W <= kMaxWindow && k + 1 <= kMaxRows;
}

// returns 0 = launched, 1 = shape not served by this kernel
int launch_vs_tile(const VsFusedArgs &a_in, const float *WpT, cudaStream_t st) {
  if (!vs_tile_supported(a_in.dw, a_in.de, a_in.W, a_in.k)) return 1;
  VsFusedArgs a = a_in;
  if (a.n_owner == 0) {                 // one destination: the kernel always goes through the owner table
    a.gE_peer[0] = a.gE; a.gR_peer[0] = a.gR;
  } else {
    SERT_REQUIRE(a.n_owner <= kMaxOwners && a.i0 >= 0 && a.i0 <= a.B, "bad instance shard");
  }
  const size_t smem = std::max((size_t)kT * (a.k + 1) * kD, (size_t)kPartFloats) * sizeof(float);
  static std::atomic<uint64_t> configured{0};
  if (first_use_on_device(configured)) {
    const int most = (int)(kT * kMaxRows * kD * sizeof(float));
    SERT_CUDA(cudaFuncSetAttribute(vs_tile_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
    SERT_CUDA(cudaFuncSetAttribute(vs_tile_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
    SERT_CUDA(cudaFuncSetAttribute(vs_tile_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
    SERT_CUDA(cudaFuncSetAttribute(vs_tile_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
  }
  static_assert(kSumsqSlots == 64, "the finalising CTA reads the slots with two warps");
  const int grid = cdiv(a.B - a.i0, kT) + 1;             // + 1: the CTA that finalises the previous loss
  const int nw = (a.dw + kD - 1) / kD;
  if (a.dw == kD) vs_tile_kernel<1, true><<<grid, kThreads, smem, st>>>(a, WpT);
  else if (nw == 1) vs_tile_kernel<1, false><<<grid, kThreads, smem, st>>>(a, WpT);
  else if (nw == 2) vs_tile_kernel<2, false><<<grid, kThreads, smem, st>>>(a, WpT);
  else vs_tile_kernel<3, false><<<grid, kThreads, smem, st>>>(a, WpT);
  SERT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sert
