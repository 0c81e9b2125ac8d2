// Vector-space training step, forward + backward, ONE WARP PER PAIR OF INSTANCES (sert/models.py:1044-1098):
// gather + window mean, tanh projection, negative-sampling loss and its gradients, back-projection and the
// scatter-add into the word-gradient rows -- all in registers, no block-level synchronisation.
//
// Why this shape: every stage is a chain of dependent HBM / L2 round trips (row gathers, entity rows, atomics)
// with a little arithmetic between them.  The projection matrix W (dw x de) and its transpose are read through
// L1 with fully coalesced 512-byte row requests (128 KB for both at d=128), each row request is shared by the
// warp's two instances, and the only shared memory is a 1 KB per-warp staging buffer that turns the h / da
// vectors into broadcast LDS.128 operands of the two matrix-vector products; representation sizes up to 384
// are supported (product-search.sh uses 300 / 128).  Measured at BASELINE configs[1] (ncu, profiles/): 45-50 us,
// issue-bound at ~14 warps per SM (21 M warp instructions, 42 % issue-slot utilisation, L1 hit rate 70 %, L2
// atomic units 7 % busy).  A tile-per-CTA variant with W resident in shared memory (git history: vs_fused.cu)
// measured the same 40-48 us and could not hold a 300 x 128 matrix, which is why this one stayed.
#include <stdlib.h>

#include "kernels.cuh"

namespace sert {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_fma(float4 &acc, float s, const float4 &v) {
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}

// out[g][c] (float4 chunk lane+32c of a length-N vector) = sum_k vec_g[k] * M[k][chunk], M row-major (K, N).
// vec_g comes from the warp's shared staging buffer (broadcast LDS.128), M rows through L1 (coalesced).
template <int kInst, int CN>
__device__ __forceinline__ void warp_matvec(const float *__restrict__ stage, int stage_ld, int K,
                                            const float4 *__restrict__ M4, int n4, int lane,
                                            float4 (&out)[kInst][CN]) {
#pragma unroll
  for (int g = 0; g < kInst; ++g)
#pragma unroll
    for (int c = 0; c < CN; ++c) out[g][c] = f4_zero();
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    float4 v[kInst];
#pragma unroll
    for (int g = 0; g < kInst; ++g) v[g] = *reinterpret_cast<const float4 *>(stage + g * stage_ld + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float4 row[CN];
#pragma unroll
      for (int c = 0; c < CN; ++c) {
        const int ch = lane + 32 * c;
        row[c] = ch < n4 ? __ldg(M4 + (size_t)(k + kk) * n4 + ch) : f4_zero();
      }
#pragma unroll
      for (int g = 0; g < kInst; ++g) {
        const float s = kk == 0 ? v[g].x : kk == 1 ? v[g].y : kk == 2 ? v[g].z : v[g].w;
#pragma unroll
        for (int c = 0; c < CN; ++c) f4_fma(out[g][c], s, row[c]);
      }
    }
  }
}

template <int kInst, int CW, int CE>
__global__ void __launch_bounds__(kThreads, kInst == 1 ? 4 : 3) vs_warp_kernel(VsFusedArgs a, const float *__restrict__ WpT) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dw = a.dw, de = a.de, W = a.W, dw4 = dw >> 2, de4 = de >> 2;
  const int stage_ld = max(dw, de);
  float *stage = smem + (size_t)warp * kInst * stage_ld;
  const int i_base = (blockIdx.x * kWarps + warp) * kInst;
  __shared__ double s_loss[kWarps];
  double my_loss = 0.0;
  const float4 *R4 = reinterpret_cast<const float4 *>(a.R);
  const float4 *E4 = reinterpret_cast<const float4 *>(a.Eemb);

  if (i_base < a.B) {
    bool ok[kInst];
#pragma unroll
    for (int g = 0; g < kInst; ++g) ok[g] = i_base + g < a.B;

    // ---- gather + window mean: h[g] in registers, staged for the first matvec ---------------------------
    float4 h[kInst][CW];
#pragma unroll
    for (int g = 0; g < kInst; ++g) {
#pragma unroll
      for (int c = 0; c < CW; ++c) h[g][c] = f4_zero();
      const int32_t *xi = a.x + (size_t)(ok[g] ? i_base + g : i_base) * W;
      for (int w0 = 0; w0 < W; w0 += 32) {
        const int nw = min(32, W - w0);
        const int idx = lane < nw ? __ldg(xi + w0 + lane) : 0;
#pragma unroll 5
        for (int w = 0; w < nw; ++w) {
          const int r = __shfl_sync(0xffffffffu, idx, w);
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            const int ch = lane + 32 * c;
            if (ch < dw4) {
              const float4 v = __ldg(R4 + (size_t)r * dw4 + ch);
              h[g][c].x += v.x; h[g][c].y += v.y; h[g][c].z += v.z; h[g][c].w += v.w;
            }
          }
        }
      }
      const float den = (float)W;
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        const int ch = lane + 32 * c;
        if (ch < dw4) {
          h[g][c].x /= den; h[g][c].y /= den; h[g][c].z /= den; h[g][c].w /= den;
          reinterpret_cast<float4 *>(stage + g * stage_ld)[ch] = h[g][c];
          if (ok[g]) reinterpret_cast<float4 *>(a.h)[(size_t)(i_base + g) * dw4 + ch] = h[g][c];
        }
      }
    }
    __syncwarp();

    // ---- t = tanh(h . Wp + bp) -----------------------------------------------------------------------------
    float4 t[kInst][CE];
    warp_matvec<kInst, CE>(stage, stage_ld, dw, reinterpret_cast<const float4 *>(a.Wp), de4, lane, t);
#pragma unroll
    for (int c = 0; c < CE; ++c) {
      const int ch = lane + 32 * c;
      const float4 b = ch < de4 ? __ldg(reinterpret_cast<const float4 *>(a.bp) + ch) : f4_zero();
#pragma unroll
      for (int g = 0; g < kInst; ++g) {
        t[g][c].x = tanhf(t[g][c].x + b.x); t[g][c].y = tanhf(t[g][c].y + b.y);
        t[g][c].z = tanhf(t[g][c].z + b.z); t[g][c].w = tanhf(t[g][c].w + b.w);
      }
    }
    __syncwarp();

    // ---- negative-sampling loss, forward and backward, one instance at a time ------------------------------
#pragma unroll
    for (int g = 0; g < kInst; ++g) {
      float4 da[CE];
      if (ok[g]) {
        const int i = i_base + g;
        float4 u[CE], du[CE];
#pragma unroll
        for (int c = 0; c < CE; ++c) {
          u[c].x = clipf_(t[g][c].x, SERT_TANH_LO, SERT_TANH_HI);
          u[c].y = clipf_(t[g][c].y, SERT_TANH_LO, SERT_TANH_HI);
          u[c].z = clipf_(t[g][c].z, SERT_TANH_LO, SERT_TANH_HI);
          u[c].w = clipf_(t[g][c].w, SERT_TANH_LO, SERT_TANH_HI);
          du[c] = f4_zero();
        }
        const float wi = a.w ? __ldg(a.w + i) : 1.0f;
        const float coef_scale = wi * a.inv_B;
        const int yi = __ldg(a.y + i);
        const int32_t *negi = a.neg + (size_t)i * a.k;
        float ell = 0.f;
        constexpr int GROUP = CE <= 1 ? 4 : 2;
        for (int j0 = 0; j0 <= a.k; j0 += GROUP) {
          int rows[GROUP];
          float4 e[GROUP][CE];
          float dots[GROUP];
#pragma unroll
          for (int q = 0; q < GROUP; ++q) {
            const int j = j0 + q;
            rows[q] = (j > a.k) ? -1 : (j == 0 ? yi : __ldg(negi + j - 1));
#pragma unroll
            for (int c = 0; c < CE; ++c) {
              const int ch = lane + 32 * c;
              e[q][c] = (rows[q] >= 0 && ch < de4) ? __ldg(E4 + (size_t)rows[q] * de4 + ch) : f4_zero();
            }
          }
#pragma unroll
          for (int q = 0; q < GROUP; ++q) {
            float sdot = 0.f;
#pragma unroll
            for (int c = 0; c < CE; ++c)
              sdot += e[q][c].x * u[c].x + e[q][c].y * u[c].y + e[q][c].z * u[c].z + e[q][c].w * u[c].w;
            dots[q] = sdot;
          }
#pragma unroll
          for (int q = 0; q < GROUP; ++q) dots[q] = warp_sum(dots[q]);
#pragma unroll
          for (int q = 0; q < GROUP; ++q) {
            if (rows[q] < 0) continue;
            const int j = j0 + q;
            const float sg = sigmoidf_(dots[q]);
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
            const bool mine = a.own.entity(rows[q]);     // table shards: another rank forms this row's gradient
            if (lane == 0 && mine) a.flagE[rows[q]] = a.stamp;
#pragma unroll
            for (int c = 0; c < CE; ++c) {
              const int ch = lane + 32 * c;
              f4_fma(du[c], coef, e[q][c]);
              if (ch < de4 && mine)
                red_add_f4(a.gE + ((size_t)rows[q] * de4 + ch) * 4,
                           make_float4(coef * u[c].x, coef * u[c].y, coef * u[c].z, coef * u[c].w));
            }
          }
        }
#pragma unroll
        for (int c = 0; c < CE; ++c) {
          const float4 tt = t[g][c];
          da[c].x = (tt.x >= SERT_TANH_LO && tt.x <= SERT_TANH_HI) ? du[c].x * (1.0f - tt.x * tt.x) : 0.f;
          da[c].y = (tt.y >= SERT_TANH_LO && tt.y <= SERT_TANH_HI) ? du[c].y * (1.0f - tt.y * tt.y) : 0.f;
          da[c].z = (tt.z >= SERT_TANH_LO && tt.z <= SERT_TANH_HI) ? du[c].z * (1.0f - tt.z * tt.z) : 0.f;
          da[c].w = (tt.w >= SERT_TANH_LO && tt.w <= SERT_TANH_HI) ? du[c].w * (1.0f - tt.w * tt.w) : 0.f;
        }
        my_loss += (double)(wi * ell);
      } else {
#pragma unroll
        for (int c = 0; c < CE; ++c) da[c] = f4_zero();
      }
#pragma unroll
      for (int c = 0; c < CE; ++c) {
        const int ch = lane + 32 * c;
        if (ch < de4) {
          reinterpret_cast<float4 *>(stage + g * stage_ld)[ch] = da[c];
          if (ok[g]) reinterpret_cast<float4 *>(a.da)[(size_t)(i_base + g) * de4 + ch] = da[c];
        }
      }
    }
    __syncwarp();

    // ---- dh = da . Wp^T (rows of the transposed copy), then scatter-add dh / W into the word-gradient rows ---
    float4 dh[kInst][CW];
    warp_matvec<kInst, CW>(stage, stage_ld, de, reinterpret_cast<const float4 *>(WpT), dw4, lane, dh);
#pragma unroll
    for (int g = 0; g < kInst; ++g) {
      if (!ok[g]) continue;
      const float den = (float)W;
#pragma unroll
      for (int c = 0; c < CW; ++c) { dh[g][c].x /= den; dh[g][c].y /= den; dh[g][c].z /= den; dh[g][c].w /= den; }
      const int32_t *xi = a.x + (size_t)(i_base + g) * W;
      for (int w0 = 0; w0 < W; w0 += 32) {
        const int nw = min(32, W - w0);
        const int idx = lane < nw ? __ldg(xi + w0 + lane) : 0;
        if (lane < nw && a.own.word(idx)) a.flagR[idx] = a.stamp;
        for (int w = 0; w < nw; ++w) {
          const int r = __shfl_sync(0xffffffffu, idx, w);
          if (!a.own.word(r)) continue;
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            const int ch = lane + 32 * c;
            if (ch < dw4) red_add_f4(a.gR + ((size_t)r * dw4 + ch) * 4, dh[g][c]);
          }
        }
      }
    }
  }

  if (lane == 0) s_loss[warp] = my_loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int wv = 0; wv < kWarps; ++wv) tot += s_loss[wv];
    if (tot != 0.0) atomicAdd(a.loss_acc, tot);
  }
}

// WpT (de, dw) <- Wp (dw, de): 64 KB at d = 128.  Only launched when the copy is stale (first step, parameters set
// from the host): in steady state the dense update of the previous step writes it (opt_kernels.cu, phase 4).
__global__ void transpose_small_kernel(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dst[(size_t)c * rows + r] = tile[threadIdx.x][j];
  }
}

template <int kInst, int CW, int CE>
int launch_i(const VsFusedArgs &a, const float *WpT, cudaStream_t st) {
  const size_t smem = (size_t)kWarps * kInst * std::max(a.dw, a.de) * sizeof(float);
  const int warps = (a.B + kInst - 1) / kInst;
  vs_warp_kernel<kInst, CW, CE><<<cdiv(warps, kWarps), kThreads, smem, st>>>(a, WpT);
  SERT_LAUNCH_CHECK();
  return 0;
}

// instances per warp: 1 fills the machine with B warps (shortest dependent chain per warp); 2 halves the L1
// traffic of the projection-matrix rows.  SERT_VS_INST overrides for experiments.
template <int CW, int CE>
int launch_t(const VsFusedArgs &a, const float *WpT, cudaStream_t st) {
  static int inst = 0;
  if (inst == 0) {
    const char *e = getenv("SERT_VS_INST");
    inst = (e && e[0] == '1') ? 1 : 2;     // measured equal at BASELINE configs[1] (0.1443 vs 0.1446 ms/step)
  }
  return inst == 2 ? launch_i<2, CW, CE>(a, WpT, st) : launch_i<1, CW, CE>(a, WpT, st);
}

}  // namespace

// returns 0 = launched, 1 = shape not supported (caller uses the per-stage kernels), -1 = error
// variant: 1 = tile kernel (csrc/vs_tile.cu) where the shape fits, else the warp kernel; 2 = warp kernel only
int launch_vs_fused(const VsFusedArgs &a, float *WpT_scratch, bool refresh_WpT, int variant, cudaStream_t st) {
  if (a.B == 0) return 0;
  if (a.dw % 4 != 0 || a.de % 4 != 0 || a.dw > 384 || a.de > 384 || WpT_scratch == nullptr) return 1;
  if (refresh_WpT) {
    transpose_small_kernel<<<dim3(cdiv(a.de, 32), cdiv(a.dw, 32)), dim3(32, 8), 0, st>>>(a.Wp, WpT_scratch, a.dw, a.de);
    SERT_LAUNCH_CHECK();
  }
  if (variant != 2) {
    const int rc = launch_vs_tile(a, WpT_scratch, st);
    if (rc <= 0) return rc;
  }
  const int cw = (a.dw / 4 + 31) / 32, ce = (a.de / 4 + 31) / 32;
#define SERT_WARP_CASE(CWv, CEv) \
  if (cw == CWv && ce == CEv) return launch_t<CWv, CEv>(a, WpT_scratch, st);
  SERT_WARP_CASE(1, 1)   // dw, de <= 128   (BASELINE configs[1]: 128 / 128)
  SERT_WARP_CASE(2, 1)   // dw <= 256
  SERT_WARP_CASE(3, 1)   // dw <= 384       (product-search.sh: 300 / 128)
  SERT_WARP_CASE(1, 2)
  SERT_WARP_CASE(2, 2)
  SERT_WARP_CASE(3, 2)
  SERT_WARP_CASE(3, 3)
#undef SERT_WARP_CASE
  return 1;
}

}  // namespace sert
