// Fused forward + backward of the vector-space model for one tile of instances per CTA
// (sert/models.py:1044-1098): gather + window mean, tanh projection, negative-sampling loss and its
// gradients, back-projection and the scatter-add into the word-gradient rows.  The projection matrix
// (dw x de fp32, 64 KB at d=128) lives in shared memory for the whole CTA and the (TB x d) intermediates
// h, t, da, dh never leave the SM; only h and da are also written to HBM for the dW = h^T.da GEMM that
// follows.  Replaces five launches (gather_pool, gemm+tanh, vs_nce, gemm dh, scatter_rows) of the
// unfused path, which stays as the general-shape fallback (launch_vs_fused returns 1 when a shape does
// not fit) and as the eval / parity-hook path.
#include "kernels.cuh"

namespace sert {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct Smem {
  float *wp;    // [dw][de+1]   (odd row stride: conflict-free for both W and W^T access)
  float *bp;    // [de]
  float *hs;    // [TB][dw]     h, later dh
  float *ts;    // [TB][de]     tanh output t
  float *das;   // [TB][de]     d loss / d pre-activation
};

__host__ __device__ inline size_t fused_smem_floats(int TB, int dw, int de) {
  return (size_t)dw * (de + 1) + de + (size_t)TB * dw + 2 * (size_t)TB * de + 16;
}

template <int TB, int CW, int CE>
__global__ void __launch_bounds__(kThreads, TB <= 16 ? 2 : 1) vs_fused_kernel(VsFusedArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int dw = a.dw, de = a.de, W = a.W;
  const int ldw = de + 1;
  Smem s;
  s.hs = smem;                                   // 16-byte aligned rows (dw % 4 == 0)
  s.ts = s.hs + (size_t)TB * dw;
  s.das = s.ts + (size_t)TB * de;
  s.bp = s.das + (size_t)TB * de;
  s.wp = s.bp + de;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0 = blockIdx.x * TB;
  const int dw4 = dw >> 2, de4 = de >> 2;
  __shared__ double s_loss[kWarps];

  // ---- phase 0: projection matrix and bias into shared memory --------------------------------------
  for (int e4 = tid; e4 < dw * de4; e4 += kThreads) {
    const int r = e4 / de4, c = (e4 - r * de4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4 *>(a.Wp) + e4);
    float *dst = s.wp + r * ldw + c;
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  for (int c = tid; c < de; c += kThreads) s.bp[c] = __ldg(a.bp + c);

  // ---- phase 1: h = mean_w R[x]  (warp per instance) --------------------------------------------------
  const float4 *R4 = reinterpret_cast<const float4 *>(a.R);
  for (int li = warp; li < TB; li += kWarps) {
    const int i = i0 + li;
    const bool inst_ok = i < a.B;
    const int32_t *xi = a.x + (size_t)(inst_ok ? i : 0) * W;
    for (int c0 = 0; c0 < dw4; c0 += 32) {
      const int c = c0 + lane;
      const bool active = c < dw4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int w0 = 0; w0 < W; w0 += 32) {
        const int nw = min(32, W - w0);
        const int idx = (lane < nw && inst_ok) ? __ldg(xi + w0 + lane) : 0;
#pragma unroll 5
        for (int w = 0; w < nw; ++w) {
          const int r = __shfl_sync(0xffffffffu, idx, w);
          if (active && inst_ok) {
            const float4 v = __ldg(R4 + (size_t)r * dw4 + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
        }
      }
      if (active) {
        const float den = (float)W;
        acc.x /= den; acc.y /= den; acc.z /= den; acc.w /= den;
        reinterpret_cast<float4 *>(s.hs + (size_t)li * dw)[c] = acc;
        if (inst_ok) reinterpret_cast<float4 *>(a.h)[(size_t)i * dw4 + c] = acc;
      }
    }
  }
  __syncthreads();

  // ---- phase 2: t = tanh(h . Wp + bp): thread -> rows {ty*RPT..}, cols {lane + 32*cc} -------------------
  constexpr int RPT = TB / kWarps;              // rows per thread (one row group per warp)
  {
    float acc[RPT][CE];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
      for (int cc = 0; cc < CE; ++cc) acc[r][cc] = 0.f;
    for (int k = 0; k < dw; ++k) {
      float hv[RPT], wv[CE];
#pragma unroll
      for (int r = 0; r < RPT; ++r) hv[r] = s.hs[(size_t)(warp * RPT + r) * dw + k];      // warp broadcast
#pragma unroll
      for (int cc = 0; cc < CE; ++cc) {
        const int c = lane + 32 * cc;
        wv[cc] = c < de ? s.wp[k * ldw + c] : 0.f;
      }
#pragma unroll
      for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int cc = 0; cc < CE; ++cc) acc[r][cc] = fmaf(hv[r], wv[cc], acc[r][cc]);
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
      for (int cc = 0; cc < CE; ++cc) {
        const int c = lane + 32 * cc;
        if (c < de) s.ts[(size_t)(warp * RPT + r) * de + c] = tanhf(acc[r][cc] + s.bp[c]);
      }
  }
  __syncthreads();

  // ---- phase 3: negative-sampling loss, forward and backward (warp per instance) ------------------------
  constexpr int MAXC = (CE + 3) / 4;             // float4 chunks per lane over de
  double my_loss = 0.0;
  const float4 *E4 = reinterpret_cast<const float4 *>(a.Eemb);
  for (int li = warp; li < TB; li += kWarps) {
    const int i = i0 + li;
    if (i >= a.B) {                              // warp-uniform
      for (int c = lane; c < de; c += 32) s.das[(size_t)li * de + c] = 0.f;
      continue;
    }
    float4 tt[MAXC], u[MAXC], du[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + 32 * c;
      tt[c] = ch < de4 ? reinterpret_cast<const float4 *>(s.ts + (size_t)li * de)[ch] : make_float4(0.f, 0.f, 0.f, 0.f);
      u[c].x = clipf_(tt[c].x, SERT_TANH_LO, SERT_TANH_HI);
      u[c].y = clipf_(tt[c].y, SERT_TANH_LO, SERT_TANH_HI);
      u[c].z = clipf_(tt[c].z, SERT_TANH_LO, SERT_TANH_HI);
      u[c].w = clipf_(tt[c].w, SERT_TANH_LO, SERT_TANH_HI);
      du[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float wi = a.w ? __ldg(a.w + i) : 1.0f;
    const float coef_scale = wi * a.inv_B;
    const int yi = __ldg(a.y + i);
    const int32_t *negi = a.neg + (size_t)i * a.k;
    float ell = 0.f;
    constexpr int GROUP = MAXC <= 1 ? 4 : 2;
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
          e[g][c] = (rows[g] >= 0 && ch < de4) ? __ldg(E4 + (size_t)rows[g] * de4 + ch)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int g = 0; g < GROUP; ++g) {
        float sdot = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          sdot += e[g][c].x * u[c].x + e[g][c].y * u[c].y + e[g][c].z * u[c].z + e[g][c].w * u[c].w;
        dots[g] = sdot;
      }
#pragma unroll
      for (int g = 0; g < GROUP; ++g) dots[g] = warp_sum(dots[g]);
#pragma unroll
      for (int g = 0; g < GROUP; ++g) {
        if (rows[g] < 0) continue;
        const int j = j0 + g;
        const float sg = sigmoidf_(dots[g]);
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
        if (lane == 0) a.flagE[rows[g]] = a.stamp;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          const int ch = lane + 32 * c;
          du[c].x += coef * e[g][c].x; du[c].y += coef * e[g][c].y;
          du[c].z += coef * e[g][c].z; du[c].w += coef * e[g][c].w;
          if (ch < de4)
            red_add_f4(a.gE + ((size_t)rows[g] * de4 + ch) * 4,
                       make_float4(coef * u[c].x, coef * u[c].y, coef * u[c].z, coef * u[c].w));
        }
      }
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int ch = lane + 32 * c;
      if (ch >= de4) continue;
      float4 o;
      o.x = (tt[c].x >= SERT_TANH_LO && tt[c].x <= SERT_TANH_HI) ? du[c].x * (1.0f - tt[c].x * tt[c].x) : 0.f;
      o.y = (tt[c].y >= SERT_TANH_LO && tt[c].y <= SERT_TANH_HI) ? du[c].y * (1.0f - tt[c].y * tt[c].y) : 0.f;
      o.z = (tt[c].z >= SERT_TANH_LO && tt[c].z <= SERT_TANH_HI) ? du[c].z * (1.0f - tt[c].z * tt[c].z) : 0.f;
      o.w = (tt[c].w >= SERT_TANH_LO && tt[c].w <= SERT_TANH_HI) ? du[c].w * (1.0f - tt[c].w * tt[c].w) : 0.f;
      reinterpret_cast<float4 *>(s.das + (size_t)li * de)[ch] = o;
      reinterpret_cast<float4 *>(a.da)[(size_t)i * de4 + ch] = o;
    }
    my_loss += (double)(wi * ell);
  }
  __syncthreads();

  // ---- phase 4: dh = da . Wp^T into the h tile, then scatter-add dh / W into the word-gradient rows -------
  {
    float acc[RPT][CW];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
      for (int cc = 0; cc < CW; ++cc) acc[r][cc] = 0.f;
    for (int n = 0; n < de; ++n) {
      float dv[RPT], wv[CW];
#pragma unroll
      for (int r = 0; r < RPT; ++r) dv[r] = s.das[(size_t)(warp * RPT + r) * de + n];
#pragma unroll
      for (int cc = 0; cc < CW; ++cc) {
        const int kd = lane + 32 * cc;
        wv[cc] = kd < dw ? s.wp[kd * ldw + n] : 0.f;          // row stride de+1: conflict-free
      }
#pragma unroll
      for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int cc = 0; cc < CW; ++cc) acc[r][cc] = fmaf(dv[r], wv[cc], acc[r][cc]);
    }
    __syncthreads();                                          // every warp is done reading nothing from hs, but keep phases apart
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
      for (int cc = 0; cc < CW; ++cc) {
        const int kd = lane + 32 * cc;
        if (kd < dw) s.hs[(size_t)(warp * RPT + r) * dw + kd] = acc[r][cc];
      }
  }
  __syncthreads();
  for (int li = warp; li < TB; li += kWarps) {
    const int i = i0 + li;
    if (i >= a.B) continue;
    const int32_t *xi = a.x + (size_t)i * W;
    for (int w0 = 0; w0 < W; w0 += 32) {
      const int nw = min(32, W - w0);
      const int idx = (lane < nw) ? __ldg(xi + w0 + lane) : 0;
      if (lane < nw) a.flagR[idx] = a.stamp;
      for (int c0 = 0; c0 < dw4; c0 += 32) {
        const int c = c0 + lane;
        const bool active = c < dw4;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
          g = reinterpret_cast<const float4 *>(s.hs + (size_t)li * dw)[c];
          const float den = (float)W;
          g.x /= den; g.y /= den; g.z /= den; g.w /= den;
        }
        for (int w = 0; w < nw; ++w) {
          const int r = __shfl_sync(0xffffffffu, idx, w);
          if (active) red_add_f4(a.gR + ((size_t)r * dw4 + c) * 4, g);
        }
      }
    }
  }

  if (lane == 0) s_loss[warp] = my_loss;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int wv = 0; wv < kWarps; ++wv) tot += s_loss[wv];
    atomicAdd(a.loss_acc, tot);
  }
}

template <int TB, int CW, int CE>
int launch_t(const VsFusedArgs &a, size_t smem_bytes, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    SERT_CUDA(cudaFuncSetAttribute(vs_fused_kernel<TB, CW, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   200 * 1024));
    configured = true;
  }
  vs_fused_kernel<TB, CW, CE><<<cdiv(a.B, TB), kThreads, smem_bytes, st>>>(a);
  SERT_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// returns 0 = launched, 1 = shape not supported by the fused kernel (caller uses the unfused path), -1 = error
int launch_vs_fused(const VsFusedArgs &a, cudaStream_t st) {
  if (a.B == 0) return 0;
  if (a.dw % 4 != 0 || a.de % 4 != 0) return 1;
  const int cw = (a.dw + 31) / 32, ce = (a.de + 31) / 32;
  // 16-instance tiles, two CTAs per SM (16 warps in flight hide the gather / LDS latencies) when the tile
  // fits twice in shared memory, else 32-instance tiles with one CTA per SM
  const size_t bytes16 = fused_smem_floats(16, a.dw, a.de) * sizeof(float);
  const size_t bytes32 = fused_smem_floats(32, a.dw, a.de) * sizeof(float);
  const bool two = bytes16 <= 110 * 1024;
  if (!two && bytes32 > 200 * 1024) return 1;
#define SERT_FUSED_CASE(CWv, CEv)                                              \
  if (cw == CWv && ce == CEv)                                                  \
    return two ? launch_t<16, CWv, CEv>(a, bytes16, st) : launch_t<32, CWv, CEv>(a, bytes32, st);
  SERT_FUSED_CASE(4, 4)     // d = 128 (BASELINE configs[1])
  SERT_FUSED_CASE(2, 2)     // d = 64
  SERT_FUSED_CASE(1, 1)     // d = 32
  SERT_FUSED_CASE(4, 2)
  SERT_FUSED_CASE(2, 4)
  SERT_FUSED_CASE(8, 4)     // dw = 256, de = 128
#undef SERT_FUSED_CASE
  return 1;
}

}  // namespace sert
