// Dense Adam / Adadelta with L2, fused with the sparse-gradient pickup and the loss finalisation.
//
// Reference semantics: the L2 term of sert/models.py:764-795 puts a
// gradient lambda/B * theta on EVERY row of every table EVERY step, and lasagne.updates.adam /
// adadelta (sert/models.py:820,922) keep two state arrays per parameter, so one training step
// streams 6 float32 arrays over all parameters: 24 B/param, the dominant HBM traffic of the step.
// This kernel is that stream and nothing more: theta, s1, s2 are read and written exactly once with
// 16-byte streaming accesses; the data gradient is only fetched for rows whose `flags` entry carries
// this step's stamp (rows the scatter kernels touched), and is zeroed in the same pass.  The same
// pass accumulates sum(theta^2) of the pre-update values (the regulariser's contribution to the
// reported loss, f64 accumulation like Theano's CPU Sum), and the last block to finish writes the
// step's scalar loss.
#include <algorithm>

#include "kernels.cuh"

namespace sert {

constexpr int kUnroll = 2;   // independent 16-byte chunks per thread per iteration (3 streams each)

template <bool ADAM>
__global__ void __launch_bounds__(256, 4) dense_update_kernel(OptimArgs a) {
  const long long total4 = a.total >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float4 *__restrict__ th4 = reinterpret_cast<float4 *>(a.theta);
  float4 *__restrict__ s14 = reinterpret_cast<float4 *>(a.s1);
  float4 *__restrict__ s24 = reinterpret_cast<float4 *>(a.s2);
  float4 *__restrict__ g4 = reinterpret_cast<float4 *>(a.grad);
  const float one_m_c1 = 1.0f - a.c1;
  const float one_m_c2 = 1.0f - a.c2;
  float sumsq = 0.f;

  for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < total4; base += stride * kUnroll) {
    long long idx[kUnroll];
    bool live[kUnroll], touched[kUnroll];
    float l2[kUnroll];
    float4 p[kUnroll], x1[kUnroll], x2[kUnroll], g[kUnroll];
    // ---- issue every load of the iteration before the first use ----
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      idx[u] = base + (long long)u * stride;
      const bool in = idx[u] < total4;
      const long long e = idx[u] << 2;
      int s = 0;
#pragma unroll
      for (int q = 1; q < kMaxSegments; ++q) s += (q < a.num_segments && e >= a.seg[q].offset) ? 1 : 0;
      const ParamSegment &sg = a.seg[s];
      live[u] = in && (e - sg.offset) < sg.count;   // false inside inter-segment padding / past the end
      touched[u] = live[u];
      l2[u] = sg.regularised ? a.l2_scale : -1.0f;  // negative: not regularised
      if (live[u] && sg.flags != nullptr) {
        const unsigned int row = (unsigned int)(e - sg.offset) / (unsigned int)sg.row_len;
        touched[u] = (__ldg(sg.flags + row) == a.stamp);
      }
      if (live[u]) {
        p[u] = ld_stream_f4(th4 + idx[u]);
        x1[u] = ld_stream_f4(s14 + idx[u]);
        x2[u] = ld_stream_f4(s24 + idx[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (touched[u]) {
        g[u] = g4[idx[u]];
        g4[idx[u]] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (!live[u]) continue;
      const bool reg = l2[u] >= 0.0f;
      const float l2s = reg ? l2[u] : 0.0f;
      if (reg) sumsq += p[u].x * p[u].x + p[u].y * p[u].y + p[u].z * p[u].z + p[u].w * p[u].w;
      float pv[4] = {p[u].x, p[u].y, p[u].z, p[u].w};
      float v1[4] = {x1[u].x, x1[u].y, x1[u].z, x1[u].w};
      float v2[4] = {x2[u].x, x2[u].y, x2[u].z, x2[u].w};
      const float gv[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gj = gv[j] + l2s * pv[j];
        if (ADAM) {
          // m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; theta <- theta - a_t m / (sqrt(v) + eps)
          const float m = a.c1 * v1[j] + one_m_c1 * gj;
          const float v = a.c2 * v2[j] + one_m_c2 * gj * gj;
          pv[j] = pv[j] - a.c0 * m / (sqrtf(v) + a.c3);
          v1[j] = m; v2[j] = v;
        } else {
          // accu <- rho accu + (1-rho) g^2 ; upd = g sqrt(delta+eps)/sqrt(accu+eps) ;
          // theta <- theta - lr upd ; delta <- rho delta + (1-rho) upd^2
          const float accu = a.c1 * v1[j] + one_m_c1 * gj * gj;
          const float upd = gj * sqrtf(v2[j] + a.c3) / sqrtf(accu + a.c3);
          pv[j] = pv[j] - a.c0 * upd;
          v1[j] = accu; v2[j] = a.c1 * v2[j] + one_m_c1 * upd * upd;
        }
      }
      st_stream_f4(th4 + idx[u], make_float4(pv[0], pv[1], pv[2], pv[3]));
      st_stream_f4(s14 + idx[u], make_float4(v1[0], v1[1], v1[2], v1[3]));
      st_stream_f4(s24 + idx[u], make_float4(v2[0], v2[1], v2[2], v2[3]));
    }
  }

  // ---- block reduction of sum(theta^2), then loss finalisation by the last block ----
  __shared__ double s_part[8];
  double d = warp_sum_d((double)sumsq);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_part[w];
    atomicAdd(a.acc + 1, tot);
    __threadfence();
    const unsigned int t = atomicAdd(a.ticket, 1u);
    if (t == gridDim.x - 1) {
      __threadfence();
      const double data = *((volatile double *)a.acc);
      const double ss = *((volatile double *)(a.acc + 1));
      if (a.loss_out != nullptr)
        *a.loss_out = (float)((float)(data * (double)a.inv_B) + (float)((double)a.reg_coeff * ss));
      a.acc[0] = 0.0;
      a.acc[1] = 0.0;
      *a.ticket = 0u;
    }
  }
}

static int launch_update(const OptimArgs &a, bool adam, cudaStream_t st) {
  SERT_REQUIRE(a.total % 4 == 0, "parameter arena must be padded to 4 floats");
  SERT_REQUIRE(a.num_segments >= 1 && a.num_segments <= kMaxSegments, "bad segment table");
  // Persistent grid: exactly the number of CTAs that are resident at once (SM count x occupancy), so
  // the grid-stride loop runs as a single wave with no tail.
  static int resident[2] = {0, 0};
  if (resident[adam ? 1 : 0] == 0) {
    int per_sm = 0, sms = kNumSMs, dev = 0;
    SERT_CUDA(cudaGetDevice(&dev));
    SERT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (adam)
      SERT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dense_update_kernel<true>, 256, 0));
    else
      SERT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dense_update_kernel<false>, 256, 0));
    resident[adam ? 1 : 0] = std::max(1, per_sm) * sms;
  }
  const long long total4 = a.total / 4;
  long long blocks = (total4 + 255) / 256;
  if (blocks > resident[adam ? 1 : 0]) blocks = resident[adam ? 1 : 0];
  if (blocks < 1) blocks = 1;
  if (adam)
    dense_update_kernel<true><<<(int)blocks, 256, 0, st>>>(a);
  else
    dense_update_kernel<false><<<(int)blocks, 256, 0, st>>>(a);
  SERT_LAUNCH_CHECK();
  return 0;
}

int launch_adam(const OptimArgs &a, cudaStream_t st) { return launch_update(a, true, st); }
int launch_adadelta(const OptimArgs &a, cudaStream_t st) { return launch_update(a, false, st); }

__global__ void finalize_eval_kernel(double *acc, float *loss_out, float inv_B) {
  *loss_out = (float)(acc[0] * (double)inv_B);
  acc[0] = 0.0;
}

int launch_finalize_eval(double *acc, float *loss_out, float inv_B, cudaStream_t st) {
  finalize_eval_kernel<<<1, 1, 0, st>>>(acc, loss_out, inv_B);
  SERT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sert
