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

// sqrt.approx / div.approx: 1-2 ulp, branch-free (the IEEE sequences carry a slow-path branch per element,
// which serialises the warps' load and compute phases and costs ~40% of the achievable bandwidth here).
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One 16-byte chunk of each stream per thread, one-shot grid (no grid-stride loop): measured on B200
// (tools/bw_probe.cu) the in-place 3-stream read-modify-write reaches 6.50 TB/s this way vs 5.3-6.2 TB/s
// for persistent grid-stride variants -- the block scheduler interleaves the load and store phases of
// many short CTAs better than a resident wave does.
// PHASE 0: everything.  PHASE 3: only the row-stamped tables (word / entity representations).  PHASE 4: only the
// dense tensors (projection matrix, bias), whose gradients are produced by two small kernels that the caller
// overlaps with phase 3 on a second stream.
template <bool ADAM, int PHASE>
__global__ void __launch_bounds__(256) dense_update_kernel(OptimArgs a) {
  const long long total4 = a.total >> 2;
  float4 *__restrict__ th4 = reinterpret_cast<float4 *>(a.theta);
  float4 *__restrict__ s14 = reinterpret_cast<float4 *>(a.s1);
  float4 *__restrict__ s24 = reinterpret_cast<float4 *>(a.s2);
  float4 *__restrict__ g4 = reinterpret_cast<float4 *>(a.grad);
  const float one_m_c1 = 1.0f - a.c1;
  const float one_m_c2 = 1.0f - a.c2;
  float sumsq = 0.f;

  const long long i4 = a.first4 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 < total4) {
    const long long e = i4 << 2;
    int s = 0;
#pragma unroll
    for (int q = 1; q < kMaxSegments; ++q) s += (q < a.num_segments && e >= a.seg[q].offset) ? 1 : 0;
    const ParamSegment &sg = a.seg[s];
    const bool live = (e - sg.offset) < sg.count;   // false only inside inter-segment padding
    bool touched = true;
    if (live && sg.flags != nullptr) {
      const unsigned int row = (unsigned int)(e - sg.offset) / (unsigned int)sg.row_len;
      touched = (__ldg(sg.flags + row) == a.stamp);
    }
    const bool mine = PHASE == 0 ? true : PHASE == 3 ? (sg.flags != nullptr) : (sg.flags == nullptr);
    if (live && mine) {
      const float4 p = th4[i4];
      const float4 x1 = s14[i4];
      const float4 x2 = s24[i4];
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (touched) g = __ldcg(g4 + i4);
      const float l2 = sg.regularised ? a.l2_scale : 0.0f;
      if (sg.regularised == 1) sumsq = p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w;
      float pv[4] = {p.x, p.y, p.z, p.w};
      float v1[4] = {x1.x, x1.y, x1.z, x1.w};
      float v2[4] = {x2.x, x2.y, x2.z, x2.w};
      const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gj = gv[j] + l2 * pv[j];
        if (ADAM) {
          // m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; theta <- theta - a_t m / (sqrt(v) + eps)
          const float m = a.c1 * v1[j] + one_m_c1 * gj;
          const float v = a.c2 * v2[j] + one_m_c2 * gj * gj;
          pv[j] = pv[j] - __fdividef(a.c0 * m, fast_sqrt(v) + a.c3);
          v1[j] = m; v2[j] = v;
        } else {
          // accu <- rho accu + (1-rho) g^2 ; upd = g sqrt(delta+eps)/sqrt(accu+eps) ;
          // theta <- theta - lr upd ; delta <- rho delta + (1-rho) upd^2
          const float accu = a.c1 * v1[j] + one_m_c1 * gj * gj;
          const float upd = __fdividef(gj * fast_sqrt(v2[j] + a.c3), fast_sqrt(accu + a.c3));
          pv[j] = pv[j] - a.c0 * upd;
          v1[j] = accu; v2[j] = a.c1 * v2[j] + one_m_c1 * upd * upd;
        }
      }
      th4[i4] = make_float4(pv[0], pv[1], pv[2], pv[3]);
      if (PHASE == 4 && a.transposed != nullptr && s == a.transposed_segment) {
        // keep the (cols, rows) copy of the projection matrix current for the next step's back-projection
        const unsigned int el = (unsigned int)(e - sg.offset);
        const unsigned int r = el / (unsigned int)sg.row_len, c = el % (unsigned int)sg.row_len;
#pragma unroll
        for (int j = 0; j < 4; ++j) a.transposed[(size_t)(c + j) * a.transposed_rows + r] = pv[j];
      }
      s14[i4] = make_float4(v1[0], v1[1], v1[2], v1[3]);
      s24[i4] = make_float4(v2[0], v2[1], v2[2], v2[3]);
      // The gradient row is zeroed only now, behind the stores that depend on its value: a store issued
      // right behind the load of the same address stalls the LSU until the load returns and costs 2.6x
      // (measured with tools/bw_probe.cu: 185 us vs 71 us for this stream on B200).
      if (touched) g4[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  // ---- block partial of sum(theta^2): fire-and-forget f64 reductions spread over 64 slots (no fence, no
  // ticket: the loss is finalised by a one-block kernel behind this one on the stream) ----
  __shared__ double s_part[8];
  double d = warp_sum_d((double)sumsq);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_part[w];
    if (tot != 0.0) atomicAdd(a.acc + 1 + (blockIdx.x & (kSumsqSlots - 1)), tot);
  }
  if (PHASE == 4) {
    // The dense tensors are a handful of blocks: the last one to finish writes the step's loss (what
    // finalize_train_kernel does behind the other phases) -- one launch less on the step's critical path.
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {                       // block-uniform
      __threadfence();
      double v = 0.0;
      if (threadIdx.x < kSumsqSlots) {  // the slots are read in parallel: 64 dependent L2 round trips cost ~15 us
        v = __ldcg(a.acc + 1 + threadIdx.x);
        a.acc[1 + threadIdx.x] = 0.0;
      }
      v = warp_sum_d(v);
      if (threadIdx.x < kSumsqSlots && (threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
        const double ss = s_part[0] + s_part[1];
        if (a.loss_out != nullptr)
          *a.loss_out = (float)((float)(__ldcg(a.acc) * (double)a.inv_B) + (float)((double)a.reg_coeff * ss));
        a.acc[0] = 0.0;
        *a.ticket = 0u;
      }
    }
  }
}

// loss[slot] = mean data loss + lambda/(2B) * sum(theta^2) (sert/models.py:745-755,773-793); resets the accumulators
__global__ void finalize_train_kernel(double *acc, float *loss_out, float inv_B, float reg_coeff) {
  __shared__ double s[kSumsqSlots];
  s[threadIdx.x] = acc[1 + threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    double ss = 0.0;
    for (int i = 0; i < kSumsqSlots; ++i) ss += s[i];
    if (loss_out != nullptr) *loss_out = (float)((float)(acc[0] * (double)inv_B) + (float)((double)reg_coeff * ss));
    acc[0] = 0.0;
  }
  acc[1 + threadIdx.x] = 0.0;
}

static int launch_update(const OptimArgs &a, bool adam, cudaStream_t st) {
  SERT_REQUIRE(a.total % 4 == 0, "parameter arena must be padded to 4 floats");
  SERT_REQUIRE(a.num_segments >= 1 && a.num_segments <= kMaxSegments, "bad segment table");
  SERT_REQUIRE(a.phase == 0 || a.phase == 3 || a.phase == 4, "bad update phase");
  SERT_REQUIRE(a.phase != 4 || a.ticket != nullptr, "phase 4 needs a ticket counter");
  SERT_REQUIRE(a.transposed == nullptr || a.seg[a.transposed_segment].row_len % 4 == 0,
               "transposed copy needs rows of 4n floats");
  const long long total4 = a.total / 4;
  const long long blocks = std::max<long long>(1, (total4 - a.first4 + 255) / 256);
  SERT_REQUIRE(blocks < (1ll << 31), "parameter arena too large for one launch");
  const int g = (int)blocks;
  if (adam) {
    if (a.phase == 3) dense_update_kernel<true, 3><<<g, 256, 0, st>>>(a);
    else if (a.phase == 4) dense_update_kernel<true, 4><<<g, 256, 0, st>>>(a);
    else dense_update_kernel<true, 0><<<g, 256, 0, st>>>(a);
  } else {
    if (a.phase == 3) dense_update_kernel<false, 3><<<g, 256, 0, st>>>(a);
    else if (a.phase == 4) dense_update_kernel<false, 4><<<g, 256, 0, st>>>(a);
    else dense_update_kernel<false, 0><<<g, 256, 0, st>>>(a);
  }
  SERT_LAUNCH_CHECK();
  if (a.phase == 0) {          // the loss is complete once the last phase of the step has run (phase 4 finalises itself)
    finalize_train_kernel<<<1, kSumsqSlots, 0, st>>>(a.acc, a.loss_out, a.inv_B, a.reg_coeff);
    SERT_LAUNCH_CHECK();
  }
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
