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

// One element of lasagne.updates.adam / adadelta (Lasagne 0.1 forms); gj already holds the L2 term.
template <bool ADAM>
__device__ __forceinline__ void update_element(float &p, float &s1, float &s2, float gj, float c0, float c1, float c2,
                                               float c3) {
  if (ADAM) {
    // m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; theta <- theta - a_t m / (sqrt(v) + eps)
    const float m = c1 * s1 + (1.0f - c1) * gj;
    const float v = c2 * s2 + (1.0f - c2) * gj * gj;
    p = p - __fdividef(c0 * m, fast_sqrt(v) + c3);
    s1 = m; s2 = v;
  } else {
    // accu <- rho accu + (1-rho) g^2 ; upd = g sqrt(delta+eps)/sqrt(accu+eps) ;
    // theta <- theta - lr upd ; delta <- rho delta + (1-rho) upd^2
    const float accu = c1 * s1 + (1.0f - c1) * gj * gj;
    const float upd = __fdividef(gj * fast_sqrt(s2 + c3), fast_sqrt(accu + c3));
    p = p - c0 * upd;
    s1 = accu; s2 = c1 * s2 + (1.0f - c1) * upd * upd;
  }
}

// ---- optimiser state as bfloat16 (sert_config.dtype_mode 1; BASELINE.json configs[1] "bf16") ----------------------
// Four consecutive state values = one 8-byte access.  Stores round stochastically: Adam's second moment moves by
// 0.1 % per step (beta2 = 0.999), less than half a bf16 ulp (0.2-0.4 %), so round-to-nearest would freeze it; adding
// 16 uniform random bits below the kept mantissa before truncating makes the stored value unbiased.  The bits come
// from a hash of (element index, step): no state, reproducible.
template <bool S16>
__device__ __forceinline__ void load_state4(const float *base, long long i4, float (&v)[4]) {
  if (S16) {
    const uint2 u = reinterpret_cast<const uint2 *>(base)[i4];
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  } else {
    const float4 x = reinterpret_cast<const float4 *>(base)[i4];
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  }
}
__device__ __forceinline__ uint32_t sr_hash(uint32_t element, uint32_t step) {
  uint32_t h = element * 0x9E3779B1u ^ step * 0x85EBCA77u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}
// top 16 bits of (bits + 16 random bits): a value is rounded away from zero with probability = its discarded fraction
__device__ __forceinline__ uint32_t bf16_sr(float x, uint32_t r16) { return (__float_as_uint(x) + r16) >> 16; }
template <bool S16>
__device__ __forceinline__ void store_state4(float *base, long long i4, const float (&v)[4], uint32_t element,
                                             uint32_t step, uint32_t salt) {
  if (S16) {
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t h = sr_hash(element + j, step);
      r[j] = salt ? (h >> 16) : (h & 0xffffu);
    }
    uint2 u;
    u.x = bf16_sr(v[0], r[0]) | (bf16_sr(v[1], r[1]) << 16);
    u.y = bf16_sr(v[2], r[2]) | (bf16_sr(v[3], r[3]) << 16);
    reinterpret_cast<uint2 *>(base)[i4] = u;
  } else {
    reinterpret_cast<float4 *>(base)[i4] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// One 16-byte chunk of each stream per thread, one-shot grid (no grid-stride loop): measured on B200
// (tools/bw_probe.cu) the in-place 3-stream read-modify-write reaches 6.50 TB/s this way vs 5.3-6.2 TB/s
// for persistent grid-stride variants -- the block scheduler interleaves the load and store phases of
// many short CTAs better than a resident wave does.
// PHASE 0: everything.  PHASE 3: only the row-stamped tables (word / entity representations).  PHASE 4: only the
// dense tensors (projection matrix, bias), whose gradients are produced by two small kernels that the caller
// overlaps with phase 3 on a second stream.
template <bool ADAM, int PHASE, bool S16>
__global__ void __launch_bounds__(256) dense_update_kernel(OptimArgs a) {
  const long long total4 = a.last4 >= 0 ? a.last4 : (a.total >> 2);
  const float4 *__restrict__ th4 = reinterpret_cast<const float4 *>(a.theta);
  float4 *__restrict__ out4 = reinterpret_cast<float4 *>(a.theta_out != nullptr ? a.theta_out : a.theta);
  float4 *__restrict__ g4 = reinterpret_cast<float4 *>(a.grad);
  float sumsq = 0.f;

  // One 16-byte chunk of theta / gradient per thread, whatever the state type: with bfloat16 state the 8-byte state
  // accesses of a warp still fill whole 32-byte sectors, and tools/bf16_state_probe.cu measures 6.23 TB/s for this
  // shape against 5.0-5.7 TB/s for 8 elements per thread (16-byte state accesses or two adjacent chunks).
  constexpr int CH = 1;
  const long long base4 = a.first4 + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * CH;
  // phase A: where each chunk lives, whether its row was touched, and every load of every chunk in flight at once
  int sgi[CH];
  bool act[CH], touched[CH];
  uint32_t push[CH];               // ranks the new chunk is sent to (bit r = rank r)
  float4 p[CH], g[CH];
  float v1[CH][4], v2[CH][4];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    const long long i4 = base4 + ch;
    act[ch] = false; touched[ch] = false; sgi[ch] = 0;
    g[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i4 < total4) {
      const long long e = i4 << 2;
      int s = 0;
#pragma unroll
      for (int q = 1; q < kMaxSegments; ++q) s += (q < a.num_segments && e >= a.seg[q].offset) ? 1 : 0;
      const ParamSegment &sg = a.seg[s];
      const bool live = (e - sg.offset) < sg.count;   // false only inside inter-segment padding
      bool tch = true, skip = false;
      push[ch] = a.n_peers > 0 ? 0xffffffffu : 0u;
      if (live && sg.flags != nullptr) {
        const unsigned int row = (unsigned int)(e - sg.offset) / (unsigned int)sg.row_len;
        const uint32_t flag = __ldg(sg.flags + row);
        tch = (flag == a.stamp);
        skip = (flag == kHotRowMark);       // updated by hot_update_kernel on the side stream
        // table shards with look-ahead: the other ranks only need the rows their next batch reads
        // (instance shards: the need word is the mask of the ranks whose instances read the row)
        if (a.n_peers > 0 && a.push_all == 0 && a.need[s] != nullptr) {
          const uint32_t nd = __ldg(a.need[s] + row);
          push[ch] = a.need_is_mask ? nd : (nd == a.need_stamp ? 0xffffffffu : 0u);
        }
      }
      const bool mine = PHASE == 0 ? true : PHASE == 3 ? (sg.flags != nullptr) : (sg.flags == nullptr);
      sgi[ch] = s;
      act[ch] = live && mine && !skip;
      touched[ch] = act[ch] && tch;
      if (act[ch]) {
        p[ch] = th4[i4];
        load_state4<S16>(a.s1, i4, v1[ch]);
        load_state4<S16>(a.s2, i4, v2[ch]);
        if (touched[ch]) g[ch] = __ldcg(g4 + i4);
      }
    }
  }
  // phase B: update and store
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
    if (act[ch]) {
      const long long i4 = base4 + ch;
      const long long e = i4 << 2;
      const int s = sgi[ch];
      const ParamSegment &sg = a.seg[s];
      const float l2 = sg.regularised ? a.l2_scale : 0.0f;
      if (sg.regularised == 1)
        sumsq += p[ch].x * p[ch].x + p[ch].y * p[ch].y + p[ch].z * p[ch].z + p[ch].w * p[ch].w;
      float pv[4] = {p[ch].x, p[ch].y, p[ch].z, p[ch].w};
      const float gv[4] = {g[ch].x, g[ch].y, g[ch].z, g[ch].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        update_element<ADAM>(pv[j], v1[ch][j], v2[ch][j], gv[j] + l2 * pv[j], a.c0, a.c1, a.c2, a.c3);
      }
      const float4 fresh = make_float4(pv[0], pv[1], pv[2], pv[3]);
      out4[i4] = fresh;
      if (push[ch])
        for (int pr = 0; pr < a.n_peers; ++pr)
          if ((push[ch] >> a.peer_rank[pr]) & 1u) reinterpret_cast<float4 *>(a.peer_theta[pr])[i4] = fresh;   // NVLink stores
      if (PHASE == 4 && a.transposed != nullptr && s == a.transposed_segment) {
        // keep the (cols, rows) copy of the projection matrix current for the next step's back-projection
        const unsigned int el = (unsigned int)(e - sg.offset);
        const unsigned int r = el / (unsigned int)sg.row_len, c = el % (unsigned int)sg.row_len;
#pragma unroll
        for (int j = 0; j < 4; ++j) a.transposed[(size_t)(c + j) * a.transposed_rows + r] = pv[j];
      }
      store_state4<S16>(a.s1, i4, v1[ch], (uint32_t)e, a.stamp, 0u);
      store_state4<S16>(a.s2, i4, v2[ch], (uint32_t)e, a.stamp, 1u);
      // The gradient row is zeroed only now, behind the stores that depend on its value: a store issued
      // right behind the load of the same address stalls the LSU until the load returns and costs 2.6x
      // (measured with tools/bw_probe.cu: 185 us vs 71 us for this stream on B200).
      if (touched[ch]) g4[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
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
  if (PHASE == 4 && a.ticket != nullptr) {
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
  SERT_REQUIRE(a.transposed == nullptr || a.seg[a.transposed_segment].row_len % 4 == 0,
               "transposed copy needs rows of 4n floats");
  const long long total4 = a.last4 >= 0 ? a.last4 : a.total / 4;
  SERT_REQUIRE(total4 <= a.total / 4 && a.first4 <= total4, "bad chunk range");
  if (a.first4 == total4) return 0;                          // an empty piece
  const long long per_block = 256;                           // 16-byte chunks per block (dense_update_kernel: CH = 1)
  const long long blocks = std::max<long long>(1, (total4 - a.first4 + per_block - 1) / per_block);
  SERT_REQUIRE(blocks < (1ll << 31), "parameter arena too large for one launch");
  const int g = (int)blocks;
#define SERT_UPD(ADAM_, S16_)                                                              \
  do {                                                                                     \
    if (a.phase == 3) dense_update_kernel<ADAM_, 3, S16_><<<g, 256, 0, st>>>(a);           \
    else if (a.phase == 4) dense_update_kernel<ADAM_, 4, S16_><<<g, 256, 0, st>>>(a);      \
    else dense_update_kernel<ADAM_, 0, S16_><<<g, 256, 0, st>>>(a);                        \
  } while (0)
  if (adam) { if (a.state_bf16) SERT_UPD(true, true); else SERT_UPD(true, false); }
  else { if (a.state_bf16) SERT_UPD(false, true); else SERT_UPD(false, false); }
#undef SERT_UPD
  SERT_LAUNCH_CHECK();
  if (a.phase == 0 && !a.no_finalize) {          // the loss is complete once the last phase of the step has run (phase 4 finalises itself)
    finalize_train_kernel<<<1, kSumsqSlots, 0, st>>>(a.acc, a.loss_out, a.inv_B, a.reg_coeff);
    SERT_LAUNCH_CHECK();
  }
  return 0;
}

// Adam + L2 for the hot word rows of the fused tile kernel (kernels.cuh: VsFusedArgs::hot_slot): the gradient of hot
// row s is the sum of its kHotReplicas private copies (plus whatever sits in the gradient row itself); the copies
// are zeroed for the next step.  One CTA per hot row, one 16-byte chunk per thread, off the step's critical path.
template <bool S16>
__global__ void __launch_bounds__(128) hot_update_kernel(HotUpdateArgs h) {
  const int s = blockIdx.x;
  const int row = __ldg(h.hot_ids + s);
  const long long row4 = (long long)(((size_t)h.table_offset + (size_t)row * h.d) / 4);
  if (row4 < h.own_lo4 || row4 >= h.own_hi4) return;   // table shards: another rank's row (its gradient is not formed here)
  float sumsq = 0.f;
  for (int c = threadIdx.x; c < h.d / 4; c += blockDim.x) {
    float4 v[kHotReplicas];
#pragma unroll
    for (int r = 0; r < kHotReplicas; ++r)
      v[r] = __ldcg(reinterpret_cast<const float4 *>(h.hot_acc + ((size_t)r * kMaxHotRows + s) * h.d) + c);
    const size_t i4 = ((size_t)h.table_offset + (size_t)row * h.d) / 4 + c;
    float4 g = __ldcg(reinterpret_cast<const float4 *>(h.grad) + i4);
#pragma unroll
    for (int r = 0; r < kHotReplicas; ++r) { g.x += v[r].x; g.y += v[r].y; g.z += v[r].z; g.w += v[r].w; }
    const float4 p = reinterpret_cast<const float4 *>(h.theta)[i4];
    float v1[4], v2[4];
    load_state4<S16>(h.s1, (long long)i4, v1);
    load_state4<S16>(h.s2, (long long)i4, v2);
    sumsq += p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w;
    float pv[4] = {p.x, p.y, p.z, p.w};
    const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      update_element<true>(pv[j], v1[j], v2[j], gv[j] + h.l2_scale * pv[j], h.c0, h.c1, h.c2, h.c3);
    const float4 fresh = make_float4(pv[0], pv[1], pv[2], pv[3]);
    reinterpret_cast<float4 *>(h.theta_out != nullptr ? h.theta_out : h.theta)[i4] = fresh;
    for (int pr = 0; pr < h.n_peers; ++pr) reinterpret_cast<float4 *>(h.peer_theta[pr])[i4] = fresh;
    store_state4<S16>(h.s1, (long long)i4, v1, (uint32_t)(i4 * 4), h.stamp, 0u);
    store_state4<S16>(h.s2, (long long)i4, v2, (uint32_t)(i4 * 4), h.stamp, 1u);
    reinterpret_cast<float4 *>(h.grad)[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kHotReplicas; ++r)
      reinterpret_cast<float4 *>(h.hot_acc + ((size_t)r * kMaxHotRows + s) * h.d)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __shared__ double s_part[4];
  double d = warp_sum_d(h.counted ? (double)sumsq : 0.0);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_part[w];
    if (tot != 0.0) atomicAdd(h.acc + 1 + (s & (kSumsqSlots - 1)), tot);
  }
}

// Instance shards: this rank's share of a hot row's gradient -> the row's owner (kernels.cuh: HotPushArgs)
__global__ void __launch_bounds__(128) hot_push_kernel(HotPushArgs h) {
  const int s = blockIdx.x;
  const int row = __ldg(h.hot_ids + s);
  int o = 0;
  for (int q = 1; q < h.n_owner; ++q) o += row >= h.r_bound[q] ? 1 : 0;
  float *dst = h.grad_peer[o] + (size_t)h.table_offset + (size_t)row * h.d;
  for (int c = threadIdx.x; c < h.d / 4; c += blockDim.x) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kHotReplicas; ++r) {
      float4 *src = reinterpret_cast<float4 *>(h.hot_acc + ((size_t)r * kMaxHotRows + s) * h.d) + c;
      const float4 v = __ldcg(src);
      g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
      *src = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    red_add_f4(dst + c * 4, g);
  }
}

int launch_hot_push(const HotPushArgs &h, cudaStream_t st) {
  if (h.n_hot <= 0) return 0;
  SERT_REQUIRE(h.d % 4 == 0 && h.table_offset % 4 == 0, "hot rows must be 16-byte aligned");
  SERT_REQUIRE(h.n_owner >= 1 && h.n_owner <= kMaxPeers + 1, "bad owner table");
  const int threads = std::min(128, std::max(32, (h.d / 4 + 31) / 32 * 32));
  hot_push_kernel<<<h.n_hot, threads, 0, st>>>(h);
  SERT_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) push_add_kernel(float4 *__restrict__ src, float *__restrict__ dst, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldcg(src + i);
  src[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  red_add_f4(dst + i * 4, v);
}

int launch_push_add(float *src, float *dst, long long n, cudaStream_t st) {
  SERT_REQUIRE(n % 4 == 0, "push_add: the range must be a multiple of 4 floats");
  if (n == 0) return 0;
  push_add_kernel<<<cdiv(n / 4, 256), 256, 0, st>>>(reinterpret_cast<float4 *>(src), dst, n / 4);
  SERT_LAUNCH_CHECK();
  return 0;
}

int launch_hot_update(const HotUpdateArgs &h, cudaStream_t st) {
  if (h.n_hot <= 0) return 0;
  SERT_REQUIRE(h.d % 4 == 0 && h.table_offset % 4 == 0, "hot rows must be 16-byte aligned");
  const int threads = std::min(128, std::max(32, (h.d / 4 + 31) / 32 * 32));
  if (h.state_bf16) hot_update_kernel<true><<<h.n_hot, threads, 0, st>>>(h);
  else hot_update_kernel<false><<<h.n_hot, threads, 0, st>>>(h);
  SERT_LAUNCH_CHECK();
  return 0;
}

__global__ void hot_mark_kernel(uint32_t *flags, const int32_t *ids, int n, uint32_t value) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) flags[ids[s]] = value;
}

int launch_hot_mark(uint32_t *flags, const int32_t *hot_ids, int n_hot, uint32_t value, cudaStream_t st) {
  if (n_hot <= 0) return 0;
  hot_mark_kernel<<<1, 64, 0, st>>>(flags, hot_ids, n_hot, value);
  SERT_LAUNCH_CHECK();
  return 0;
}

int launch_finalize_train(double *acc, float *loss_out, float inv_B, float reg_coeff, cudaStream_t st) {
  finalize_train_kernel<<<1, kSumsqSlots, 0, st>>>(acc, loss_out, inv_B, reg_coeff);
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
