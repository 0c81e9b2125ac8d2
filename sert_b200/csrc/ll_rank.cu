// Log-linear entity ranking on the device: LogLinearCallback.process (bin/query.py:204-233) with
// aggregate_distribution(mode='product') (sert/inference.py:170-183) behind the per-term softmax of predict_fn
// (sert/models.py:880-890).  For query j with terms t = 0..T_j-1 (rows first_j + t of the (rows*W, E) per-term
// matrix):
//     s_j[e]   = sum_t log p_t[e]      over the terms with p_t[e] > 0   (np.ma.log(p).filled(0): exact zeros are skipped)
//     rel_j[e] = exp(s_j[e]) / sum_e' exp(s_j[e'])                       (distribution /= distribution.sum())
//     order    = argsort(rel_j) descending over ALL E entities           (np.argsort(distribution)[::-1])
// plus what the callback writes to its debug file: the normalised entropy of every term's distribution and of rel_j.
// Nothing of size (rows, W, E) or (queries, E) crosses PCIe: the host receives (queries, top) ids and relevances.
// The full descending order is a bitonic sort of 64-bit keys (score bits | ~entity id: ties order by lower id): one
// shared-memory kernel for lists of up to 4096 keys, global compare-exchange steps above that.
#include <math.h>

#include <algorithm>
#include <vector>

#include "kernels.cuh"
#include "ll_kernels.cuh"
#include "topk_keys.cuh"

namespace sert {

namespace {

constexpr int kSortTile = 4096;     // keys a CTA sorts / merges in shared memory (32 KB)

__device__ __forceinline__ double block_sum_d(double v, double *sm) {
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
  r = warp_sum_d(r);
  __syncthreads();
  return r;    // valid in warp 0
}

// rel[j,e] = exp(sum_t log p_t[e]) (unnormalised) and mass[j] += sum_e rel[j,e].  PROBS: `Z` holds probabilities;
// else logits with per-row softmax statistics (p = expf(z - max) / sum, the expression of predict_fn's softmax).
template <bool PROBS>
__global__ void __launch_bounds__(256) ll_rank_aggregate_kernel(const float *__restrict__ Z, const float *__restrict__ rmax,
                                                                const float *__restrict__ rsum,
                                                                const int32_t *__restrict__ first,
                                                                const int32_t *__restrict__ nterms, int E, long long ldz,
                                                                float *__restrict__ rel, double *__restrict__ mass) {
  __shared__ double sm[32];
  const int j = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const long long r0 = first[j];
  const int T = nterms[j];
  float value = 0.f;
  if (e < E) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      const long long r = r0 + t;
      const float z = Z[r * ldz + e];
      const float p = PROBS ? z : expf(z - rmax[r]) / rsum[r];
      if (p > 0.f) acc += logf(p);
    }
    value = expf(acc);
    rel[(size_t)j * E + e] = value;
  }
  const double total = block_sum_d((double)value, sm);
  if (threadIdx.x == 0) atomicAdd(mass + j, total);
}

// rel /= mass; keys; sum and sum p ln p of the normalised distribution (entropy of the final ranking distribution)
__global__ void __launch_bounds__(256) ll_rank_keys_kernel(float *__restrict__ rel, const double *__restrict__ mass, int E,
                                                           int q_begin, long long n_pow2,
                                                           unsigned long long *__restrict__ keys,
                                                           double *__restrict__ norm_sum, double *__restrict__ plogp) {
  __shared__ double sm[32];
  const int j = q_begin + blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  if (e < n_pow2) {
    unsigned long long key = 0ull;
    if (e < E) {
      const float total = (float)mass[j];                    // distribution.sum() is a float32 scalar in the reference
      const float p = rel[(size_t)j * E + e] / total;
      rel[(size_t)j * E + e] = p;
      key = make_key(p, (unsigned int)e);
      s1 = (double)p;
      if (p > 0.f) s2 = (double)p * log((double)p);
    }
    keys[(size_t)blockIdx.y * n_pow2 + e] = key;
  }
  s1 = block_sum_d(s1, sm);
  s2 = block_sum_d(s2, sm);
  if (threadIdx.x == 0) {
    atomicAdd(norm_sum + j, s1);
    atomicAdd(plogp + j, s2);
  }
}

// normalised Shannon entropy of every term row's distribution (compute_normalised_entropy, bin/query.py:370-376):
// scipy.stats.entropy renormalises pk, so H = ln(S1) - S2 / S1 with S1 = sum p, S2 = sum p ln p; divided by ln E.
template <bool PROBS>
__global__ void __launch_bounds__(256) ll_term_entropy_kernel(const float *__restrict__ Z, const float *__restrict__ rmax,
                                                              const float *__restrict__ rsum,
                                                              const int32_t *__restrict__ term_rows, int E, long long ldz,
                                                              float *__restrict__ out) {
  __shared__ double sm[32];
  __shared__ double s1_all;
  const long long r = term_rows[blockIdx.x];
  double s1 = 0.0, s2 = 0.0;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float z = Z[r * ldz + e];
    const float p = PROBS ? z : expf(z - rmax[r]) / rsum[r];
    s1 += (double)p;
    if (p > 0.f) s2 += (double)p * log((double)p);
  }
  s1 = block_sum_d(s1, sm);
  if (threadIdx.x == 0) s1_all = s1;
  s2 = block_sum_d(s2, sm);
  if (threadIdx.x == 0) out[blockIdx.x] = (float)((log(s1_all) - s2 / s1_all) / log((double)E));
}

// ---- bitonic sort, descending, of `segments` independent lists of n_pow2 keys ---------------------------------
__device__ __forceinline__ void cmpx(unsigned long long &a, unsigned long long &b, bool desc) {
  if ((a < b) == desc) { const unsigned long long t = a; a = b; b = t; }
}

// stages size = 2 .. tile of every tile (first == true), or the strides tile/2 .. 1 of stage `size` (merge)
__global__ void __launch_bounds__(512) bitonic_tile_kernel(unsigned long long *__restrict__ keys, long long n_pow2,
                                                           int tile, long long size_in, bool first) {
  extern __shared__ unsigned long long sk[];
  const long long tiles_per_seg = n_pow2 / tile;
  const long long seg = blockIdx.x / tiles_per_seg, tl = blockIdx.x % tiles_per_seg;
  unsigned long long *base = keys + seg * n_pow2 + tl * tile;
  const long long g0 = tl * tile;                           // index of the tile's first key inside its list
  for (int i = threadIdx.x; i < tile; i += blockDim.x) sk[i] = base[i];
  __syncthreads();
  for (long long size = first ? 2 : size_in; size <= (first ? (long long)tile : size_in); size <<= 1) {
    for (int stride = (int)((size >> 1) < tile ? (size >> 1) : (tile >> 1)); stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (tile >> 1); t += blockDim.x) {
        const int l = 2 * t - (t & (stride - 1));
        const bool desc = (((g0 + l) & size) == 0);
        cmpx(sk[l], sk[l + stride], desc);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < tile; i += blockDim.x) base[i] = sk[i];
}

// one compare-exchange step (stride >= tile) of stage `size` over every list
__global__ void __launch_bounds__(256) bitonic_global_kernel(unsigned long long *__restrict__ keys, long long n_pow2,
                                                             long long size, long long stride, long long pairs) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < pairs;
       t += (long long)gridDim.x * blockDim.x) {
    const long long seg = t / (n_pow2 >> 1), u = t % (n_pow2 >> 1);
    const long long l = 2 * u - (u & (stride - 1));
    unsigned long long *b = keys + seg * n_pow2;
    unsigned long long x = b[l], y = b[l + stride];
    const bool desc = ((l & size) == 0);
    if ((x < y) == desc) { b[l] = y; b[l + stride] = x; }
  }
}

__global__ void __launch_bounds__(256) ll_rank_emit_kernel(const unsigned long long *__restrict__ keys, long long n_pow2,
                                                           int top, int q_begin, int32_t *__restrict__ out_idx,
                                                           float *__restrict__ out_rel) {
  const int j = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= top) return;
  const unsigned long long key = keys[(size_t)j * n_pow2 + i];
  out_idx[(size_t)(q_begin + j) * top + i] = (int32_t)key_row(key);
  out_rel[(size_t)(q_begin + j) * top + i] = key_score(key);
}

int sort_desc(unsigned long long *keys, long long segments, long long n_pow2, cudaStream_t st) {
  const int tile = (int)std::min<long long>(n_pow2, kSortTile);
  static std::atomic<uint64_t> configured{0};
  if (first_use_on_device(configured))
    SERT_CUDA(cudaFuncSetAttribute(bitonic_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortTile * 8));
  const long long blocks = segments * (n_pow2 / tile);
  SERT_REQUIRE(blocks < (1ll << 31), "too many sort tiles");
  bitonic_tile_kernel<<<(int)blocks, 512, (size_t)tile * 8, st>>>(keys, n_pow2, tile, 0, true);
  SERT_LAUNCH_CHECK();
  const long long pairs = segments * (n_pow2 >> 1);
  for (long long size = 2ll * tile; size <= n_pow2; size <<= 1) {
    for (long long stride = size >> 1; stride >= tile; stride >>= 1) {
      bitonic_global_kernel<<<(int)std::min<long long>((pairs + 255) / 256, kNumSMs * 32), 256, 0, st>>>(
          keys, n_pow2, size, stride, pairs);
      SERT_LAUNCH_CHECK();
    }
    bitonic_tile_kernel<<<(int)blocks, 512, (size_t)tile * 8, st>>>(keys, n_pow2, tile, size, false);
    SERT_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace

size_t ll_rank_scratch_bytes(int nq, int E, int n_terms) {
  long long n_pow2 = 32;
  while (n_pow2 < E) n_pow2 <<= 1;
  // keys of one chunk of queries (at most 256 MB), three doubles per query, term entropies, int arrays
  const long long per_query = n_pow2 * 8;
  const long long chunk = std::max<long long>(1, std::min<long long>(nq, (256ll << 20) / per_query));
  return (size_t)(chunk * per_query) + (size_t)nq * (3 * sizeof(double) + 2 * sizeof(int32_t) + 64) +
         (size_t)n_terms * (sizeof(float) + sizeof(int32_t)) + 4096;
}

// Z: (rows, E) logits with statistics (probs == false) or probabilities (probs == true), device.  first / nterms: host.
// rel_dev: (nq, E) scratch.  Outputs: host arrays.  Synchronises the stream.
int ll_rank(const float *Z, const float *rmax, const float *rsum, bool probs, long long ldz, int E,
            const int32_t *first_host, const int32_t *nterms_host, int nq, int top, float *rel_dev, void *scratch,
            size_t scratch_bytes, int32_t *out_idx_host, float *out_rel_host, float *out_term_entropy_host,
            float *out_entropy_host, float *out_mass_host, cudaStream_t st) {
  SERT_REQUIRE(nq >= 0 && E >= 1 && top >= 1 && top <= E, "bad ranking shape");
  if (nq == 0) return 0;
  long long n_terms = 0;
  for (int j = 0; j < nq; ++j) n_terms += nterms_host[j];
  SERT_REQUIRE(scratch_bytes >= ll_rank_scratch_bytes(nq, E, (int)n_terms), "ranking scratch too small");
  long long n_pow2 = 32;
  while (n_pow2 < E) n_pow2 <<= 1;
  const long long per_query = n_pow2 * 8;
  const int chunk = (int)std::max<long long>(1, std::min<long long>(nq, (256ll << 20) / per_query));
  char *p = static_cast<char *>(scratch);
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(p); p += (size_t)chunk * per_query;
  double *mass = reinterpret_cast<double *>(p); p += (size_t)nq * sizeof(double);
  double *norm_sum = reinterpret_cast<double *>(p); p += (size_t)nq * sizeof(double);
  double *plogp = reinterpret_cast<double *>(p); p += (size_t)nq * sizeof(double);
  int32_t *first = reinterpret_cast<int32_t *>(p); p += align_up((size_t)nq * sizeof(int32_t), 16);
  int32_t *nterms = reinterpret_cast<int32_t *>(p); p += align_up((size_t)nq * sizeof(int32_t), 16);
  int32_t *term_rows = reinterpret_cast<int32_t *>(p); p += align_up((size_t)n_terms * sizeof(int32_t), 16);
  float *term_entropy = reinterpret_cast<float *>(p);

  std::vector<int32_t> rows_host((size_t)n_terms);
  {
    size_t o = 0;
    for (int j = 0; j < nq; ++j)
      for (int t = 0; t < nterms_host[j]; ++t) rows_host[o++] = first_host[j] + t;
  }
  SERT_CUDA(cudaMemcpyAsync(first, first_host, (size_t)nq * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  SERT_CUDA(cudaMemcpyAsync(nterms, nterms_host, (size_t)nq * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  SERT_CUDA(cudaMemcpyAsync(term_rows, rows_host.data(), (size_t)n_terms * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  SERT_CUDA(cudaMemsetAsync(mass, 0, (size_t)nq * 3 * sizeof(double), st));
  dim3 grid(cdiv(E, 256), nq);
  if (probs) ll_rank_aggregate_kernel<true><<<grid, 256, 0, st>>>(Z, rmax, rsum, first, nterms, E, ldz, rel_dev, mass);
  else ll_rank_aggregate_kernel<false><<<grid, 256, 0, st>>>(Z, rmax, rsum, first, nterms, E, ldz, rel_dev, mass);
  SERT_LAUNCH_CHECK();
  if (out_term_entropy_host && n_terms > 0) {
    if (probs) ll_term_entropy_kernel<true><<<(int)n_terms, 256, 0, st>>>(Z, rmax, rsum, term_rows, E, ldz, term_entropy);
    else ll_term_entropy_kernel<false><<<(int)n_terms, 256, 0, st>>>(Z, rmax, rsum, term_rows, E, ldz, term_entropy);
    SERT_LAUNCH_CHECK();
  }
  int32_t *out_idx_dev = nullptr;
  float *out_rel_dev = nullptr;
  SERT_CUDA(cudaMallocAsync(&out_idx_dev, (size_t)nq * top * sizeof(int32_t), st));
  SERT_CUDA(cudaMallocAsync(&out_rel_dev, (size_t)nq * top * sizeof(float), st));
  for (int q0 = 0; q0 < nq; q0 += chunk) {
    const int n = std::min(chunk, nq - q0);
    dim3 kgrid((unsigned)((n_pow2 + 255) / 256), n);
    ll_rank_keys_kernel<<<kgrid, 256, 0, st>>>(rel_dev, mass, E, q0, n_pow2, keys, norm_sum, plogp);
    SERT_LAUNCH_CHECK();
    if (sort_desc(keys, n, n_pow2, st)) return -1;
    dim3 egrid(cdiv(top, 256), n);
    ll_rank_emit_kernel<<<egrid, 256, 0, st>>>(keys, n_pow2, top, q0, out_idx_dev, out_rel_dev);
    SERT_LAUNCH_CHECK();
  }
  SERT_CUDA(cudaMemcpyAsync(out_idx_host, out_idx_dev, (size_t)nq * top * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  SERT_CUDA(cudaMemcpyAsync(out_rel_host, out_rel_dev, (size_t)nq * top * sizeof(float), cudaMemcpyDeviceToHost, st));
  std::vector<double> sums((size_t)nq * 3);
  SERT_CUDA(cudaMemcpyAsync(sums.data(), mass, (size_t)nq * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (out_term_entropy_host && n_terms > 0)
    SERT_CUDA(cudaMemcpyAsync(out_term_entropy_host, term_entropy, (size_t)n_terms * sizeof(float),
                              cudaMemcpyDeviceToHost, st));
  SERT_CUDA(cudaFreeAsync(out_idx_dev, st));
  SERT_CUDA(cudaFreeAsync(out_rel_dev, st));
  SERT_CUDA(cudaStreamSynchronize(st));
  for (int j = 0; j < nq; ++j) {
    const double s1 = sums[(size_t)nq + j], s2 = sums[(size_t)2 * nq + j];
    if (out_mass_host) out_mass_host[j] = (float)s1;
    if (out_entropy_host) out_entropy_host[j] = s1 > 0.0 ? (float)((log(s1) - s2 / s1) / log((double)E)) : 0.f;
  }
  return 0;
}

}  // namespace sert
