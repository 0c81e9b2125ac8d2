// Launchers for the log-linear row kernels (see ll_kernels.cu).
#pragma once

#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"

namespace sert {

int launch_ll_row_stats(const float *Z, int64_t rows, int E, int64_t ldz, float *rmax, float *rsum,
                        cudaStream_t st);
int launch_ll_softmax_inplace(float *Z, int64_t rows, int E, int64_t ldz, const float *rmax,
                              const float *rsum, cudaStream_t st);
int launch_ll_joint(const float *Z, const float *rmax, const float *rsum, float *S, int B, int W, int E,
                    int64_t ldz, int64_t lds, cudaStream_t st, float *lrsum_scratch, float2 *sstats = nullptr);
// sstats (optional, (B, ll_joint_slots(E)) float2): per instance and slot of kJointSlot entities, (max, sum of
// exp(S - max)) of the joint logits the pass has just written -- launch_ll_instance then needs no pass of its own over
// S for the softmax statistics.  Written only when E and the row strides are multiples of 4 (the vector path);
// launch_ll_joint returns 1 (instead of 0) when it has written them.
constexpr int kJointSlot = 1024;
static inline int ll_joint_slots(long long E) { return (int)((E + kJointSlot - 1) / kJointSlot); }

struct LlInstanceArgs {
  const float *S;          // (B,E) joint logits
  float *DS;               // (B,E) out: d loss / d S (train) or nullptr
  long long lds;
  int B, E;
  const long long *indptr; // CSR row pointers of this batch's first row (B+1 entries readable)
  long long nnz_base;      // subtracted from indptr values before indexing indices/data
  const int32_t *indices;
  const float *data;
  const float *w;          // (B,) or nullptr
  float inv_B;
  float *ell_out;          // (B,) or nullptr
  double *loss_acc;
  bool train;
  const float2 *sstats = nullptr;   // per-slot softmax statistics of S left by launch_ll_joint, or nullptr
  int slots = 0;
};
int launch_ll_instance(const LlInstanceArgs &a, cudaStream_t st);

// dZ[i,w,:] = p * (dp - sum_e dp*p), dp = ds/clip(p) on the unclipped entries; in place over Z.
// mode 0: both halves; 1: racc[r] = partial row sum only; 2: apply with racc[r] given (entity-sharded step).
int launch_ll_dz(float *Z, const float *rmax, const float *rsum, const float *DS, int B, int W, int E,
                 int64_t ldz, int64_t lds, cudaStream_t st, int mode = 0, float *racc = nullptr);

// (rows, slots) pairs of (max, sum exp(v - max)) over 64-column slices -> per-row softmax statistics
int launch_ll_combine_slices(const float2 *stats, int64_t rows, int slots, int64_t ld, float *rmax, float *rsum,
                             cudaStream_t st);

// ---- fused backward tail of the tensor-core path: acc_r in the log domain, then dZ straight into the split bf16
// operands of the two gradient GEMMs (see ll_kernels.cu).  lrsum = log(rsum) (launch_ll_joint leaves it in its scratch).
int launch_ll_racc_log(const float *Z, const float *rmax, const float *lrsum, const float *DS, int B, int W, int E,
                       int64_t ldz, int64_t lds, float *racc, cudaStream_t st);
int launch_ll_dz_split(const float *Z, const float *rmax, const float *lrsum, const float *racc, const float *DS,
                       int B, int W, int E, int64_t ldz, int64_t lds, int terms, __nv_bfloat16 *dZs, __nv_bfloat16 *dZT_s,
                       cudaStream_t st);   // terms: 3 = [hi|hi|mid] / [hi|mid|hi] rows, 2 = [hi|mid] (pair operands)

// ---- entity-sharded softmax pieces (columns [e_begin, e_begin+E) of the entity axis live on this rank) ----
// parts [shard][2][rows] of gathered (row max, row sum) -> global statistics
int launch_ll_combine_stats(const float *parts, int shards, int64_t rows, float *rmax, float *rsum,
                            cudaStream_t st);
// label terms owned by this shard: partial ell (into loss_acc / ell_out) and partial sum_e do*o (adot_out, train)
int launch_ll_shard_labels(const LlInstanceArgs &a, const float *smax, const float *ssum, int e_begin,
                           float *adot_out, cudaStream_t st);
// ds over the shard's columns given the all-reduced adot
int launch_ll_shard_ds(const LlInstanceArgs &a, const float *smax, const float *ssum, int e_begin,
                       const float *adot_all, cudaStream_t st);

// ---- entity ranking of queries (ll_rank.cu): product of the terms' distributions, renormalised, ranked over all E ----
size_t ll_rank_scratch_bytes(int nq, int E, int n_terms);
int ll_rank(const float *Z, const float *rmax, const float *rsum, bool probs, long long ldz, int E,
            const int32_t *first_host, const int32_t *nterms_host, int nq, int top, float *rel_dev, void *scratch,
            size_t scratch_bytes, int32_t *out_idx_host, float *out_rel_host, float *out_term_entropy_host,
            float *out_entropy_host, float *out_mass_host, cudaStream_t st);

}  // namespace sert
