// Launchers for the log-linear row kernels (see ll_kernels.cu).
#pragma once

#include <algorithm>

#include "common.cuh"

namespace sert {

int launch_ll_row_stats(const float *Z, int64_t rows, int E, int64_t ldz, float *rmax, float *rsum,
                        cudaStream_t st);
int launch_ll_softmax_inplace(float *Z, int64_t rows, int E, int64_t ldz, const float *rmax,
                              const float *rsum, cudaStream_t st);
int launch_ll_joint(const float *Z, const float *rmax, const float *rsum, float *S, int B, int W, int E,
                    int64_t ldz, int64_t lds, cudaStream_t st);

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
};
int launch_ll_instance(const LlInstanceArgs &a, cudaStream_t st);

// dZ[i,w,:] = p * (dp - sum_e dp*p), dp = ds/clip(p) on the unclipped entries; in place over Z.
int launch_ll_dz(float *Z, const float *rmax, const float *rsum, const float *DS, int B, int W, int E,
                 int64_t ldz, int64_t lds, cudaStream_t st);

}  // namespace sert
