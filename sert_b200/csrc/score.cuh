// Running top-k state shared by the score-tile producers (fp32 FMA tile in score.cu, tcgen05 tile in
// gemm_tc.cu) and the prune / merge kernels.
#pragma once

#include <algorithm>

#include "common.cuh"

namespace sert {

struct TopkState {
  float *entities = nullptr;            // (rows,d) float32, row-major (L2-normalised when asked)
  long long rows = 0;
  int d = 0;
  long long row_begin = 0;              // global id of local row 0 (row-sharded scoring)
  int max_queries = 0;
  int cap = 0;                          // candidate slots per query (power of two, >= 2*max_k)
  unsigned long long *cand = nullptr;   // (max_queries, cap) keys
  unsigned long long *tau = nullptr;    // (max_queries,) key of the current k-th best (0 = none yet)
  int *count = nullptr;                 // (max_queries,) used slots
};

int launch_normalise_rows(const float *in, float *out, int64_t rows, int d, cudaStream_t st);
int topk_prepare(int cap);
int topk_sweep(const TopkState &s, const float *queries_dev, int Q, int k, int32_t *out_idx, float *out_score,
               cudaStream_t st);
int launch_topk_merge(const int32_t *idx, const float *score, int parts, int Q, int k, int32_t *out_idx,
                      float *out_score, cudaStream_t st);

}  // namespace sert
