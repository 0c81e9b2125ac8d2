// Running top-k state shared by the score-tile producers (fp32 FMA tile in score.cu, tcgen05 tile in
// gemm_tc.cu) and the prune / merge kernels.
#pragma once

#include <algorithm>

#include <cuda_bf16.h>

#include "common.cuh"

namespace sert {

enum ScoreMode { SCORE_FMA = 0, SCORE_TENSOR = 1 };

struct TopkState {
  float *entities = nullptr;            // (rows,d) float32, row-major (L2-normalised when asked)
  long long rows = 0;
  int d = 0;
  long long row_begin = 0;              // global id of local row 0 (row-sharded scoring)
  int max_queries = 0;
  int cap = 0;                          // candidate slots per query (power of two, >= 2*max_k)
  unsigned long long *cand = nullptr;   // (max_queries, cap) keys
  unsigned long long *tau = nullptr;    // (max_queries,) key of the current k-th best (0 = none yet)
  int *count = nullptr;                 // (max_queries,) used slots (may exceed cap after an overflow)
  int *overflow = nullptr;              // set when an append found its list full
  // tensor-core mode (tcgen05 bf16x3 GEMM): split operands, K-major, Kt = terms * padded d
  int mode = SCORE_TENSOR;
  int terms = 3;
  int kt = 0;
  __nv_bfloat16 *ent_split = nullptr;   // (rows, kt)
  __nv_bfloat16 *q_split = nullptr;     // (max_queries, kt)
  // coarse-then-exact sweep (score.cu: topk_sweep): the hi.hi term alone selects candidates, with a per-query score
  // margin that bounds its rounding error rigorously
  int coarse = 1;
  float *margin = nullptr;              // (max_queries,)
  float ent_norm_max = 0.f;             // largest L2 norm of an entity row
  float ent_err_max = 0.f;              // largest L2 norm of (row - bf16(row)): the rows' share of the coarse error
  float ent_err2_max = 0.f;             // ... of (row - hi - mid): the rows' share when the coarse GEMM takes two blocks
  // Blocks of the split operands the coarse GEMM multiplies: 1 = [hi].[hi]; 2 = [hi|hi].[hi|mid] = q_hi.(e_hi + e_mid),
  // which removes the entity rows' rounding from the error (the margin band halves).  Two blocks double the MMA work;
  // taken when one block is a K of 128 or less, where the epilogue paces the sweep and the extra MMAs are free.
  int coarse_blocks = 1;
  // seeded single-launch sweep (score.cu: topk_sweep): group maxima of a strided row sample, (max_queries, kSeedGroups)
  float *gmax = nullptr;
  int seeded = 1;                       // 0: always the multi-chunk sweep (tests, diagnostics)
  long long *stats = nullptr;           // host counters: [0] sweeps answered by the seeded path, [1] sweeps that fell back
};

constexpr int kSeedGroups = 512;        // most group maxima per query the threshold seed selects from

int launch_normalise_rows(const float *in, float *out, int64_t rows, int d, cudaStream_t st);
int topk_prepare(int cap);
int topk_sweep(const TopkState &s, const float *queries_dev, int Q, int k, int32_t *out_idx, float *out_score,
               cudaStream_t st, int *deferred = nullptr, bool allow_seeded = true);
// part p's lists start at idx + p * part_stride / score + p * part_stride (0 = Q * k: dense [part][q][k] arrays)
int launch_topk_merge(const int32_t *idx, const float *score, int parts, int Q, int k, int32_t *out_idx,
                      float *out_score, cudaStream_t st, size_t part_stride = 0);

}  // namespace sert
