// tcgen05 / TMEM / TMA GEMM for the two tensor-core-shaped ops of the SERT hot path:
//   * the query x all-entities scoring GEMM (bin/query.py:304-318 replaced; epilogue = running top-k filter)
//   * the word x entity projection GEMM of the log-linear model (sert/models.py:846-849) and its gradients.
// C[M,N] (f32) = A[M,K'] . B[N,K']^T with bf16 operands, both K-major ("TN"), fp32 accumulation in TMEM.
// fp32 accuracy comes from a 3-term bf16 split of each operand laid out along K (split_bf16 below):
//   A'' = [A_hi | A_hi | A_mid],  B'' = [B_hi | B_mid | B_hi]   =>  A''.B''^T = hi.hi + hi.mid + mid.hi
// (relative error ~2^-16 per product instead of 2^-8), so one plain bf16 GEMM of depth K' = 3K does it.
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"

namespace sert {

enum TcEpilogueMode { TC_EPI_STORE = 0, TC_EPI_TOPK = 1, TC_EPI_GROUPMAX = 2 };

struct TcEpilogue {
  int mode = TC_EPI_STORE;
  // TC_EPI_STORE: C[m, n] = acc (+ bias[n])
  float *C = nullptr;
  long long ldc = 0;
  const float *bias = nullptr;
  // TC_EPI_STORE, optional: row `extra_row` of C goes to extra_dst[n] instead of C (a row of ones appended to A turns
  // the GEMM's last output row into the column sums of B^T -- the bias gradient of the log-linear layer for free)
  int extra_row = -1;
  float *extra_dst = nullptr;
  // TC_EPI_STORE: C += acc (float4 reductions) instead of C = acc; the launcher may then split a deep K range into
  // slices that are scheduled as independent tiles.  C must hold the initial value (zeros) before the launch.
  int accumulate = 0;
  // TC_EPI_STORE, optional: softmax statistics of the stored values (after the bias), per row and per 64-column
  // slice -- row_stats[m * stats_ld + slot] = (max, sum of exp(v - max)) with slot = (n - n_begin) / 64 -- so that the
  // row-softmax of the log-linear logits needs no pass of its own over the (B*W, E) matrix
  // (launch_ll_combine_slices folds the slices).  Slots of a row that lie beyond n_end are not written.
  float2 *row_stats = nullptr;
  int stats_ld = 0;
  // TC_EPI_TOPK: rows are queries, columns are entity rows [n_begin, n_end) of B
  const unsigned long long *tau = nullptr;
  int *count = nullptr;
  unsigned long long *cand = nullptr;
  int cap = 0;
  long long row_offset = 0;     // global id of B row 0
  int *overflow = nullptr;      // set to 1 when a candidate list is full
  // TC_EPI_GROUPMAX (threshold seeding of the scoring sweep, score.cu): every n-tile must be full (256 valid rows).
  // gmax[m * gmax_ld + tile * (256 / group) + c / group] = max over the `group` (8, 16, 32 or 64) columns around column c
  // of n-tile `tile`; nothing else is written.
  float *gmax = nullptr;
  int gmax_ld = 0;
  int group = 64;
  // every mode: n-tile t covers B rows [n_begin + t * 256 * tile_stride, +256): a strided sample of the rows when > 1
  int tile_stride = 1;
};

enum SplitRole { SPLIT_A = 0, SPLIT_B = 1 };

// K padded to a multiple of 64 (one 128-byte swizzle row of bf16); returns the padded K of ONE term.
static inline int tc_padded_k(int K) { return (int)align_up((size_t)K, 64); }

// dst (rows, terms * Kp) bf16 <- split of src (rows, K) f32 (row stride ld_src); terms in {1, 2, 3}
// (2 = [hi | mid] for launch_gemm_tc_pair, role ignored).
int launch_split_bf16(const float *src, long long rows, int K, long long ld_src, int terms, SplitRole role,
                      __nv_bfloat16 *dst, cudaStream_t st);

// Transposing variant: dst (C, terms * R64) bf16 <- src (R, C) f32, i.e. the K-major split of src^T.
int launch_split_bf16_t(const float *src, long long R, long long C, long long ld_src, int terms, SplitRole role,
                        __nv_bfloat16 *dst, cudaStream_t st);

// Runs the GEMM over B rows [n_begin, n_end) (columns of C).  A: (M, Kt) bf16, B: (N_total, Kt) bf16,
// Kt = terms * Kp a multiple of 64.  For TC_EPI_STORE column n of C is B row n.
int launch_gemm_tc(const __nv_bfloat16 *A, int M, const __nv_bfloat16 *B, long long N_total, long long n_begin,
                   long long n_end, int Kt, const TcEpilogue &epi, cudaStream_t st);
// Pair operands: A (M, 2 Kp) and B (N_total, 2 Kp) hold [hi | mid] (split terms = 2); the kernel forms
// hi.hi + hi.mid + mid.hi itself from two ring stages per 64-column block -- the same three products as the 3-term
// layout from two thirds of the operand traffic and storage.  Store epilogue only.
int launch_gemm_tc_pair(const __nv_bfloat16 *A, int M, const __nv_bfloat16 *B, long long N_total, long long n_begin,
                        long long n_end, int Kp, const TcEpilogue &epi, cudaStream_t st);
// Pair operands with an N-MAJOR B: Bn holds k_rows rows (the K index) of [hi | mid] blocks of n_pad columns (the N
// index; row stride 2 * n_pad) -- e.g. the rows of a matrix whose TRANSPOSE is the B operand.  The kernel loads
// 64 x 64 boxes and multiplies through an MN-major shared-memory descriptor, so no transposed copy is needed.
// A: (M, 2 Kp) K-major pair operand with Kp >= k_rows (columns beyond k_rows must be zero or meet zero rows of Bn:
// rows >= k_rows of Bn read as zeros).
int launch_gemm_tc_pair_bn(const __nv_bfloat16 *A, int M, const __nv_bfloat16 *Bn, long long k_rows, long long n_pad,
                           long long N_total, long long n_begin, long long n_end, int Kp, const TcEpilogue &epi,
                           cudaStream_t st);
// Same with explicit row strides (elements): the first Kt columns of wider operands, e.g. the hi.hi term alone of
// 3-term split operands (Kt = Kp, lda = ldb = 3 Kp).
int launch_gemm_tc_ld(const __nv_bfloat16 *A, long long lda, int M, const __nv_bfloat16 *B, long long ldb,
                      long long N_total, long long n_begin, long long n_end, int Kt, const TcEpilogue &epi,
                      cudaStream_t st);

}  // namespace sert
