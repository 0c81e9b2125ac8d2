// Launcher declarations for the sm_100a kernels of libsert_b200.
// Each launcher enqueues on `st` and returns 0 / -1 (error text via sert::set_error).
#pragma once

#include "common.cuh"

namespace sert {

// ---- embedding gather (sert/models.py:180) fused with the window mean (sert/models.py:226,1051) ---
// out[i,:] = (sum_w R[x[i,w],:]) / denom        (pooled, (B,d))
int launch_gather_pool(const int32_t *x, const float *R, float *out, int B, int W, int d, float denom,
                       cudaStream_t st);
// out[i*W+w,:] = R[x[i,w],:]                      (unpooled rows, (B*W,d); log-linear)
int launch_gather_rows(const int32_t *x, const float *R, float *out, int64_t rows, int d, cudaStream_t st);

// ---- small fp32 GEMMs on CUDA cores (vector-space projection, sert/models.py:1057-1061, and its grads) --
enum GemmEpilogue { EPI_STORE = 0, EPI_BIAS_TANH = 1, EPI_ATOMIC_ADD = 2, EPI_BIAS = 3 };
// C[M,N] = epi(op(A)[M,K] * op(B)[K,N]); a_t: A stored (K,M) row-major; b_t: B stored (N,K) row-major.
// split_k > 1 requires EPI_ATOMIC_ADD.
int launch_gemm_f32(const float *A, const float *Bm, float *C, int M, int N, int K, bool a_t, bool b_t,
                    int lda, int ldb, int ldc, GemmEpilogue epi, const float *bias, int split_k,
                    cudaStream_t st);
// out[n] += sum_m A[m,n]   (bias gradient)
int launch_colsum_atomic(const float *A, float *out, int M, int N, cudaStream_t st);

// ---- vector-space negative-sampling loss, forward + backward (sert/models.py:1072-1098,893-902) ---
// Table shards (sert_model_set_table_shard_comm): every rank computes the whole batch, but accumulates gradient rows
// (and stamps them touched) only for the rows of the two tables it updates itself.  Default: everything.
struct RowOwner {
  int e_lo = 0, e_hi = 0x7fffffff;    // entity rows [e_lo, e_hi)
  int r_lo = 0, r_hi = 0x7fffffff;    // word rows
  __host__ __device__ bool entity(int row) const { return row >= e_lo && row < e_hi; }
  __host__ __device__ bool word(int row) const { return row >= r_lo && row < r_hi; }
};

constexpr int kMaxOwners = 8;       // ranks of a table-shard group (kMaxPeers + 1, one NVSwitch domain)
struct VsNceArgs {
  const float *t;        // (B,de) tanh(h.Wp+bp), unclipped
  const float *Eemb;     // (E,de)
  const int32_t *y;      // (B,)
  const int32_t *neg;    // (B,k)
  const float *w;        // (B,) instance weights or nullptr (=1)
  float *gE;             // (E,de) gradient accumulator (train) or nullptr
  uint32_t *flagE;       // (E,) touched stamps
  uint32_t stamp;
  float *da;             // (B,de) d loss / d pre-activation (train)
  double *loss_acc;      // += sum_i w_i*ell_i (train) or sum_i ell_i (eval)
  float *dbg_scores;     // (B,1+k) logits or nullptr
  float *dbg_u;          // (B,de) clipped projection or nullptr
  float *dbg_ell;        // (B,) or nullptr
  int B, k, de;
  float inv_B;
  bool train;
  RowOwner own;
};
int launch_vs_nce(const VsNceArgs &a, cudaStream_t st);

// fused gather -> tanh projection -> loss fwd/bwd -> back-projection -> scatter, one warp per pair of instances
// (csrc/vs_warp.cu).  Returns 0 = launched, 1 = shape not supported (use the per-stage kernels), -1 = error.
struct VsFusedArgs {
  const int32_t *x;       // (B,W)
  const float *R;         // (V,dw)
  const float *Wp;        // (dw,de)
  const float *bp;        // (de,)
  const float *Eemb;      // (E,de)
  const int32_t *y;       // (B,)
  const int32_t *neg;     // (B,k)
  const float *w;         // (B,) or nullptr
  float *gE; uint32_t *flagE;
  float *gR; uint32_t *flagR;
  uint32_t stamp;
  float *h;               // (B,dw) out (for dW = h^T.da)
  float *da;              // (B,de) out
  double *loss_acc;
  int B, W, k, dw, de;
  float inv_B;
  // Hot word rows (tile kernel only; nullptr / 0 = none).  A word id that occurs thousands of times per batch makes
  // thousands of row additions land on the same four L2 lines, which the L2 slices serialise (tools/red_probe.cu:
  // 44 MB of row additions take 14-20 us on uniform rows and 39-44 us on Zipf rows; plain stores behave the same).
  // Additions to a hot row go to one of `hot_replicas` private copies instead (chosen by CTA), and
  // launch_hot_update applies their sum (the dense update skips rows flagged kHotRowMark).
  const int8_t *hot_slot = nullptr;   // (V,) slot of a hot word id, -1 otherwise
  float *hot_acc = nullptr;           // (hot_replicas, kMaxHotRows, dw) zero between steps
  int hot_replicas = 0;               // power of two
  // Loss of the PREVIOUS step, finalised by one extra CTA of the tile kernel (its accumulators are complete by
  // then and nothing else on the step's critical path needs them): nullptr = nothing pending.
  double *fin_acc = nullptr;
  float *fin_loss = nullptr;
  float fin_inv_B = 0.f, fin_reg_coeff = 0.f;
  RowOwner own;
  // Instance shards (table-shard mode 3, tile kernel only): this rank runs the instances [i0, B) of the batch -- h and
  // da rows are written at i - i0 -- and adds every gradient row (and its touched stamp) straight into the arena of
  // the rank that updates it, over NVLink for the other ranks' rows: rank o owns the entity rows
  // [e_bound[o], e_bound[o + 1]) and the word rows [r_bound[o], r_bound[o + 1]).  n_owner == 0: gE / gR / flagE / flagR
  // above are the only destination.  Row ids must stay below 2^28 (the owner travels in the bits above).
  int i0 = 0;
  int n_owner = 0;
  int e_bound[kMaxOwners + 1] = {}, r_bound[kMaxOwners + 1] = {};
  float *gE_peer[kMaxOwners] = {}, *gR_peer[kMaxOwners] = {};
  uint32_t *flagE_peer[kMaxOwners] = {}, *flagR_peer[kMaxOwners] = {};
};
constexpr int kMaxHotRows = 32;
constexpr int kHotReplicas = 16;
constexpr uint32_t kHotRowMark = 0xffffffffu;   // flag value of a hot row: the dense update leaves the row alone
// true when launch_vs_tile serves this shape
bool vs_tile_supported(int dw, int de, int W, int k);
// WpT_scratch: (de, dw) floats, the transpose of Wp; refresh_WpT = recompute it first (it is stale).
// variant: 1 = CTA-per-8-instances tile kernel (csrc/vs_tile.cu) when d_w = d_e = 128, window <= 32, k <= 15, else the
// warp kernel; 2 = warp kernel only.
int launch_vs_fused(const VsFusedArgs &a, float *WpT_scratch, bool refresh_WpT, int variant, cudaStream_t st);
// the tile kernel alone: 0 = launched, 1 = shape not served, -1 = error.  WpT: (de, dw) transpose of Wp.
int launch_vs_tile(const VsFusedArgs &a, const float *WpT, cudaStream_t st);

// scatter-add of dh/denom into the word-gradient rows (autodiff of the gather, AdvancedIncSubtensor)
int launch_scatter_rows(const int32_t *x, const float *dh, float *gR, uint32_t *flagR, uint32_t stamp,
                        int B, int W, int d, float denom, cudaStream_t st, int row_lo = 0, int row_hi = 0x7fffffff);
// Table shards with look-ahead: need_e / need_r [row] = stamp for every row the NEXT batch (x, y, neg) will read, so
// that the update kernels send only those rows' new values to the other ranks
int launch_mark_needed(const int32_t *x, const int32_t *y, const int32_t *neg, int B, int W, int k, uint32_t *need_r,
                       uint32_t *need_e, uint32_t stamp, cudaStream_t st);

// Instance shards: need_*[row] |= 1 << r for every row that rank r's instances [i_bound[r], i_bound[r + 1]) of the NEXT
// batch read (the arrays are zeroed first); the update kernels send a row only to the ranks whose bit is set
int launch_mark_needed_by(const int32_t *x, const int32_t *y, const int32_t *neg, int B, int W, int k, uint32_t *need_r,
                          uint32_t *need_e, const int *i_bound, int n_ranks, cudaStream_t st);

// uniform negatives with replacement over [0,E) (sert/models.py:956-973); Philox4x32-10
int launch_sample_negatives(int32_t *out, int64_t n, int64_t E, uint64_t seed, uint64_t step,
                            cudaStream_t st);

// ---- dense optimisers with L2 (lasagne.updates.adam / adadelta; sert/models.py:764-795,820,922) -----
struct ParamSegment {
  long long offset;      // float offset inside the arena arrays
  long long count;       // number of floats (multiple of 4, zero padded)
  int row_len;           // floats per row for flag lookup
  int regularised;       // 0: no L2 (bias); 1: L2 gradient + counted in the reported loss; 2: L2 gradient only
                         // (replicated tensor of an entity-sharded model: rank 0 alone reports its norm)
  const uint32_t *flags; // per-row touched stamps or nullptr (= always read the gradient)
};
constexpr int kMaxSegments = 4;
constexpr int kMaxPeers = 7;        // other ranks of a table-shard group (one NVSwitch domain: 8 GPUs)
constexpr int kSumsqSlots = 64;     // acc[0] = data-loss sum, acc[1..64] = partial sums of theta^2
struct OptimArgs {
  float *theta, *s1, *s2, *grad;   // arena arrays of `total` floats
  long long total;                 // multiple of 4
  ParamSegment seg[kMaxSegments];
  int num_segments;
  uint32_t stamp;
  float l2_scale;                  // lambda / B
  // Adam: c0 = a_t, c1 = beta1, c2 = beta2, c3 = eps.  Adadelta: c0 = lr, c1 = rho, c3 = eps.
  float c0, c1, c2, c3;
  // loss finalisation by the last block: loss[slot] = data_acc*inv_B + reg_coeff * sum(theta^2)
  double *acc;                     // 1 + kSumsqSlots doubles, see above
  float *loss_out;
  float inv_B, reg_coeff;
  bool no_finalize = false;        // phase 0: leave the loss to a later launch_finalize_train
  int phase = 0;                   // 0 = everything, 3 = row-stamped tables only, 4 = dense tensors only
  long long first4 = 0;            // first 16-byte chunk to process (phase 4 starts at the first dense tensor)
  unsigned int *ticket = nullptr;  // phase 4: block counter (zero between launches); the last block finalises the
                                   // loss.  nullptr: the caller finalises later (launch_finalize_train / the next
                                   // step's tile kernel)
  // phase 4, optional: segment `transposed_segment` ((rows, row_len) row-major) is also written transposed here
  float *transposed = nullptr;
  int transposed_segment = 0, transposed_rows = 0;
  // 1: s1 / s2 are arrays of bfloat16 (sert_config.dtype_mode 1), written with stochastic rounding -- 16 instead of
  // 24 bytes per parameter and step.  Element offsets are the same as for theta.
  int state_bf16 = 0;
  // Table shards (vs_train_step with sert_model_set_table_shard_comm): the launch covers the 16-byte chunks
  // [first4, last4) only (this rank's piece of the tables; last4 < 0: to the end).  The new theta goes to theta_out
  // (nullptr: in place) and to the same offset of every peer_theta[p] -- the other ranks' copies, written over NVLink
  // by this kernel's own stores, so the update IS the exchange.  With look-ahead (push_all == 0) a chunk of a
  // segment with need flags is sent only when its row is marked need_stamp (the next batch reads it).
  long long last4 = -1;
  float *theta_out = nullptr;
  int n_peers = 0;
  float *peer_theta[kMaxPeers] = {};
  const uint32_t *need[kMaxSegments] = {};
  uint32_t need_stamp = 0;
  int push_all = 1;
  // instance shards: the need words are bit masks over the ranks (launch_mark_needed_by); peer_rank[p] = rank behind
  // peer_theta[p]
  int need_is_mask = 0;
  int peer_rank[kMaxPeers] = {};
};

// Instance shards: the gradient of hot word row s gathered on this rank (sum of its private copies, zeroed here) is
// added into the gradient row of the rank that updates it (grad_peer[o] = that rank's gradient arena, r_bound as in
// VsFusedArgs); launch_hot_update on the owner then sees the sum over all ranks in the row itself.
struct HotPushArgs {
  float *hot_acc;                  // (kHotReplicas, kMaxHotRows, d)
  const int32_t *hot_ids;
  int n_hot, d;
  long long table_offset;          // float offset of the word table inside the gradient arenas
  int n_owner;
  int r_bound[kMaxPeers + 2];
  float *grad_peer[kMaxPeers + 1];
};
int launch_hot_push(const HotPushArgs &h, cudaStream_t st);
// dst[0 .. n) += src[0 .. n); src <- 0 (n a multiple of 4, 16-byte aligned): a rank's share of the dense tensors'
// gradient, accumulated locally, goes to the rank that updates them in one pass of vector reductions over NVLink
int launch_push_add(float *src, float *dst, long long n, cudaStream_t st);

// ---- barrier + accumulator exchange of a table-shard group over NVLink peer memory (csrc/peer_sync.cu) ----------
constexpr int kAccSlots = 1 + kSumsqSlots;     // acc[0] = data-loss sum, acc[1..64] = partial sums of theta^2
struct PeerSyncBlock {
  uint32_t flag[kMaxPeers + 1][32];            // flag[r][0]: the last epoch rank r has reached (one 128-byte line each)
  double inbox[2][kMaxPeers + 1][72];          // [epoch parity][writer rank][accumulator]
};
struct PeerBarrierArgs {
  PeerSyncBlock *blk[kMaxPeers + 1] = {};      // rank r's block as mapped in this process (blk[rank] = the local one)
  int rank = 0, world = 1;
  uint32_t epoch = 0;                          // 1, 2, 3, ... (the same sequence on every rank)
  double *acc = nullptr;                       // kAccSlots local accumulators, replaced by the sum over the ranks
  int acc_first = 0;                           // accumulators below this index are left alone
  unsigned int *error = nullptr;               // page-locked host word, set when a peer never arrives
};
int launch_peer_barrier(const PeerBarrierArgs &a, cudaStream_t st);

// Adam + L2 of the hot word rows (gradient = sum of the private copies), see opt_kernels.cu
struct HotUpdateArgs {
  float *theta, *s1, *s2, *grad;   // arena arrays
  long long table_offset;          // float offset of the word table inside them
  int d;                           // floats per row
  float *hot_acc;                  // (kHotReplicas, kMaxHotRows, d)
  const int32_t *hot_ids;
  int n_hot;
  float l2_scale, c0, c1, c2, c3;
  double *acc;                     // sum(theta^2) goes to acc[1 + slot]
  int counted;                     // 1: this rank reports the table's norm in the loss
  int state_bf16 = 0;              // as OptimArgs::state_bf16
  uint32_t stamp = 0;              // seeds the stochastic rounding
  // table shards: hot rows outside the chunks [own_lo4, own_hi4) belong to another rank (nothing to do here)
  long long own_lo4 = 0, own_hi4 = 0x7fffffffffffffffll;
  float *theta_out = nullptr;
  int n_peers = 0;
  float *peer_theta[kMaxPeers] = {};
};
int launch_hot_update(const HotUpdateArgs &h, cudaStream_t st);
// flags[hot_ids[s]] = value  (kHotRowMark to hand the rows to launch_hot_update, 0 to hand them back)
int launch_hot_mark(uint32_t *flags, const int32_t *hot_ids, int n_hot, uint32_t value, cudaStream_t st);
// train loss finalisation as a kernel of its own: loss_out = acc[0]*inv_B + reg_coeff*sum(acc[1..64]); acc <- 0
int launch_finalize_train(double *acc, float *loss_out, float inv_B, float reg_coeff, cudaStream_t st);
int launch_adam(const OptimArgs &a, cudaStream_t st);
int launch_adadelta(const OptimArgs &a, cudaStream_t st);
// eval loss finalisation: loss_out = acc[0]*inv_B ; acc[0] = 0
int launch_finalize_eval(double *acc, float *loss_out, float inv_B, cudaStream_t st);

}  // namespace sert
