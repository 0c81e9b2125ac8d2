// extern "C" surface of libsert_b200.so: model life cycle, device-resident data set, the
// train / eval / predict sequences of both models.  See include/sert_b200.h for the contract and the
// reference interfaces each entry point replaces.
#include <math.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "comm.cuh"
#include "kernels.cuh"
#include "gemm_tc.cuh"
#include "ll_kernels.cuh"

namespace sert {

static thread_local std::string g_error;
static std::atomic<uint64_t> g_launches{0};
void set_error(const std::string &msg) { g_error = msg; }
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool first_use_on_device(std::atomic<uint64_t> &mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  const uint64_t bit = 1ull << (dev & 63);
  return (mask.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}

// ---- bump allocator over the caller-provided HBM arena ------------------------------------------
struct Bump {
  char *base;
  size_t off = 0;
  explicit Bump(void *b) : base(static_cast<char *>(b)) {}
  template <typename T>
  T *take(size_t count) {
    off = align_up(off, 256);
    T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

struct Dataset {
  int64_t n = 0;
  const int32_t *x = nullptr;
  const int32_t *y = nullptr;
  const int64_t *indptr = nullptr;
  const int32_t *indices = nullptr;
  const float *data = nullptr;
  const float *w = nullptr;
};

}  // namespace sert

using namespace sert;

// Blocks per split operand of the log-linear GEMMs: 2 = pair operands [hi | mid] (gemm_tc.cuh: launch_gemm_tc_pair,
// default), 3 = the [hi|hi|mid] x [hi|mid|hi] layout (SERT_LL_TERMS=3, for A/B runs).  Read once per process.
static int ll_terms() {
  static const int t = [] {
    const char *e = getenv("SERT_LL_TERMS");
    return e != nullptr && e[0] == '3' ? 3 : 2;
  }();
  return t;
}

struct sert_model {
  sert_config cfg;
  cudaStream_t st = nullptr;
  char *arena = nullptr;
  size_t arena_bytes = 0;
  // parameter arena: theta / state1 / state2 / grad share one layout
  float *theta = nullptr, *s1 = nullptr, *s2 = nullptr, *grad = nullptr;
  long long total = 0;
  long long off[4] = {0, 0, 0, 0};    // by SERT_PARAM_*: float offset
  long long cnt[4] = {0, 0, 0, 0};    // logical float count (0 = tensor absent)
  ParamSegment seg[kMaxSegments];
  int nseg = 0;
  uint32_t *flagR = nullptr, *flagE = nullptr;
  uint32_t stamp = 0;
  int64_t step = 0;                   // Adam's t
  uint64_t sample_calls = 0;
  double *acc = nullptr;
  float *losses = nullptr;
  Dataset ds[2];
  // vector-space workspaces
  float *h = nullptr, *t = nullptr, *da = nullptr, *dh = nullptr, *WpT = nullptr;
  // hot word rows of the fused tile kernel (kernels.cuh: VsFusedArgs::hot_slot)
  int8_t *hot_slot = nullptr;
  int32_t *hot_ids = nullptr;
  float *hot_acc = nullptr;
  int n_hot = 0;
  int32_t *neg = nullptr;
  float *dbg_scores = nullptr, *dbg_u = nullptr, *dbg_ell = nullptr;
  // log-linear workspaces
  float *X = nullptr, *Z = nullptr, *S = nullptr, *DS = nullptr, *dX = nullptr, *rmax = nullptr, *rsum = nullptr;
  float *lrsum = nullptr;          // log of rsum (joint pass in the log domain)
  float2 *zstats = nullptr;        // (B*W, ceil(E/64)) per-slice softmax statistics from the GEMM epilogue
  // bf16x3 split operands of the three word x entity GEMMs (tcgen05 path, csrc/gemm_tc.cu), all K-major:
  __nv_bfloat16 *Xs = nullptr;     // (B*W, 3*dw64)   A of  Z  = X . Wd
  __nv_bfloat16 *WdT_s = nullptr;  // (E,   3*dw64)   B of  Z  (transposed split of Wd (dw,E))
  __nv_bfloat16 *dZs = nullptr;    // (B*W, 3*E64)    A of  dX = dZ . Wd^T
  __nv_bfloat16 *Wd_s = nullptr;   // (dw,  3*E64)    B of  dX
  __nv_bfloat16 *XT_s = nullptr;   // (dw,  3*BW64)   A of  dWd = X^T . dZ
  __nv_bfloat16 *dZT_s = nullptr;  // (E,   3*BW64)   B of  dWd
  bool use_tensor = true;
  // entity-sharded (column-parallel) log-linear step: this rank owns the columns [e_begin, e_begin + cfg.entities)
  // of Wd / bd; R is replicated.  `exchange` performs the collectives on device buffers (include/sert_b200.h).
  int shard_rank = 0, shard_world = 1;
  int64_t e_begin = 0, e_total = 0;
  sert_exchange_fn exchange = nullptr;
  void *exchange_ctx = nullptr;
  sert_comm *comm = nullptr;       // set: the exchanges are NCCL collectives issued by the library itself (comm.cu)
  // table-sharded vector-space step (sert_model_set_table_shard_comm): this rank updates the 16-byte chunks
  // [table_lo4[rank], table_lo4[rank + 1]) of the two tables (the LAST rank also the dense tensors) for the whole group.
  // Mode 1: the updated pieces are broadcast in place by NCCL.  Mode 2: theta lives in two library-owned buffers that
  // every rank maps (CUDA IPC); the update kernels store the new values into the NEXT buffer of every rank over
  // NVLink, and the buffers swap behind the step's one all-reduce.
  sert_comm *table_comm = nullptr;
  int table_mode = 0;
  long long table_lo4[kMaxPeers + 2] = {};
  RowOwner table_own;              // this rank's rows of the two tables (pieces end on row boundaries)
  uint32_t *need_r = nullptr, *need_e = nullptr;   // (V,), (E,): rows the next batch reads (look-ahead, mode 2)
  int32_t *neg_alt = nullptr;      // negatives of the next step, drawn one step ahead (look-ahead with device sampling)
  bool neg_presampled = false;
  float *arena_theta = nullptr;    // theta's place in the arena (mode 2 moves m.theta out of it)
  float *pp[2] = {nullptr, nullptr};
  void *pp_peers[2][kMaxPeers + 1] = {};
  int pp_cur = 0;
  // barrier + sum(theta^2) exchange of the group over NVLink peer memory (csrc/peer_sync.cu); replaces the step's one
  // ncclAllReduce in mode 2 unless SERT_TABLE_SHARD_BARRIER=nccl
  PeerSyncBlock *sync_blk = nullptr;
  void *sync_peers[kMaxPeers + 1] = {};
  uint32_t sync_epoch = 0;
  unsigned int *sync_error = nullptr;   // page-locked host word
  // Mode 3 (instance shards): on top of mode 2, every rank runs the forward / backward of its own instances
  // [inst_bound[rank], inst_bound[rank + 1]) only and adds the gradient rows into the arena of the rank that updates
  // them: gradients and touched stamps live in one library-owned block [grad (total) | flagE (E) | flagR (V)] that all
  // ranks map (gsh_peers); m.grad / m.flagE / m.flagR point into it while the shard is attached.
  // SERT_TABLE_SHARD_TRACE=1: CUDA events at the phase boundaries of the sharded step on the model's stream; the mean
  // phase durations are printed to stderr when the shard is released (diagnostic)
  std::vector<std::vector<cudaEvent_t>> trace;
  float *gsh = nullptr;
  void *gsh_peers[kMaxPeers + 1] = {};
  float *arena_grad = nullptr;
  uint32_t *arena_flagE = nullptr, *arena_flagR = nullptr;
  int inst_bound[kMaxPeers + 2] = {};
  int e_bound[kMaxPeers + 2] = {}, r_bound[kMaxPeers + 2] = {};
  void *rank_scratch = nullptr;    // sert_ll_rank_queries: sort keys and per-query sums (library-owned, grown on demand)
  size_t rank_scratch_bytes = 0;
  float2 *sstats = nullptr;        // (B, ll_joint_slots(E)) softmax statistics of S per slot, left by the joint pass
  float *xstats = nullptr;         // [kMaxShards][2][B*W] gathered (row max, row sum)
  float *smax = nullptr, *ssum = nullptr, *adot = nullptr, *racc = nullptr;
  // host-batch staging
  int32_t *stage_x = nullptr, *stage_y = nullptr, *stage_neg = nullptr, *stage_indices = nullptr;
  int64_t *stage_indptr = nullptr;
  float *stage_w = nullptr, *stage_data = nullptr;
  size_t stage_nnz_cap = 0;
  // optional per-kernel timing of the dense update (bench.py's roofline leg)
  int use_fused = 1;                  // vector-space step: 0 = per-stage kernels, 1 = fused (tile kernel, else warp kernel), 2 = fused warp kernel
  // second stream + fork/join events: the two small dense-gradient kernels overlap the table update
  bool overlap = true;
  bool wpt_valid = false;             // WpT holds the transpose of the current projection matrix
  // lazy vector-space steps (vs_train_step): accumulator bank and loss slot of the step whose loss is not written yet
  int pending_bank = -1;
  float *pending_loss = nullptr;
  bool hot_marked = false;            // flagR of the hot word rows carries kHotRowMark
  // pipelined host batches (sert_train_batch_host_async): copy streams, second staging set, loss ring
  static constexpr int kPipe = 8;
  bool pipe_ready = false;
  cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  bool ev_free_set[2] = {false, false};
  cudaEvent_t ev_fused = nullptr;     // recorded behind the forward/backward kernels of a step (want_fused_event)
  bool want_fused_event = false;
  cudaEvent_t ev_loss[kPipe] = {};
  int32_t *stage2_x = nullptr, *stage2_y = nullptr, *stage2_neg = nullptr;
  float *stage2_w = nullptr;
  float *pipe_losses = nullptr;       // device, kPipe slots
  float *pin_loss = nullptr;          // pinned host, kPipe slots
  int64_t host_steps = 0;             // tickets issued so far
  int64_t host_unfetched = -1;        // ticket whose loss is not yet on its way to the host
  cudaStream_t st2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool profile = false;
  double prof_bytes = 0.0;            // algorithmic bytes of the launches timed in profile mode
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
};

namespace sert {

static bool is_vs(const sert_config &c) { return c.kind == SERT_KIND_VECTORSPACE; }
// loss accumulators: two banks of (data-loss sum, kSumsqSlots partial sums of theta^2), 8 spare doubles (ticket)
constexpr int kAccBank = 1 + kSumsqSlots + 7;
constexpr int kAccDoubles = 2 * kAccBank + 8;
static double *acc_bank(sert_model &m, int bank) { return m.acc + (size_t)bank * kAccBank; }
constexpr int kMaxShards = 16;

static int validate(const sert_config &c) {
  SERT_REQUIRE(c.kind == SERT_KIND_LOGLINEAR || c.kind == SERT_KIND_VECTORSPACE, "unknown model kind");
  SERT_REQUIRE(c.batch > 0, "batch_size must be positive");                     // sert/models.py:315
  SERT_REQUIRE(c.window >= 1, "window_size must be >= 1");                      // sert/models.py:696
  SERT_REQUIRE(c.vocab > 0 && c.vocab < (1ll << 31), "vocabulary size out of range");
  SERT_REQUIRE(c.entities > 1 && c.entities < (1ll << 31), "number of entities out of range");  // bin/train.py:100
  SERT_REQUIRE(c.word_dim > 0 && c.word_dim % 4 == 0, "word representation size must be a positive multiple of 4");
  SERT_REQUIRE(c.loss_slots > 0, "loss_slots must be positive");
  SERT_REQUIRE(c.lambda >= 0.f, "regularization lambda must be >= 0");
  SERT_REQUIRE(c.dtype_mode == 0 || c.dtype_mode == 1, "dtype_mode must be 0 (float32) or 1 (bfloat16 optimiser state)");
  SERT_REQUIRE(c.reserved1 == 0, "reserved config fields must be zero");
  if (is_vs(c)) {
    SERT_REQUIRE(c.entity_dim > 0 && c.entity_dim % 4 == 0,
                 "entity representation size must be a positive multiple of 4");
    SERT_REQUIRE(c.num_negatives >= 1, "num_negative_samples must be positive");  // sert/models.py:948
  }
  SERT_REQUIRE((long long)c.vocab * c.word_dim < (1ll << 32), "word table above 2^32 elements");
  SERT_REQUIRE((long long)c.entities * (is_vs(c) ? c.entity_dim : c.word_dim) < (1ll << 32),
               "entity table above 2^32 elements");
  return 0;
}

// Lays the model out in the arena (base == nullptr: size query only).
static size_t carve(sert_model &m, void *base) {
  const sert_config &c = m.cfg;
  Bump b(base);
  const long long B = c.batch, W = c.window, V = c.vocab, E = c.entities, dw = c.word_dim;
  const long long de = is_vs(c) ? c.entity_dim : 0;
  // ---- parameter layout: the optimiser's parameter order, sert/models.py:542-543,1105 ----
  long long o = 0;
  m.nseg = 0;
  auto add = [&](int which, long long count, int row_len, int regularised) {
    m.off[which] = o;
    m.cnt[which] = count;
    ParamSegment &s = m.seg[m.nseg++];
    s.offset = o;
    s.count = (long long)align_up((size_t)count, 4);
    s.row_len = row_len;
    s.regularised = regularised;
    s.flags = nullptr;
    o += (long long)align_up((size_t)count, 64);   // 256-byte aligned tensors
  };
  if (is_vs(c)) {
    add(SERT_PARAM_ENTITY_REPR, E * de, (int)de, 1);
    add(SERT_PARAM_WORD_REPR, V * dw, (int)dw, 1);
    add(SERT_PARAM_DENSE_W, dw * de, (int)de, 1);
    add(SERT_PARAM_DENSE_B, de, (int)de, 0);
  } else {
    add(SERT_PARAM_WORD_REPR, V * dw, (int)dw, 1);
    add(SERT_PARAM_DENSE_W, dw * E, (int)E, 1);
    add(SERT_PARAM_DENSE_B, E, (int)E, 0);
  }
  m.total = o;
  const bool train = c.inference_only == 0;
  m.theta = b.take<float>(o);
  if (c.dtype_mode == 1) {      // bfloat16 optimiser state: same element offsets, half the bytes
    m.s1 = train ? reinterpret_cast<float *>(b.take<uint16_t>(o)) : nullptr;
    m.s2 = train ? reinterpret_cast<float *>(b.take<uint16_t>(o)) : nullptr;
  } else {
    m.s1 = train ? b.take<float>(o) : nullptr;
    m.s2 = train ? b.take<float>(o) : nullptr;
  }
  m.grad = train ? b.take<float>(o) : nullptr;
  m.flagR = train ? b.take<uint32_t>(V) : nullptr;
  m.flagE = (train && is_vs(c)) ? b.take<uint32_t>(E) : nullptr;
  m.acc = b.take<double>(kAccDoubles);
  m.losses = b.take<float>(c.loss_slots + 1);   // last slot: scratch for parity hooks
  if (is_vs(c)) {
    const long long k = c.num_negatives;
    m.h = b.take<float>(B * dw);
    m.t = b.take<float>(B * de);
    m.da = train ? b.take<float>(B * de) : nullptr;
    m.dh = train ? b.take<float>(B * dw) : nullptr;
    m.WpT = train ? b.take<float>(dw * de) : nullptr;
    m.hot_slot = train ? b.take<int8_t>(V) : nullptr;
    m.hot_ids = train ? b.take<int32_t>(kMaxHotRows) : nullptr;
    m.hot_acc = train ? b.take<float>((long long)kHotReplicas * kMaxHotRows * dw) : nullptr;
    m.neg = b.take<int32_t>(B * k);
    m.dbg_scores = b.take<float>(B * (k + 1));
    m.dbg_u = b.take<float>(B * de);
    m.dbg_ell = b.take<float>(B);
    m.stage_neg = b.take<int32_t>(B * k);
    m.stage_y = b.take<int32_t>(B);
    if (train) {
      m.stage2_x = b.take<int32_t>(B * W);
      m.stage2_y = b.take<int32_t>(B);
      m.stage2_neg = b.take<int32_t>(B * k);
      m.stage2_w = b.take<float>(B);
      m.pipe_losses = b.take<float>(sert_model::kPipe);
    }
  } else {
    m.X = b.take<float>(B * W * dw);
    m.Z = b.take<float>(B * W * E);
    m.S = b.take<float>(B * E);
    m.DS = train ? b.take<float>(B * E) : nullptr;
    m.dX = train ? b.take<float>(B * W * dw) : nullptr;
    m.rmax = b.take<float>(B * W);
    m.rsum = b.take<float>(B * W);
    m.lrsum = b.take<float>(B * W);
    m.zstats = b.take<float2>(B * W * ((E + 63) / 64));
    m.sstats = b.take<float2>(B * ll_joint_slots(E));
    m.xstats = b.take<float>(kMaxShards * 2 * B * W);
    m.smax = b.take<float>(B);
    m.ssum = b.take<float>(B);
    m.adot = b.take<float>(B);
    m.racc = b.take<float>(B * W);
    const long long dw64 = tc_padded_k((int)dw), E64 = tc_padded_k((int)E), BW64 = tc_padded_k((int)(B * W));
    const long long T = ll_terms();      // blocks per split operand
    m.Xs = b.take<__nv_bfloat16>(B * W * T * dw64);
    m.WdT_s = b.take<__nv_bfloat16>(E * T * dw64);
    if (train) {
      m.dZs = b.take<__nv_bfloat16>(B * W * T * E64);
      m.Wd_s = b.take<__nv_bfloat16>(dw * T * E64);
      m.XT_s = b.take<__nv_bfloat16>((dw + 1) * T * BW64);    // + a row of ones: the bias gradient falls out of gWd's GEMM
      m.dZT_s = b.take<__nv_bfloat16>(E * T * BW64);
    }
    m.dbg_ell = b.take<float>(B);
    m.stage_indptr = b.take<int64_t>(B + 1);
    m.stage_nnz_cap = (size_t)B * 64;            // host-streamed batches: up to 64 labels per row on average
    m.stage_indices = b.take<int32_t>(m.stage_nnz_cap);
    m.stage_data = b.take<float>(m.stage_nnz_cap);
  }
  m.stage_x = b.take<int32_t>(B * W);
  m.stage_w = b.take<float>(B);
  // flags are looked up through the segment table
  for (int s = 0; s < m.nseg; ++s) {
    if (m.seg[s].offset == m.off[SERT_PARAM_WORD_REPR]) m.seg[s].flags = m.flagR;
    if (is_vs(c) && m.seg[s].offset == m.off[SERT_PARAM_ENTITY_REPR]) m.seg[s].flags = m.flagE;
  }
  return align_up(b.off, 256);
}

static float adam_alpha_f32(int64_t t) {
  // a_t = lr*sqrt(1-beta2^t)/(1-beta1^t) evaluated in float32 (t is a floatX scalar in the graph)
  const float tf = (float)t;
  const float b1 = 0.9f, b2 = 0.999f, lr = 1e-3f;
  return lr * sqrtf(1.0f - powf(b2, tf)) / (1.0f - powf(b1, tf));
}

static OptimArgs optim_args(sert_model &m, float *loss_out) {
  OptimArgs a;
  a.theta = m.theta; a.s1 = m.s1; a.s2 = m.s2; a.grad = m.grad;
  a.state_bf16 = m.cfg.dtype_mode == 1 ? 1 : 0;
  a.total = m.total;
  for (int s = 0; s < m.nseg; ++s) a.seg[s] = m.seg[s];
  for (int s = m.nseg; s < kMaxSegments; ++s) a.seg[s] = ParamSegment{0, 0, 1, 0, nullptr};
  a.num_segments = m.nseg;
  a.stamp = m.stamp;
  const float B = (float)m.cfg.batch;
  a.l2_scale = m.cfg.lambda > 0.f ? m.cfg.lambda / B : 0.f;
  a.acc = m.acc; a.loss_out = loss_out;
  a.inv_B = 1.0f / B;
  a.reg_coeff = m.cfg.lambda > 0.f ? m.cfg.lambda / (2.0f * B) : 0.f;
  if (m.table_comm != nullptr && m.table_mode >= 2) {
    const int r = m.table_comm->rank, next = m.pp_cur ^ 1;
    a.theta_out = m.pp[next];
    for (int p = 0; p < m.table_comm->world; ++p)
      if (p != r) { a.peer_rank[a.n_peers] = p; a.peer_theta[a.n_peers++] = static_cast<float *>(m.pp_peers[next][p]); }
    a.need_is_mask = m.table_mode == 3 ? 1 : 0;
    static const char *dbg = getenv("SERT_TABLE_SHARD_DEBUG");
    if (dbg && strchr(dbg, 'p')) a.n_peers = 0;
    for (int sg = 0; sg < m.nseg; ++sg)
      a.need[sg] = m.seg[sg].flags == m.flagR ? m.need_r : m.seg[sg].flags == m.flagE ? m.need_e : nullptr;
  }
  return a;
}

// stamp of the step after `stamp` (vs_train_step advances it the same way)
static uint32_t next_stamp(uint32_t stamp) { return stamp + 1 == kHotRowMark ? 1u : stamp + 1; }

constexpr int kTracePhases = 6;   // step start | tile kernel | side stream joined | barrier A | table update | barrier B
static bool trace_on() {
  static const char *e = getenv("SERT_TABLE_SHARD_TRACE");
  return e != nullptr && e[0] == '1';
}
static void trace_mark(sert_model &m, int phase) {
  if (!trace_on() || m.table_comm == nullptr || m.trace.size() > 400) return;
  if (phase == 0) m.trace.emplace_back();
  if (m.trace.empty()) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, m.st);
  m.trace.back().push_back(e);
}
static void trace_report(sert_model &m) {
  if (m.trace.empty()) return;
  cudaStreamSynchronize(m.st);
  double sum[kTracePhases] = {}, gap = 0.0;
  int n = 0, ngap = 0;
  for (size_t i = 20; i < m.trace.size(); ++i) {
    const auto &ev = m.trace[i];
    if ((int)ev.size() != kTracePhases) continue;
    for (int p = 1; p < kTracePhases; ++p) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev[p - 1], ev[p]) == cudaSuccess) sum[p] += ms;
    }
    if (i + 1 < m.trace.size() && (int)m.trace[i + 1].size() == kTracePhases) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev[kTracePhases - 1], m.trace[i + 1][0]) == cudaSuccess) { gap += ms; ++ngap; }
    }
    ++n;
  }
  if (n > 0)
    fprintf(stderr, "[table shard trace] rank %d mode %d steps %d: mark %.1f us | tile %.1f | side join %.1f | barrier A %.1f | "
            "table update %.1f | join + barrier B %.1f | gap to next step %.1f\n", m.table_comm->rank, m.table_mode, n,
            0.0, 1e3 * sum[1] / n, 1e3 * sum[2] / n, 1e3 * sum[3] / n, 1e3 * sum[4] / n, 1e3 * sum[5] / n,
            ngap ? 1e3 * gap / ngap : 0.0);
  for (auto &ev : m.trace)
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  m.trace.clear();
}

// One barrier of the table-shard group on the model's stream; acc[acc_first .. 64] become their sums over the ranks.
static int peer_barrier(sert_model &m, double *acc, int acc_first) {
  PeerBarrierArgs b;
  for (int r = 0; r < m.table_comm->world; ++r) b.blk[r] = static_cast<PeerSyncBlock *>(m.sync_peers[r]);
  b.rank = m.table_comm->rank; b.world = m.table_comm->world;
  b.epoch = ++m.sync_epoch;
  b.acc = acc; b.acc_first = acc_first;
  b.error = m.sync_error;
  return launch_peer_barrier(b, m.st);
}

// A peer that never reached a barrier (peer_sync.cu: kTimeoutNs) leaves the parameters of this rank undefined.
static int peer_barrier_check(sert_model &m) {
  if (m.sync_error != nullptr && *static_cast<volatile unsigned int *>(m.sync_error) != 0u) {
    set_error("table shards: a rank of the group did not reach the step's barrier (timeout); the model is invalid");
    return -1;
  }
  return 0;
}

// Table shards: what follows the update kernels of a step on the model's stream.  Mode 1 broadcasts every owner's
// piece of theta in place.  Both modes sum the owners' partial sum(theta^2) (the loss is finalised after this), and
// in mode 2 that all-reduce is also the step's barrier: it completes on a rank only after every rank has finished
// its update kernels, i.e. after every store into this rank's next buffer has landed; then the buffers swap.
static int table_exchange(sert_model &m, double *acc) {
  sert_comm *c = m.table_comm;
  if (m.table_mode == 1) {
    size_t off[kMaxPeers + 1], len[kMaxPeers + 1];
    for (int r = 0; r < c->world; ++r) {
      off[r] = (size_t)m.table_lo4[r] * 16;
      len[r] = (size_t)(m.table_lo4[r + 1] - m.table_lo4[r]) * 16;
    }
    if (comm_gather_pieces(c, m.theta, off, len, m.st)) return -1;
    const long long dense0 = m.off[SERT_PARAM_DENSE_W];
    if (comm_broadcast(c, m.theta + dense0, (size_t)(m.total - dense0) * sizeof(float), c->world - 1, m.st)) return -1;
  }
  // diagnostic (timing only, results invalid): SERT_TABLE_SHARD_DEBUG containing 'a' skips the all-reduce, 'p' the
  // peer stores, 'm' the look-ahead marks (every row is sent)
  static const char *dbg = getenv("SERT_TABLE_SHARD_DEBUG");
  if (dbg && strchr(dbg, 'a')) {
  } else if (m.sync_blk != nullptr) {
    if (peer_barrier(m, acc, m.table_mode == 3 ? 0 : 1)) return -1;   // instance shards: the data loss is partial too
  } else if (comm_all_reduce_sum_f64(c, acc + 1, kSumsqSlots, m.st)) {
    return -1;
  }
  if (m.table_mode >= 2) {
    m.pp_cur ^= 1;
    m.theta = m.pp[m.pp_cur];
  }
  return 0;
}

static int pick_split_k(int M, int N, int K) {
  const long long tiles = (long long)cdiv(M, 64) * cdiv(N, 64);
  long long want = (2ll * kNumSMs + tiles - 1) / tiles;
  const long long max_split = std::max(1, K / 64);
  if (want > max_split) want = max_split;
  if (want < 1) want = 1;
  return (int)want;
}

// dense update, optionally bracketed by CUDA events on the model's stream (profile mode)
static int timed_update(sert_model &m, const OptimArgs &o, bool adam) {
  if (!m.profile) return adam ? launch_adam(o, m.st) : launch_adadelta(o, m.st);
  // the events bracket the streaming kernel alone; the one-block loss finalisation follows outside them
  cudaEvent_t a, b;
  SERT_CUDA(cudaEventCreate(&a));
  SERT_CUDA(cudaEventCreate(&b));
  OptimArgs k = o;
  k.no_finalize = true;
  SERT_CUDA(cudaEventRecord(a, m.st));
  const int rc = adam ? launch_adam(k, m.st) : launch_adadelta(k, m.st);
  SERT_CUDA(cudaEventRecord(b, m.st));
  m.prof_events.emplace_back(a, b);
  // algorithmic bytes of this launch: theta and two state arrays read and written, float32
  long long params = 0;
  for (int sg = 0; sg < o.num_segments; ++sg)
    if (o.phase == 0 || (o.phase == 3) == (o.seg[sg].flags != nullptr)) params += o.seg[sg].count;
  m.prof_bytes = (o.state_bf16 ? 16.0 : 24.0) * (double)params;   // theta rw + two state arrays rw
  if (rc || o.phase != 0 || o.no_finalize) return rc;
  return launch_finalize_train(o.acc, o.loss_out, o.inv_B, o.reg_coeff, m.st);
}

// ---- vector space: one training step on device-resident batch pointers --------------------------
static int vs_forward(sert_model &m, const int32_t *x, cudaStream_t st) {
  const sert_config &c = m.cfg;
  float *R = m.theta + m.off[SERT_PARAM_WORD_REPR];
  float *Wp = m.theta + m.off[SERT_PARAM_DENSE_W];
  float *bp = m.theta + m.off[SERT_PARAM_DENSE_B];
  if (launch_gather_pool(x, R, m.h, c.batch, c.window, c.word_dim, (float)c.window, st)) return -1;
  return launch_gemm_f32(m.h, Wp, m.t, c.batch, c.entity_dim, c.word_dim, false, false, c.word_dim,
                         c.entity_dim, c.entity_dim, EPI_BIAS_TANH, bp, 1, st);
}

// Loss of a "lazy" training step (below) that no later tile kernel has finalised: one small kernel on the stream.
static int flush_pending(sert_model &m) {
  if (m.pending_bank < 0) return 0;
  const float B = (float)m.cfg.batch;
  const float reg = m.cfg.lambda > 0.f ? m.cfg.lambda / (2.0f * B) : 0.f;
  const int rc = launch_finalize_train(acc_bank(m, m.pending_bank), m.pending_loss, 1.0f / B, reg, m.st);
  m.pending_bank = -1;
  m.pending_loss = nullptr;
  return rc;
}

// Hot word rows belong to launch_hot_update only while the lazy tile path runs (their flags carry kHotRowMark).
static int set_hot_marks(sert_model &m, bool on) {
  if (m.hot_marked == on || m.n_hot == 0) return 0;
  m.hot_marked = on;
  return launch_hot_mark(m.flagR, m.hot_ids, m.n_hot, on ? kHotRowMark : 0u, m.st);
}

// The batch the NEXT vs_train_step call will be given (table shards with look-ahead: only the rows it reads are
// sent to the other ranks by this step's update).  x == nullptr: unknown, every row is sent.
struct NextBatch {
  const int32_t *x = nullptr, *y = nullptr;
  const int32_t *neg = nullptr;      // nullptr with sampled == true: its negatives are drawn now, one step ahead
  bool sampled = false;
};

static int vs_train_step(sert_model &m, const int32_t *x, const int32_t *y, const float *w,
                         const int32_t *neg, float *loss_out, const NextBatch &next = NextBatch()) {
  const sert_config &c = m.cfg;
  cudaStream_t st = m.st;
  const int B = c.batch, dw = c.word_dim, de = c.entity_dim;
  float *Wp = m.theta + m.off[SERT_PARAM_DENSE_W];
  const bool overlap = m.overlap && m.st2 != nullptr;
  // "Lazy" step = fused tile kernel + second stream.  Its critical path is two kernels, the tile kernel and the
  // Adam stream over the two tables; everything else -- gradients and update of the projection matrix and bias,
  // update of the hot word rows -- runs on the second stream under the table update, and the scalar loss is
  // written by the NEXT step's tile kernel (or by flush_pending when no step follows).
  const bool lazy = overlap && m.use_fused == 1 && m.WpT != nullptr && vs_tile_supported(dw, de, c.window, c.num_negatives);
  // instance shards (table-shard mode 3) exist on the lazy path only
  const bool inst = m.table_comm != nullptr && m.table_mode == 3;
  SERT_REQUIRE(!inst || lazy, "instance shards need the fused tile kernel and the second stream (sert_model_set_fused / _overlap)");
  if (!lazy && flush_pending(m)) return -1;
  if (set_hot_marks(m, lazy)) return -1;
  m.stamp += 1;
  if (m.stamp == kHotRowMark) m.stamp = 1;          // never collides in practice (2^32 steps); keeps the mark unique
  if (neg == nullptr) {
    if (m.neg_presampled) {                         // drawn by the previous step (look-ahead)
      std::swap(m.neg, m.neg_alt);
      m.neg_presampled = false;
    } else if (launch_sample_negatives(m.neg, (int64_t)B * c.num_negatives, c.entities, c.seed, m.sample_calls++, st)) {
      return -1;
    }
    neg = m.neg;
  }
  const bool sharded = m.table_comm != nullptr;
  // table shards: the LAST rank updates projection and bias -- their gradient GEMM, column sum and update are a chain
  // of small kernels beside the table update, and the hot word rows (the most frequent, lowest ids) sit in an earlier
  // rank's piece
  const bool dense_owner = !sharded || m.table_comm->rank == m.table_comm->world - 1;
  // table shards, look-ahead: mark the rows the next batch reads; the update kernels send only those
  bool push_all = true;
  static const char *ts_dbg = getenv("SERT_TABLE_SHARD_DEBUG");
  const bool traced = sharded && lazy;
  if (traced) trace_mark(m, 0);
  if (sharded && m.table_mode >= 2 && next.x != nullptr && m.table_comm->world > 1 && !(ts_dbg && strchr(ts_dbg, 'm'))) {
    const int32_t *next_neg = next.neg;
    if (next_neg == nullptr && next.sampled) {
      if (launch_sample_negatives(m.neg_alt, (int64_t)B * c.num_negatives, c.entities, c.seed, m.sample_calls++, st))
        return -1;
      m.neg_presampled = true;
      next_neg = m.neg_alt;
    }
    if (next_neg != nullptr) {
      if (inst) {
        // one block (V + E words): which ranks' instances of the next batch read each row
        SERT_CUDA(cudaMemsetAsync(m.need_r, 0, ((size_t)c.vocab + (size_t)c.entities) * sizeof(uint32_t), st));
        if (launch_mark_needed_by(next.x, next.y, next_neg, B, c.window, c.num_negatives, m.need_r, m.need_e,
                                  m.inst_bound, m.table_comm->world, st))
          return -1;
      } else if (launch_mark_needed(next.x, next.y, next_neg, B, c.window, c.num_negatives, m.need_r, m.need_e,
                                    next_stamp(m.stamp), st)) {
        return -1;
      }
      push_all = false;
    }
  }
  const int64_t t_next = m.step + 1;
  const int bank = lazy ? (m.pending_bank == 0 ? 1 : 0) : 0;
  double *acc = acc_bank(m, bank);
  VsFusedArgs f;
  f.x = x; f.R = m.theta + m.off[SERT_PARAM_WORD_REPR]; f.Wp = Wp; f.bp = m.theta + m.off[SERT_PARAM_DENSE_B];
  f.Eemb = m.theta + m.off[SERT_PARAM_ENTITY_REPR]; f.y = y; f.neg = neg; f.w = w;
  f.gE = m.grad + m.off[SERT_PARAM_ENTITY_REPR]; f.flagE = m.flagE;
  f.gR = m.grad + m.off[SERT_PARAM_WORD_REPR]; f.flagR = m.flagR; f.stamp = m.stamp;
  f.h = m.h; f.da = m.da; f.loss_acc = acc;
  f.B = B; f.W = c.window; f.k = c.num_negatives; f.dw = dw; f.de = de; f.inv_B = 1.0f / (float)B;
  f.own = m.table_own;
  if (inst) {
    const int world = m.table_comm->world;
    f.i0 = m.inst_bound[m.table_comm->rank];
    f.B = m.inst_bound[m.table_comm->rank + 1];
    f.n_owner = world;
    for (int r = 0; r <= world; ++r) { f.e_bound[r] = m.e_bound[r]; f.r_bound[r] = m.r_bound[r]; }
    for (int r = 0; r < world; ++r) {
      float *g = static_cast<float *>(m.gsh_peers[r]);
      uint32_t *fl = reinterpret_cast<uint32_t *>(g + m.total);
      f.gE_peer[r] = g + m.off[SERT_PARAM_ENTITY_REPR];
      f.gR_peer[r] = g + m.off[SERT_PARAM_WORD_REPR];
      f.flagE_peer[r] = fl;
      f.flagR_peer[r] = fl + c.entities;
    }
  }
  if (lazy && m.n_hot > 0) {
    f.hot_slot = m.hot_slot; f.hot_acc = m.hot_acc; f.hot_replicas = kHotReplicas;
  }
  if (lazy && m.pending_bank >= 0) {
    f.fin_acc = acc_bank(m, m.pending_bank); f.fin_loss = m.pending_loss;
    f.fin_inv_B = 1.0f / (float)B; f.fin_reg_coeff = c.lambda > 0.f ? c.lambda / (2.0f * (float)B) : 0.f;
  }
  const int fused = m.use_fused ? launch_vs_fused(f, m.WpT, !m.wpt_valid, m.use_fused, st) : 1;
  if (fused == 0) m.wpt_valid = true;
  if (fused < 0) return -1;
  if (traced) trace_mark(m, 1);
  // table shards: the projection matrix arrives from its owner, whose update alone can keep the transposed copy
  if (!dense_owner) m.wpt_valid = false;
  if (m.want_fused_event) SERT_CUDA(cudaEventRecord(m.ev_fused, st));   // the previous step's loss is final here
  if (lazy) { m.pending_bank = -1; m.pending_loss = nullptr; }      // the tile kernel has taken care of it
  if (fused == 1) {
    // general-shape path: one kernel per stage
    if (vs_forward(m, x, st)) return -1;
    VsNceArgs a;
    a.t = m.t; a.Eemb = m.theta + m.off[SERT_PARAM_ENTITY_REPR]; a.y = y; a.neg = neg; a.w = w;
    a.gE = m.grad + m.off[SERT_PARAM_ENTITY_REPR]; a.flagE = m.flagE; a.stamp = m.stamp; a.da = m.da;
    a.loss_acc = acc; a.dbg_scores = nullptr; a.dbg_u = nullptr; a.dbg_ell = nullptr;
    a.B = B; a.k = c.num_negatives; a.de = de; a.inv_B = 1.0f / (float)B; a.train = true;
    a.own = m.table_own;
    if (launch_vs_nce(a, st)) return -1;
    // dh = da . Wp^T
    if (launch_gemm_f32(m.da, Wp, m.dh, B, dw, de, false, true, de, de, dw, EPI_STORE, nullptr, 1, st)) return -1;
    if (launch_scatter_rows(x, m.dh, m.grad + m.off[SERT_PARAM_WORD_REPR], m.flagR, m.stamp, B, c.window, dw,
                            (float)c.window, st, m.table_own.r_lo, m.table_own.r_hi))
      return -1;
  }
  // The gradients of the dense tensors (gWp = h^T . da by split-K, gbp = colsum(da)) are only consumed by the
  // last 16.5k parameters of the arena, so they run on a second stream concurrently with the Adam stream over
  // the two tables (phase 3); the dense tensors are updated behind them (phase 4).
  cudaStream_t side = overlap ? m.st2 : st;
  if (overlap) {
    SERT_CUDA(cudaEventRecord(m.ev_fork, st));
    SERT_CUDA(cudaStreamWaitEvent(side, m.ev_fork, 0));
  }
  if (inst) {
    // this rank's instances' share of the dense gradients goes straight into the arena of the rank that updates the
    // projection (float atomics over NVLink), its share of the hot rows' gradients to their owners
    const int world = m.table_comm->world, Bl = f.B - f.i0;
    // (accumulated in this rank's own -- otherwise unused -- dense gradient rows first: the split-K GEMM's scalar atomics
    // are local, one pass of vector reductions carries the sum over NVLink)
    if (launch_gemm_f32(m.h, m.da, m.grad + m.off[SERT_PARAM_DENSE_W], dw, de, Bl, true, false, dw, de, de,
                        EPI_ATOMIC_ADD, nullptr, pick_split_k(dw, de, Bl), side))
      return -1;
    if (launch_colsum_atomic(m.da, m.grad + m.off[SERT_PARAM_DENSE_B], Bl, de, side)) return -1;
    if (!dense_owner) {
      const long long dense0 = m.off[SERT_PARAM_DENSE_W];
      if (launch_push_add(m.grad + dense0, static_cast<float *>(m.gsh_peers[world - 1]) + dense0, m.total - dense0, side))
        return -1;
    }
    if (m.n_hot > 0) {
      HotPushArgs hp;
      hp.hot_acc = m.hot_acc; hp.hot_ids = m.hot_ids; hp.n_hot = m.n_hot; hp.d = dw;
      hp.table_offset = m.off[SERT_PARAM_WORD_REPR];
      hp.n_owner = world;
      for (int r = 0; r <= world; ++r) hp.r_bound[r] = m.r_bound[r];
      for (int r = 0; r < world; ++r) hp.grad_peer[r] = static_cast<float *>(m.gsh_peers[r]);
      if (launch_hot_push(hp, side)) return -1;
    }
    // every rank's gradient rows must have landed in their owners' arenas before any update reads them
    SERT_CUDA(cudaEventRecord(m.ev_join, side));
    SERT_CUDA(cudaStreamWaitEvent(st, m.ev_join, 0));
    if (traced) trace_mark(m, 2);
    if (!(ts_dbg && strchr(ts_dbg, 'a')) && peer_barrier(m, nullptr, 0)) return -1;
    if (traced) trace_mark(m, 3);
    SERT_CUDA(cudaEventRecord(m.ev_fork, st));
    SERT_CUDA(cudaStreamWaitEvent(side, m.ev_fork, 0));
  } else if (traced) {
    trace_mark(m, 2);
    trace_mark(m, 3);
  }
  if (!inst && dense_owner) {
    if (launch_gemm_f32(m.h, m.da, m.grad + m.off[SERT_PARAM_DENSE_W], dw, de, B, true, false, dw, de, de,
                        EPI_ATOMIC_ADD, nullptr, pick_split_k(dw, de, B), side))
      return -1;
    if (launch_colsum_atomic(m.da, m.grad + m.off[SERT_PARAM_DENSE_B], B, de, side)) return -1;
  }
  m.step = t_next;
  OptimArgs o = optim_args(m, loss_out);
  o.acc = acc;
  o.c0 = adam_alpha_f32(m.step); o.c1 = 0.9f; o.c2 = 0.999f; o.c3 = 1e-8f;
  o.push_all = push_all ? 1 : 0;
  o.need_stamp = next_stamp(m.stamp);
  if (!overlap) {
    m.wpt_valid = false;
    if (!sharded) return timed_update(m, o, true);
    // table shards: this rank's piece of the tables, then (their owner) the dense tensors, then the exchange
    OptimArgs piece = o;
    piece.no_finalize = true;
    piece.phase = 3;
    piece.first4 = m.table_lo4[m.table_comm->rank];
    piece.last4 = m.table_lo4[m.table_comm->rank + 1];
    if (timed_update(m, piece, true)) return -1;
    if (dense_owner) {
      OptimArgs tail = o;
      tail.phase = 4;
      tail.first4 = m.off[SERT_PARAM_DENSE_W] / 4;
      tail.loss_out = nullptr;
      if (launch_adam(tail, st)) return -1;
    }
    if (table_exchange(m, acc)) return -1;
    return launch_finalize_train(acc, loss_out, o.inv_B, o.reg_coeff, st);
  }
  OptimArgs dense = o;
  dense.phase = 4;
  dense.first4 = m.off[SERT_PARAM_DENSE_W] / 4;       // W and b are the last two tensors of the arena
  if (m.wpt_valid && de % 4 == 0) {                   // the update keeps the transposed copy current: no transpose launch
    dense.transposed = m.WpT;
    dense.transposed_rows = dw;
    for (int s = 0; s < m.nseg; ++s)
      if (m.seg[s].offset == m.off[SERT_PARAM_DENSE_W]) dense.transposed_segment = s;
  } else {
    m.wpt_valid = false;
  }
  OptimArgs tables = o;
  tables.loss_out = nullptr;
  tables.phase = 3;
  if (sharded) {
    tables.first4 = m.table_lo4[m.table_comm->rank];
    tables.last4 = m.table_lo4[m.table_comm->rank + 1];
  }
  if (lazy) {
    // second stream: dense tensors and hot rows; first stream: the two tables; join; loss left pending
    dense.loss_out = nullptr;
    if (dense_owner && launch_adam(dense, side)) return -1;
    if (m.n_hot > 0) {
      HotUpdateArgs h;
      h.theta = m.theta; h.s1 = m.s1; h.s2 = m.s2; h.grad = m.grad;
      h.table_offset = m.off[SERT_PARAM_WORD_REPR]; h.d = dw;
      h.hot_acc = m.hot_acc; h.hot_ids = m.hot_ids; h.n_hot = m.n_hot;
      h.l2_scale = o.l2_scale; h.c0 = o.c0; h.c1 = o.c1; h.c2 = o.c2; h.c3 = o.c3;
      h.acc = acc; h.counted = 1;
      h.state_bf16 = o.state_bf16; h.stamp = o.stamp;
      if (sharded) { h.own_lo4 = tables.first4; h.own_hi4 = tables.last4; }
      h.theta_out = o.theta_out; h.n_peers = o.n_peers;
      for (int p = 0; p < o.n_peers; ++p) h.peer_theta[p] = o.peer_theta[p];
      if (launch_hot_update(h, side)) return -1;
    }
    SERT_CUDA(cudaEventRecord(m.ev_join, side));
    if (timed_update(m, tables, true)) return -1;     // profile mode: events around the table stream, in situ
    if (traced) trace_mark(m, 4);
    SERT_CUDA(cudaStreamWaitEvent(st, m.ev_join, 0));
    if (sharded && table_exchange(m, acc)) return -1;
    if (traced) trace_mark(m, 5);
    m.pending_bank = bank;
    m.pending_loss = loss_out;
    return 0;
  }
  SERT_CUDA(cudaEventRecord(m.ev_join, side));
  if (timed_update(m, tables, true)) return -1;
  SERT_CUDA(cudaStreamWaitEvent(st, m.ev_join, 0));
  if (sharded) {                  // the loss waits for the other ranks' share of sum(theta^2)
    dense.loss_out = nullptr;
    if ((dense_owner && launch_adam(dense, st)) || table_exchange(m, acc)) return -1;
    return launch_finalize_train(acc, loss_out, o.inv_B, o.reg_coeff, st);
  }
  dense.ticket = reinterpret_cast<unsigned int *>(m.acc + kAccDoubles - 4);
  return launch_adam(dense, st);
}

static int vs_eval_step(sert_model &m, const int32_t *x, const int32_t *y, const int32_t *neg,
                        float *loss_out, bool debug) {
  const sert_config &c = m.cfg;
  cudaStream_t st = m.st;
  if (flush_pending(m)) return -1;
  if (neg == nullptr) {
    // the eval loss draws its own negatives (loss_fn is instantiated twice, sert/models.py:745-752)
    if (launch_sample_negatives(m.neg, (int64_t)c.batch * c.num_negatives, c.entities, c.seed,
                                m.sample_calls++, st))
      return -1;
    neg = m.neg;
  }
  if (vs_forward(m, x, st)) return -1;
  VsNceArgs a;
  a.t = m.t; a.Eemb = m.theta + m.off[SERT_PARAM_ENTITY_REPR]; a.y = y; a.neg = neg; a.w = nullptr;
  a.gE = nullptr; a.flagE = nullptr; a.stamp = 0; a.da = nullptr; a.loss_acc = m.acc;
  a.dbg_scores = debug ? m.dbg_scores : nullptr; a.dbg_u = debug ? m.dbg_u : nullptr;
  a.dbg_ell = debug ? m.dbg_ell : nullptr;
  a.B = c.batch; a.k = c.num_negatives; a.de = c.entity_dim; a.inv_B = 1.0f / (float)c.batch; a.train = false;
  if (launch_vs_nce(a, st)) return -1;
  return launch_finalize_eval(m.acc, loss_out, 1.0f / (float)c.batch, st);
}

__global__ void fill_bf16_kernel(__nv_bfloat16 *dst, long long n, float value) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(value);
}

// ---- log-linear ---------------------------------------------------------------------------------
// The tcgen05 kernel works on 128 x 256 output tiles with one persistent CTA per SM: it is used when the
// output has enough tiles to occupy the chip, the fp32 FMA tiles (with split-K) otherwise.
static bool ll_tensor(const sert_model &m, long long M, long long N) {
  return m.use_tensor && ((M + 127) / 128) * ((N + 255) / 256) >= 32;
}

// C = A . B^T over split operands of ll_terms() blocks of Kp columns
static int ll_gemm(const __nv_bfloat16 *A, int M, const __nv_bfloat16 *B, long long N_total, long long n_begin,
                   long long n_end, int Kp, const TcEpilogue &ep, cudaStream_t st) {
  return ll_terms() == 2 ? launch_gemm_tc_pair(A, M, B, N_total, n_begin, n_end, Kp, ep, st)
                         : launch_gemm_tc(A, M, B, N_total, n_begin, n_end, 3 * Kp, ep, st);
}

static int ll_forward(sert_model &m, const int32_t *x, int rows /* instances */, cudaStream_t st,
                      float *rmax_out = nullptr, float *rsum_out = nullptr) {
  const sert_config &c = m.cfg;
  const int E = (int)c.entities, dw = c.word_dim;
  const long long BW = (long long)rows * c.window;
  float *R = m.theta + m.off[SERT_PARAM_WORD_REPR];
  float *Wd = m.theta + m.off[SERT_PARAM_DENSE_W];
  float *bd = m.theta + m.off[SERT_PARAM_DENSE_B];
  if (launch_gather_rows(x, R, m.X, BW, dw, st)) return -1;
  if (ll_tensor(m, BW, E)) {
    // Z = X . Wd + bd on the tensor cores: bf16x3 split of X (A) and of Wd^T (B), fp32 accumulation
    const int T = ll_terms();
    if (launch_split_bf16(m.X, BW, dw, dw, T, SPLIT_A, m.Xs, st)) return -1;
    if (launch_split_bf16_t(Wd, dw, E, E, T, SPLIT_B, m.WdT_s, st)) return -1;
    TcEpilogue ep;
    ep.mode = TC_EPI_STORE; ep.C = m.Z; ep.ldc = E; ep.bias = bd;
    // the epilogue leaves per-slice softmax statistics: no pass over Z for the row maxima and sums
    const int slots = cdiv(E, 64);
    ep.row_stats = m.zstats; ep.stats_ld = slots;
    if (ll_gemm(m.Xs, (int)BW, m.WdT_s, E, 0, E, tc_padded_k(dw), ep, st)) return -1;
    return launch_ll_combine_slices(m.zstats, BW, slots, slots, rmax_out ? rmax_out : m.rmax,
                                    rsum_out ? rsum_out : m.rsum, st);
  } else {
    if (launch_gemm_f32(m.X, Wd, m.Z, (int)BW, E, dw, false, false, dw, E, E, EPI_BIAS, bd, 1, st)) return -1;
  }
  return launch_ll_row_stats(m.Z, BW, E, E, rmax_out ? rmax_out : m.rmax, rsum_out ? rsum_out : m.rsum, st);
}

static LlInstanceArgs ll_instance_args(sert_model &m, const int64_t *indptr, long long nnz_base,
                                       const int32_t *indices, const float *data, const float *w, bool train,
                                       float *ell_out) {
  LlInstanceArgs a;
  a.S = m.S; a.DS = train ? m.DS : nullptr; a.lds = m.cfg.entities; a.B = m.cfg.batch; a.E = (int)m.cfg.entities;
  a.indptr = reinterpret_cast<const long long *>(indptr); a.nnz_base = nnz_base; a.indices = indices;
  a.data = data; a.w = w; a.inv_B = 1.0f / (float)m.cfg.batch; a.ell_out = ell_out; a.loss_acc = m.acc;
  a.train = train;
  return a;
}

// gWd += X^T . dZ ; gbd += colsum(dZ) ; dX = dZ . Wd^T   (dZ lives in m.Z)
static int ll_backward_gemms(sert_model &m, int BW, int E, int dw, float *Wd, cudaStream_t st) {
  if (ll_tensor(m, dw, E)) {
    // gWd = X^T . dZ : A = X^T (dw, B*W), B = dZ^T (E, B*W), both as transposed bf16x3 splits
    const int T = ll_terms();
    if (launch_split_bf16_t(m.X, BW, dw, dw, T, SPLIT_A, m.XT_s, st)) return -1;
    if (launch_split_bf16_t(m.Z, BW, E, E, T, SPLIT_B, m.dZT_s, st)) return -1;
    TcEpilogue ep;
    ep.mode = TC_EPI_STORE; ep.C = m.grad + m.off[SERT_PARAM_DENSE_W]; ep.ldc = E;   // overwrites the (zeroed) grad
    if (ll_gemm(m.XT_s, dw, m.dZT_s, E, 0, E, tc_padded_k(BW), ep, st)) return -1;
  } else {
    if (launch_gemm_f32(m.X, m.Z, m.grad + m.off[SERT_PARAM_DENSE_W], dw, E, BW, true, false, dw, E, E,
                        EPI_ATOMIC_ADD, nullptr, pick_split_k(dw, E, BW), st))
      return -1;
  }
  if (launch_colsum_atomic(m.Z, m.grad + m.off[SERT_PARAM_DENSE_B], BW, E, st)) return -1;
  if (ll_tensor(m, BW, dw)) {
    // dX = dZ . Wd^T : A = dZ (B*W, E), B = Wd (dw, E)
    const int T = ll_terms();
    if (launch_split_bf16(m.Z, BW, E, E, T, SPLIT_A, m.dZs, st)) return -1;
    if (launch_split_bf16(Wd, dw, E, E, T, SPLIT_B, m.Wd_s, st)) return -1;
    TcEpilogue ep;
    ep.mode = TC_EPI_STORE; ep.C = m.dX; ep.ldc = dw;
    if (ll_gemm(m.dZs, BW, m.Wd_s, dw, 0, dw, tc_padded_k(E), ep, st)) return -1;
  } else {
    if (launch_gemm_f32(m.Z, Wd, m.dX, BW, dw, E, false, true, E, E, dw, EPI_STORE, nullptr, 1, st)) return -1;
  }
  return 0;
}

// Both gradient GEMMs on tensor cores: dZ is written once, directly as their split bf16 operands (ll_kernels.cu).
static bool ll_fused_tail(const sert_model &m, long long BW, long long E, long long dw) {
  return ll_tensor(m, dw, E) && ll_tensor(m, BW, dw);
}

// racc (complete over all shards) -> dZs / dZT_s -> gWd (+ gbd through the ones row of X^T) and dX
static int ll_backward_fused(sert_model &m, int B, int W, int E, int dw, float *Wd, cudaStream_t st) {
  const int BW = B * W;
  const int T = ll_terms();
  // Pair operands: gWd = X^T . dZ reads the rows of dZs as an N-major B operand, so dZ is written ONCE (8 instead of
  // 16 GB of operand writes at BASELINE configs[4]).  SERT_LL_BN=0: the transposed copy dZT_s as K-major B.
  static const char *bn_env = getenv("SERT_LL_BN");
  const bool bn = T == 2 && !(bn_env != nullptr && bn_env[0] == '0');
  if (launch_ll_dz_split(m.Z, m.rmax, m.lrsum, m.racc, m.DS, B, W, E, E, E, T, m.dZs, bn ? nullptr : m.dZT_s, st)) return -1;
  {
    if (launch_split_bf16_t(m.X, BW, dw, dw, T, SPLIT_A, m.XT_s, st)) return -1;
    TcEpilogue ep;
    ep.mode = TC_EPI_STORE; ep.C = m.grad + m.off[SERT_PARAM_DENSE_W]; ep.ldc = E;   // overwrites the (zeroed) grads
    ep.extra_row = dw; ep.extra_dst = m.grad + m.off[SERT_PARAM_DENSE_B];
    if (bn) {
      if (launch_gemm_tc_pair_bn(m.XT_s, dw + 1, m.dZs, BW, tc_padded_k(E), E, 0, E, tc_padded_k(BW), ep, st)) return -1;
    } else if (ll_gemm(m.XT_s, dw + 1, m.dZT_s, E, 0, E, tc_padded_k(BW), ep, st)) {
      return -1;
    }
  }
  {
    if (launch_split_bf16(Wd, dw, E, E, T, SPLIT_B, m.Wd_s, st)) return -1;
    TcEpilogue ep;
    ep.mode = TC_EPI_STORE; ep.C = m.dX; ep.ldc = dw;
    ep.accumulate = 1;                                  // split-K over the entity axis (K = 3 E64)
    SERT_CUDA(cudaMemsetAsync(m.dX, 0, (size_t)BW * dw * sizeof(float), st));
    if (ll_gemm(m.dZs, BW, m.Wd_s, dw, 0, dw, tc_padded_k(E), ep, st)) return -1;
  }
  return 0;
}

static int ll_train_step(sert_model &m, const int32_t *x, const int64_t *indptr, long long nnz_base,
                         const int32_t *indices, const float *data, const float *w, float *loss_out) {
  const sert_config &c = m.cfg;
  cudaStream_t st = m.st;
  const int B = c.batch, W = c.window, E = (int)c.entities, dw = c.word_dim;
  const int BW = B * W;
  float *Wd = m.theta + m.off[SERT_PARAM_DENSE_W];
  m.stamp += 1;
  if (ll_forward(m, x, B, st)) return -1;
  {
    const int have = launch_ll_joint(m.Z, m.rmax, m.rsum, m.S, B, W, E, E, E, st, m.lrsum, m.sstats);
    if (have < 0) return -1;
    LlInstanceArgs ia = ll_instance_args(m, indptr, nnz_base, indices, data, w, true, nullptr);
    if (have == 1) { ia.sstats = m.sstats; ia.slots = ll_joint_slots(E); }
    if (launch_ll_instance(ia, st)) return -1;
  }
  if (ll_fused_tail(m, BW, E, dw)) {
    if (launch_ll_racc_log(m.Z, m.rmax, m.lrsum, m.DS, B, W, E, E, E, m.racc, st)) return -1;
    if (ll_backward_fused(m, B, W, E, dw, Wd, st)) return -1;
  } else {
    if (launch_ll_dz(m.Z, m.rmax, m.rsum, m.DS, B, W, E, E, E, st)) return -1;
    if (ll_backward_gemms(m, BW, E, dw, Wd, st)) return -1;
  }
  if (launch_scatter_rows(x, m.dX, m.grad + m.off[SERT_PARAM_WORD_REPR], m.flagR, m.stamp, BW, 1, dw, 1.0f, st))
    return -1;
  m.step += 1;
  OptimArgs o = optim_args(m, loss_out);
  o.c0 = 1.0f; o.c1 = 0.95f; o.c2 = 0.f; o.c3 = 1e-6f;   // lasagne.updates.adadelta defaults
  return timed_update(m, o, false);
}

static int ll_eval_step(sert_model &m, const int32_t *x, const int64_t *indptr, long long nnz_base,
                        const int32_t *indices, const float *data, float *loss_out, bool debug) {
  const sert_config &c = m.cfg;
  cudaStream_t st = m.st;
  const int B = c.batch, W = c.window, E = (int)c.entities;
  if (ll_forward(m, x, B, st)) return -1;
  {
    const int have = launch_ll_joint(m.Z, m.rmax, m.rsum, m.S, B, W, E, E, E, st, m.lrsum, m.sstats);
    if (have < 0) return -1;
    LlInstanceArgs ia = ll_instance_args(m, indptr, nnz_base, indices, data, nullptr, false, debug ? m.dbg_ell : nullptr);
    if (have == 1) { ia.sstats = m.sstats; ia.slots = ll_joint_slots(E); }
    if (launch_ll_instance(ia, st)) return -1;
  }
  return launch_finalize_eval(m.acc, loss_out, 1.0f / (float)B, st);
}

// ---- log-linear, entity-sharded (SURVEY.md 8(e) "column-parallel softmax") -----------------------
// cfg.entities is the shard width; CSR label indices are global entity ids.  Exchange points of one
// training step (all through m.exchange, ordered on the model's stream):
//   (1) all-gather of the per-word (row max, row sum)           2*B*W floats per rank
//   (2) all-gather of the joint (row max, row sum)              2*B   floats per rank
//   (3) all-reduce of sum_e do*o per instance                   B     floats
//   (4) all-reduce of sum_e dp*p per word                       B*W   floats
//   (5) all-reduce of the partial dX = dZ_loc . Wd_loc^T        B*W*dw floats
// The word table R and its Adadelta state are replicated and receive the identical update on every rank.
static int xchg(sert_model &m, int op, float *buf, size_t count) {
  SERT_CUDA(cudaGetLastError());
  const int rc = m.exchange(m.exchange_ctx, op, buf, count);
  SERT_REQUIRE(rc == 0, "exchange callback failed");
  return 0;
}

// this rank's slot of the gathered statistics: [shard][2][rows]
static float *ll_my_stats(sert_model &m, long long rows) { return m.xstats + (size_t)m.shard_rank * 2 * rows; }
// all-gathers the per-shard (row max, row sum) written to ll_my_stats() and combines them into (gmax, gsum)
static int ll_shard_combine(sert_model &m, long long rows, float *gmax, float *gsum) {
  if (xchg(m, SERT_XCHG_ALLGATHER, m.xstats, (size_t)2 * rows)) return -1;
  return launch_ll_combine_stats(m.xstats, m.shard_world, rows, gmax, gsum, m.st);
}

static int ll_shard_forward(sert_model &m, const int32_t *x, const int64_t *indptr, long long nnz_base,
                            const int32_t *indices, const float *data, const float *w, bool train,
                            float *ell_out) {
  const sert_config &c = m.cfg;
  const int B = c.batch, W = c.window, E = (int)c.entities;
  const long long BW = (long long)B * W;
  float *mine = ll_my_stats(m, BW);
  if (ll_forward(m, x, B, m.st, mine, mine + BW)) return -1;      // local columns of Z and their statistics
  if (ll_shard_combine(m, BW, m.rmax, m.rsum)) return -1;
  if (launch_ll_joint(m.Z, m.rmax, m.rsum, m.S, B, W, E, E, E, m.st, m.lrsum) < 0) return -1;
  mine = ll_my_stats(m, B);
  if (launch_ll_row_stats(m.S, B, E, E, mine, mine + B, m.st)) return -1;
  if (ll_shard_combine(m, B, m.smax, m.ssum)) return -1;
  LlInstanceArgs a = ll_instance_args(m, indptr, nnz_base, indices, data, w, train, ell_out);
  return launch_ll_shard_labels(a, m.smax, m.ssum, (int)m.e_begin, train ? m.adot : nullptr, m.st);
}

static int ll_train_step_sharded(sert_model &m, const int32_t *x, const int64_t *indptr, long long nnz_base,
                                 const int32_t *indices, const float *data, const float *w, float *loss_out) {
  const sert_config &c = m.cfg;
  cudaStream_t st = m.st;
  const int B = c.batch, W = c.window, E = (int)c.entities, dw = c.word_dim;
  const int BW = B * W;
  float *Wd = m.theta + m.off[SERT_PARAM_DENSE_W];
  m.stamp += 1;
  if (ll_shard_forward(m, x, indptr, nnz_base, indices, data, w, true, nullptr)) return -1;
  if (xchg(m, SERT_XCHG_ALLREDUCE_SUM, m.adot, (size_t)B)) return -1;
  LlInstanceArgs a = ll_instance_args(m, indptr, nnz_base, indices, data, w, true, nullptr);
  if (launch_ll_shard_ds(a, m.smax, m.ssum, (int)m.e_begin, m.adot, st)) return -1;
  if (ll_fused_tail(m, BW, E, dw)) {
    if (launch_ll_racc_log(m.Z, m.rmax, m.lrsum, m.DS, B, W, E, E, E, m.racc, st)) return -1;
    if (xchg(m, SERT_XCHG_ALLREDUCE_SUM, m.racc, (size_t)BW)) return -1;
    if (ll_backward_fused(m, B, W, E, dw, Wd, st)) return -1;
  } else {
    if (launch_ll_dz(m.Z, m.rmax, m.rsum, m.DS, B, W, E, E, E, st, 1, m.racc)) return -1;
    if (xchg(m, SERT_XCHG_ALLREDUCE_SUM, m.racc, (size_t)BW)) return -1;
    if (launch_ll_dz(m.Z, m.rmax, m.rsum, m.DS, B, W, E, E, E, st, 2, m.racc)) return -1;
    if (ll_backward_gemms(m, BW, E, dw, Wd, st)) return -1;
  }
  if (xchg(m, SERT_XCHG_ALLREDUCE_SUM, m.dX, (size_t)BW * dw)) return -1;
  if (launch_scatter_rows(x, m.dX, m.grad + m.off[SERT_PARAM_WORD_REPR], m.flagR, m.stamp, BW, 1, dw, 1.0f, st))
    return -1;
  m.step += 1;
  OptimArgs o = optim_args(m, loss_out);
  o.c0 = 1.0f; o.c1 = 0.95f; o.c2 = 0.f; o.c3 = 1e-6f;
  return timed_update(m, o, false);
}

static int ll_eval_step_sharded(sert_model &m, const int32_t *x, const int64_t *indptr, long long nnz_base,
                                const int32_t *indices, const float *data, float *loss_out, bool debug) {
  if (ll_shard_forward(m, x, indptr, nnz_base, indices, data, nullptr, false, debug ? m.dbg_ell : nullptr))
    return -1;
  return launch_finalize_eval(m.acc, loss_out, 1.0f / (float)m.cfg.batch, m.st);
}

// Per-batch losses of a sharded model are partial sums (labels owned by the shard, norm of the shard's columns;
// rank 0 adds the replicated word table): one all-reduce over the slots completes them.
static int ll_reduce_losses(sert_model &m, int32_t first_slot, int64_t n) {
  if (m.shard_world <= 1 || n == 0) return 0;
  return xchg(m, SERT_XCHG_ALLREDUCE_SUM, m.losses + first_slot, (size_t)n);
}

static int check_batch(sert_model *m, int split, int64_t b) {
  SERT_REQUIRE(split == SERT_SPLIT_TRAIN || split == SERT_SPLIT_VALIDATE, "unknown split");
  const Dataset &d = m->ds[split];
  SERT_REQUIRE(d.x != nullptr, "no data set attached for this split");
  SERT_REQUIRE(b >= 0 && (b + 1) * (int64_t)m->cfg.batch <= d.n,
               "batch index out of range (incomplete batches are ignored, sert/models.py:355-359)");
  return 0;
}

}  // namespace sert

// =================================================================================================
// extern "C"
// =================================================================================================
extern "C" {

int sert_abi_version(void) { return SERT_ABI_VERSION; }
const char *sert_last_error(void) { return g_error.c_str(); }
uint64_t sert_launch_count(void) { return g_launches.load(); }

int sert_model_arena_bytes(const sert_config *cfg, size_t *bytes) {
  SERT_REQUIRE(cfg && bytes, "null argument");
  if (validate(*cfg)) return -1;
  sert_model tmp;
  tmp.cfg = *cfg;
  *bytes = carve(tmp, nullptr);
  return 0;
}

static int table_shard_partition(const sert_model &m, int world, long long *e_row, long long *r_row, long long *lo4,
                                 int *inst);

// Host only (no device needed): how sert_model_set_table_shard_comm would cut a model of this configuration over `world`
// ranks.  Arrays of world + 1 entries.
int sert_table_shard_plan(const sert_config *cfg, int32_t world, int64_t *entity_row_bounds, int64_t *word_row_bounds,
                          int64_t *float_bounds, int32_t *instance_bounds) {
  SERT_REQUIRE(cfg && entity_row_bounds && word_row_bounds && float_bounds && instance_bounds, "null argument");
  if (validate(*cfg)) return -1;
  SERT_REQUIRE(is_vs(*cfg), "table shards are for vector-space models");
  sert_model tmp;
  tmp.cfg = *cfg;
  carve(tmp, nullptr);
  long long e_row[kMaxPeers + 2], r_row[kMaxPeers + 2], lo4[kMaxPeers + 2];
  int inst[kMaxPeers + 2];
  if (table_shard_partition(tmp, world, e_row, r_row, lo4, inst)) return -1;
  for (int r = 0; r <= world; ++r) {
    entity_row_bounds[r] = e_row[r];
    word_row_bounds[r] = r_row[r];
    float_bounds[r] = lo4[r] * 4;
    instance_bounds[r] = inst[r];
  }
  return 0;
}

int sert_model_create(const sert_config *cfg, void *arena_dev, size_t arena_bytes, void *stream,
                      sert_model **out) {
  SERT_REQUIRE(cfg && arena_dev && out, "null argument");
  if (validate(*cfg)) return -1;
  SERT_REQUIRE(((uintptr_t)arena_dev & 255) == 0, "arena must be 256-byte aligned");
  int ndev = 0;
  SERT_CUDA(cudaGetDeviceCount(&ndev));
  SERT_REQUIRE(ndev > 0, "no CUDA device: libsert_b200 has no CPU fallback");
  sert_model *m = new sert_model();
  m->cfg = *cfg;
  const size_t need = carve(*m, arena_dev);
  if (need > arena_bytes) {
    delete m;
    set_error("arena too small: need " + std::to_string(need) + " bytes");
    return -1;
  }
  m->arena = static_cast<char *>(arena_dev);
  m->arena_bytes = arena_bytes;
  m->st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(arena_dev, 0, need, m->st);
  if (e != cudaSuccess) {
    delete m;
    set_error(std::string("cudaMemsetAsync: ") + cudaGetErrorString(e));
    return -1;
  }
  if (!is_vs(*cfg) && m->XT_s != nullptr) {
    // the row of ones behind X^T (split [hi | hi | mid] = [1 | 1 | 0], or [hi | mid] = [1 | 0] for pair operands, over
    // the B*W real rows): see ll_backward_fused
    const long long BW = (long long)cfg->batch * cfg->window, BW64 = tc_padded_k((int)BW);
    __nv_bfloat16 *row = m->XT_s + (size_t)cfg->word_dim * ll_terms() * BW64;
    fill_bf16_kernel<<<cdiv(BW, 256), 256, 0, m->st>>>(row, BW, 1.0f);
    if (ll_terms() == 3) fill_bf16_kernel<<<cdiv(BW, 256), 256, 0, m->st>>>(row + BW64, BW, 1.0f);
    count_launch(2);
  }
  if (is_vs(*cfg)) {
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    if (cudaStreamCreateWithPriority(&m->st2, cudaStreamNonBlocking, greatest) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      m->st2 = nullptr;                  // no overlap stream: the step runs on one stream
      cudaGetLastError();
    }
  }
  *out = m;
  return 0;
}

static void table_shard_release(sert_model *m);

int sert_model_destroy(sert_model *m) {
  if (m) {
    cudaStreamSynchronize(m->st);
    if (m->st2) {
      cudaStreamSynchronize(m->st2);
      cudaStreamDestroy(m->st2);
    }
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    if (m->st_h2d) { cudaStreamSynchronize(m->st_h2d); cudaStreamDestroy(m->st_h2d); }
    if (m->st_d2h) { cudaStreamSynchronize(m->st_d2h); cudaStreamDestroy(m->st_d2h); }
    for (int q = 0; q < 2; ++q) {
      if (m->ev_copied[q]) cudaEventDestroy(m->ev_copied[q]);
      if (m->ev_free[q]) cudaEventDestroy(m->ev_free[q]);
    }
    if (m->ev_fused) cudaEventDestroy(m->ev_fused);
    for (int q = 0; q < sert_model::kPipe; ++q)
      if (m->ev_loss[q]) cudaEventDestroy(m->ev_loss[q]);
    if (m->pin_loss) cudaFreeHost(m->pin_loss);
    if (m->rank_scratch) cudaFree(m->rank_scratch);
    table_shard_release(m);
    delete m;
  }
  return 0;
}

static float *tensor_ptr(sert_model *m, int which, int slot) {
  float *base = slot == SERT_STATE_PARAM ? m->theta : slot == SERT_STATE_S1 ? m->s1 : m->s2;
  return base + m->off[which];
}

int sert_model_set_tensor(sert_model *m, int which, int slot, const float *host, size_t count) {
  SERT_REQUIRE(m && host, "null argument");
  SERT_REQUIRE(which >= 0 && which < 4 && m->cnt[which] > 0, "model has no such tensor");
  SERT_REQUIRE(slot >= 0 && slot <= 2, "bad state slot");
  SERT_REQUIRE(slot == SERT_STATE_PARAM || m->cfg.inference_only == 0, "inference_only models keep no optimiser state");
  SERT_REQUIRE((long long)count == m->cnt[which], "tensor size mismatch");
  if (slot != SERT_STATE_PARAM && m->cfg.dtype_mode == 1) {
    // bfloat16 state: round to nearest even on the host (checkpoints written by this mode are exact in bf16)
    std::vector<uint16_t> tmp(count);
    for (size_t i = 0; i < count; ++i) {
      uint32_t u;
      memcpy(&u, host + i, 4);
      tmp[i] = (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
    }
    uint16_t *dst = reinterpret_cast<uint16_t *>(slot == SERT_STATE_S1 ? m->s1 : m->s2) + m->off[which];
    SERT_CUDA(cudaMemcpyAsync(dst, tmp.data(), count * sizeof(uint16_t), cudaMemcpyHostToDevice, m->st));
    SERT_CUDA(cudaStreamSynchronize(m->st));
    return 0;
  }
  SERT_CUDA(cudaMemcpyAsync(tensor_ptr(m, which, slot), host, count * sizeof(float), cudaMemcpyHostToDevice, m->st));
  SERT_CUDA(cudaStreamSynchronize(m->st));
  if (which == SERT_PARAM_DENSE_W && slot == SERT_STATE_PARAM) m->wpt_valid = false;
  return 0;
}

int sert_model_get_tensor(sert_model *m, int which, int slot, float *host, size_t count) {
  SERT_REQUIRE(m && host, "null argument");
  SERT_REQUIRE(which >= 0 && which < 4 && m->cnt[which] > 0, "model has no such tensor");
  SERT_REQUIRE(slot >= 0 && slot <= 2, "bad state slot");
  SERT_REQUIRE(slot == SERT_STATE_PARAM || m->cfg.inference_only == 0, "inference_only models keep no optimiser state");
  SERT_REQUIRE((long long)count == m->cnt[which], "tensor size mismatch");
  if (slot != SERT_STATE_PARAM && m->cfg.dtype_mode == 1) {
    std::vector<uint16_t> tmp(count);
    const uint16_t *src = reinterpret_cast<const uint16_t *>(slot == SERT_STATE_S1 ? m->s1 : m->s2) + m->off[which];
    SERT_CUDA(cudaMemcpyAsync(tmp.data(), src, count * sizeof(uint16_t), cudaMemcpyDeviceToHost, m->st));
    SERT_CUDA(cudaStreamSynchronize(m->st));
    for (size_t i = 0; i < count; ++i) {
      const uint32_t u = (uint32_t)tmp[i] << 16;
      memcpy(host + i, &u, 4);
    }
    return 0;
  }
  SERT_CUDA(cudaMemcpyAsync(host, tensor_ptr(m, which, slot), count * sizeof(float), cudaMemcpyDeviceToHost, m->st));
  SERT_CUDA(cudaStreamSynchronize(m->st));
  return 0;
}

int sert_model_set_step(sert_model *m, int64_t t) {
  SERT_REQUIRE(m && t >= 0, "bad argument");
  m->step = t;
  return 0;
}
int sert_model_get_step(sert_model *m, int64_t *t) {
  SERT_REQUIRE(m && t, "null argument");
  *t = m->step;
  return 0;
}

int sert_model_get_sampler(sert_model *m, uint64_t *seed, uint64_t *draws) {
  SERT_REQUIRE(m && seed && draws, "null argument");
  *seed = m->cfg.seed;
  *draws = m->sample_calls;
  return 0;
}
int sert_model_set_sampler(sert_model *m, uint64_t seed, uint64_t draws) {
  SERT_REQUIRE(m, "null model");
  m->cfg.seed = seed;
  m->sample_calls = draws;
  return 0;
}

int sert_model_set_fused(sert_model *m, int enable) {
  SERT_REQUIRE(m, "null model");
  SERT_REQUIRE(enable >= 0 && enable <= 2, "fused mode must be 0 (per-stage), 1 (auto) or 2 (warp kernel)");
  m->use_fused = enable;
  return 0;
}

int sert_model_set_hot_words(sert_model *m, const int32_t *ids_host, int32_t n) {
  SERT_REQUIRE(m && (ids_host || n == 0), "null argument");
  SERT_REQUIRE(is_vs(m->cfg) && m->cfg.inference_only == 0, "hot word rows apply to a trainable vector-space model");
  SERT_REQUIRE(n >= 0 && n <= kMaxHotRows, "at most 32 hot word rows");
  std::vector<int8_t> slot((size_t)m->cfg.vocab, (int8_t)-1);
  for (int32_t s = 0; s < n; ++s) {
    SERT_REQUIRE(ids_host[s] >= 0 && ids_host[s] < m->cfg.vocab, "hot word id out of range");
    SERT_REQUIRE(slot[ids_host[s]] < 0, "duplicate hot word id");
    slot[ids_host[s]] = (int8_t)s;
  }
  if (flush_pending(*m) || set_hot_marks(*m, false)) return -1;      // the old rows go back to the dense update
  SERT_CUDA(cudaStreamSynchronize(m->st));
  SERT_CUDA(cudaMemcpyAsync(m->hot_slot, slot.data(), slot.size(), cudaMemcpyHostToDevice, m->st));
  if (n > 0) SERT_CUDA(cudaMemcpyAsync(m->hot_ids, ids_host, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, m->st));
  SERT_CUDA(cudaStreamSynchronize(m->st));
  m->n_hot = n;
  return 0;
}

int sert_model_set_overlap(sert_model *m, int enable) {
  SERT_REQUIRE(m, "null model");
  m->overlap = enable != 0;
  return 0;
}

int sert_model_set_tensor_cores(sert_model *m, int enable) {
  SERT_REQUIRE(m, "null model");
  m->use_tensor = enable != 0;
  return 0;
}

int sert_model_set_entity_shard(sert_model *m, int32_t rank, int32_t world, int64_t entity_begin,
                                int64_t entities_total, sert_exchange_fn fn, void *ctx) {
  SERT_REQUIRE(m && fn, "null argument");
  SERT_REQUIRE(!is_vs(m->cfg), "entity sharding of the training step is defined for the log-linear model");
  SERT_REQUIRE(world >= 1 && world <= kMaxShards && rank >= 0 && rank < world, "bad shard rank / world size");
  SERT_REQUIRE(entity_begin >= 0 && entity_begin + m->cfg.entities <= entities_total && entities_total < (1ll << 31),
               "shard columns out of range");
  m->shard_rank = rank; m->shard_world = world; m->e_begin = entity_begin; m->e_total = entities_total;
  m->exchange = fn; m->exchange_ctx = ctx;
  // the replicated word table is L2-regularised on every rank but its norm enters the reported loss once
  for (int s = 0; s < m->nseg; ++s)
    if (m->seg[s].offset == m->off[SERT_PARAM_WORD_REPR]) m->seg[s].regularised = rank == 0 ? 1 : 2;
  return 0;
}

// the exchange "callback" of a model that shards over a library communicator: NCCL on the model's stream
static int nccl_exchange(void *ctx, int32_t op, float *buf, size_t count) {
  sert_model *m = static_cast<sert_model *>(ctx);
  if (op == SERT_XCHG_ALLREDUCE_SUM) return comm_all_reduce_sum_f32(m->comm, buf, count, m->st);
  if (op == SERT_XCHG_ALLGATHER)
    return comm_all_gather(m->comm, buf + (size_t)m->comm->rank * count, buf, count * sizeof(float), m->st);
  set_error("unknown exchange op");
  return -1;
}

int sert_model_set_entity_shard_comm(sert_model *m, sert_comm *comm, int64_t entity_begin, int64_t entities_total) {
  SERT_REQUIRE(m && comm, "null argument");
  m->comm = comm;
  return sert_model_set_entity_shard(m, comm->rank, comm->world, entity_begin, entities_total, nccl_exchange, m);
}

// Table shards: pieces of (nearly) equal float counts that end on row boundaries (row index a multiple of 4: whole
// 16-byte chunks), so that a gradient row is formed and updated by exactly one rank.  The entity table comes first in
// the arena (carve).  e_row / r_row [r], [r + 1] = rank r's entity / word rows, lo4 [r], [r + 1] = its 16-byte chunks of
// the arena arrays, inst [r], [r + 1] = its instances of a batch (tiles of 8) for instance shards.
static int table_shard_partition(const sert_model &m, int world, long long *e_row, long long *r_row, long long *lo4,
                                 int *inst) {
  SERT_REQUIRE(world >= 1 && world <= kMaxPeers + 1, "table shards span at most 8 ranks (one NVLink domain)");
  SERT_REQUIRE(m.off[SERT_PARAM_ENTITY_REPR] < m.off[SERT_PARAM_WORD_REPR] &&
               m.off[SERT_PARAM_WORD_REPR] < m.off[SERT_PARAM_DENSE_W] &&
               m.off[SERT_PARAM_DENSE_W] < m.off[SERT_PARAM_DENSE_B], "unexpected parameter layout");
  const long long tables4 = m.off[SERT_PARAM_DENSE_W] / 4;      // the two tables come first in the arena (carve)
  const long long E = m.cfg.entities, V = m.cfg.vocab, de = m.cfg.entity_dim, dw = m.cfg.word_dim;
  const long long offE = m.off[SERT_PARAM_ENTITY_REPR], offR = m.off[SERT_PARAM_WORD_REPR];
  for (int r = 0; r <= world; ++r) {
    const long long t = (E * de + V * dw) * r / world;     // floats before the boundary
    if (r == world) { e_row[r] = E; r_row[r] = V; lo4[r] = tables4; }
    else if (t < E * de) { e_row[r] = (t / de) & ~3ll; r_row[r] = 0; lo4[r] = (offE + e_row[r] * de) / 4; }
    else { e_row[r] = E; r_row[r] = ((t - E * de) / dw) & ~3ll; lo4[r] = (offR + r_row[r] * dw) / 4; }
  }
  lo4[0] = 0;
  const int tiles = cdiv(m.cfg.batch, 8);
  for (int r = 0; r <= world; ++r) inst[r] = (int)std::min<long long>(m.cfg.batch, (long long)tiles * r / world * 8);
  return 0;
}

static void table_shard_release(sert_model *m) {
  if (m->table_comm == nullptr) return;
  trace_report(*m);
  if (m->gsh != nullptr) {
    // gradients (zero between steps) and touched stamps return to the arena
    cudaMemcpyAsync(m->arena_flagE, m->flagE, (size_t)m->cfg.entities * sizeof(uint32_t), cudaMemcpyDeviceToDevice, m->st);
    cudaMemcpyAsync(m->arena_flagR, m->flagR, (size_t)m->cfg.vocab * sizeof(uint32_t), cudaMemcpyDeviceToDevice, m->st);
    cudaStreamSynchronize(m->st);
    for (int sg = 0; sg < m->nseg; ++sg) {
      if (m->seg[sg].flags == m->flagR) m->seg[sg].flags = m->arena_flagR;
      else if (m->seg[sg].flags == m->flagE) m->seg[sg].flags = m->arena_flagE;
    }
    m->grad = m->arena_grad; m->flagE = m->arena_flagE; m->flagR = m->arena_flagR;
    comm_unmap_peers(m->table_comm, m->gsh_peers);
    cudaFree(m->gsh);
    m->gsh = nullptr;
  }
  if (m->table_mode >= 2) {
    // theta returns to its place in the arena
    cudaMemcpyAsync(m->arena_theta, m->theta, (size_t)m->total * sizeof(float), cudaMemcpyDeviceToDevice, m->st);
    cudaStreamSynchronize(m->st);
    m->theta = m->arena_theta;
    for (int b = 0; b < 2; ++b) {
      comm_unmap_peers(m->table_comm, m->pp_peers[b]);
      if (m->pp[b]) cudaFree(m->pp[b]);
      m->pp[b] = nullptr;
    }
  }
  if (m->sync_blk) {
    comm_unmap_peers(m->table_comm, m->sync_peers);
    cudaFree(m->sync_blk);
    m->sync_blk = nullptr;
  }
  if (m->sync_error) { cudaFreeHost(m->sync_error); m->sync_error = nullptr; }
  m->sync_epoch = 0;
  if (m->need_r) cudaFree(m->need_r);      // one block: need_r (V words), then need_e (E words)
  m->need_r = m->need_e = nullptr;
  if (m->neg_alt) {
    // m.neg and neg_alt swap roles (look-ahead): the arena's buffer must be the one that stays
    const char *lo = m->arena, *hi = m->arena + m->arena_bytes;
    if (!(reinterpret_cast<const char *>(m->neg) >= lo && reinterpret_cast<const char *>(m->neg) < hi)) std::swap(m->neg, m->neg_alt);
    cudaFree(m->neg_alt);
    m->neg_alt = nullptr;
  }
  m->neg_presampled = false;
  m->table_own = RowOwner();
  m->table_comm = nullptr;
  m->table_mode = 0;
}

int sert_model_set_table_shard_comm(sert_model *m, sert_comm *comm, int32_t peer_stores) {
  SERT_REQUIRE(m != nullptr, "null model");
  SERT_REQUIRE(is_vs(m->cfg) && m->cfg.inference_only == 0, "table shards are for trainable vector-space models");
  if (flush_pending(*m)) return -1;
  SERT_CUDA(cudaStreamSynchronize(m->st));
  if (m->st2) SERT_CUDA(cudaStreamSynchronize(m->st2));
  table_shard_release(m);
  if (comm == nullptr) return 0;
  SERT_REQUIRE(comm->world <= kMaxPeers + 1, "table shards span at most 8 ranks (one NVLink domain)");
  SERT_REQUIRE(peer_stores >= 0 && peer_stores <= 2, "peer_stores: 0 = NCCL broadcasts, 1 = peer stores, 2 = instance shards");
  if (peer_stores == 2) {
    const sert_config &c = m->cfg;
    SERT_REQUIRE(vs_tile_supported(c.word_dim, c.entity_dim, c.window, c.num_negatives) && m->WpT != nullptr && m->st2 != nullptr,
                 "instance shards need the fused tile kernel (entity_dim 128, word_dim <= 384, window <= 32, <= 15 negatives)");
    SERT_REQUIRE(c.batch >= 8 * comm->world, "instance shards need at least one tile of 8 instances per rank");
    SERT_REQUIRE(c.vocab < (1 << 28) && c.entities < (1 << 28), "instance shards: row ids must stay below 2^28");
  }
  const long long E = m->cfg.entities, V = m->cfg.vocab;
  long long e_row[kMaxPeers + 2], r_row[kMaxPeers + 2];
  int inst_rows[kMaxPeers + 2];
  if (table_shard_partition(*m, comm->world, e_row, r_row, m->table_lo4, inst_rows)) return -1;
  m->table_own.e_lo = (int)e_row[comm->rank]; m->table_own.e_hi = (int)e_row[comm->rank + 1];
  m->table_own.r_lo = (int)r_row[comm->rank]; m->table_own.r_hi = (int)r_row[comm->rank + 1];
  // one model: every rank starts from rank 0's parameters, optimiser state and step, and draws rank 0's negatives
  if (comm_broadcast(comm, m->theta, (size_t)m->total * sizeof(float), 0, m->st)) return -1;
  const size_t state_el = m->cfg.dtype_mode == 1 ? 2 : 4;
  if (comm_broadcast(comm, m->s1, (size_t)m->total * state_el, 0, m->st)) return -1;
  if (comm_broadcast(comm, m->s2, (size_t)m->total * state_el, 0, m->st)) return -1;
  {
    unsigned long long host[3] = {(unsigned long long)m->cfg.seed, (unsigned long long)m->sample_calls,
                                  (unsigned long long)m->step};
    unsigned long long *dev = nullptr;
    SERT_CUDA(cudaMalloc(&dev, sizeof(host)));
    const bool ok = cudaMemcpyAsync(dev, host, sizeof(host), cudaMemcpyHostToDevice, m->st) == cudaSuccess &&
                    comm_broadcast(comm, dev, sizeof(host), 0, m->st) == 0 &&
                    cudaMemcpyAsync(host, dev, sizeof(host), cudaMemcpyDeviceToHost, m->st) == cudaSuccess &&
                    cudaStreamSynchronize(m->st) == cudaSuccess;
    cudaFree(dev);
    SERT_REQUIRE(ok, "table shards: could not broadcast the sampler state");
    m->cfg.seed = (decltype(m->cfg.seed))host[0];
    m->sample_calls = host[1];
    m->step = (int64_t)host[2];
  }
  m->wpt_valid = false;
  if (peer_stores) {
    SERT_CUDA(cudaMalloc(&m->need_r, (size_t)(V + E) * sizeof(uint32_t)));
    m->need_e = m->need_r + V;
    SERT_CUDA(cudaMalloc(&m->neg_alt, (size_t)m->cfg.batch * m->cfg.num_negatives * sizeof(int32_t)));
    SERT_CUDA(cudaMemsetAsync(m->need_r, 0, (size_t)(V + E) * sizeof(uint32_t), m->st));
    m->neg_presampled = false;
    m->arena_theta = m->theta;
    for (int b = 0; b < 2; ++b) {
      SERT_CUDA(cudaMalloc(&m->pp[b], (size_t)m->total * sizeof(float)));
      SERT_CUDA(cudaMemcpyAsync(m->pp[b], m->theta, (size_t)m->total * sizeof(float), cudaMemcpyDeviceToDevice, m->st));
    }
    int rc = 0;
    for (int b = 0; b < 2 && rc == 0; ++b) rc = comm_map_peers(comm, m->pp[b], m->pp_peers[b], m->st);
    if (rc) {
      for (int b = 0; b < 2; ++b) {
        comm_unmap_peers(comm, m->pp_peers[b]);
        cudaFree(m->pp[b]);
        m->pp[b] = nullptr;
      }
      return -1;
    }
    m->pp_cur = 0;
    m->theta = m->pp[0];
    const char *how = getenv("SERT_TABLE_SHARD_BARRIER");
    SERT_REQUIRE(peer_stores != 2 || !(how && strcmp(how, "nccl") == 0), "instance shards use the peer-memory barrier");
    if (!(how && strcmp(how, "nccl") == 0)) {
      SERT_CUDA(cudaMalloc(&m->sync_blk, sizeof(PeerSyncBlock)));
      SERT_CUDA(cudaMemsetAsync(m->sync_blk, 0, sizeof(PeerSyncBlock), m->st));
      SERT_CUDA(cudaHostAlloc(&m->sync_error, sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable));
      *m->sync_error = 0u;
      m->sync_epoch = 0;
      // the all-gather inside comm_map_peers orders every rank's memset before any peer's first flag store
      if (comm_map_peers(comm, m->sync_blk, m->sync_peers, m->st)) {
        cudaFree(m->sync_blk);
        m->sync_blk = nullptr;
        return -1;
      }
    }
  }
  if (peer_stores == 2) {
    // instance shards: gradients and touched stamps move into one block that every rank maps
    const size_t bytes = (size_t)m->total * sizeof(float) + (size_t)(E + V) * sizeof(uint32_t);
    SERT_CUDA(cudaMalloc(&m->gsh, bytes));
    uint32_t *fl = reinterpret_cast<uint32_t *>(m->gsh + m->total);
    SERT_CUDA(cudaMemcpyAsync(m->gsh, m->grad, (size_t)m->total * sizeof(float), cudaMemcpyDeviceToDevice, m->st));
    SERT_CUDA(cudaMemcpyAsync(fl, m->flagE, (size_t)E * sizeof(uint32_t), cudaMemcpyDeviceToDevice, m->st));
    SERT_CUDA(cudaMemcpyAsync(fl + E, m->flagR, (size_t)V * sizeof(uint32_t), cudaMemcpyDeviceToDevice, m->st));
    if (comm_map_peers(comm, m->gsh, m->gsh_peers, m->st)) {
      cudaFree(m->gsh);
      m->gsh = nullptr;
      return -1;
    }
    m->arena_grad = m->grad; m->arena_flagE = m->flagE; m->arena_flagR = m->flagR;
    for (int sg = 0; sg < m->nseg; ++sg) {
      if (m->seg[sg].flags == m->flagR) m->seg[sg].flags = fl + E;
      else if (m->seg[sg].flags == m->flagE) m->seg[sg].flags = fl;
    }
    m->grad = m->gsh; m->flagE = fl; m->flagR = fl + E;
    for (int r = 0; r <= comm->world; ++r) {
      m->inst_bound[r] = inst_rows[r];
      m->e_bound[r] = (int)e_row[r];
      m->r_bound[r] = (int)r_row[r];
    }
  }
  m->table_comm = comm;
  m->table_mode = peer_stores == 2 ? 3 : peer_stores ? 2 : 1;
  SERT_CUDA(cudaStreamSynchronize(m->st));
  return 0;
}

// Table shards: each rank's optimiser state is current only inside its own piece; this makes s1 / s2 whole on every
// rank (for a checkpoint).  Collective.
int sert_model_gather_table_state(sert_model *m) {
  SERT_REQUIRE(m != nullptr, "null model");
  if (m->table_comm == nullptr) return 0;
  if (flush_pending(*m)) return -1;
  sert_comm *c = m->table_comm;
  const size_t el = m->cfg.dtype_mode == 1 ? 2 : 4;
  size_t off[kMaxPeers + 1], len[kMaxPeers + 1];
  for (int r = 0; r < c->world; ++r) {
    off[r] = (size_t)m->table_lo4[r] * 4 * el;
    len[r] = (size_t)(m->table_lo4[r + 1] - m->table_lo4[r]) * 4 * el;
  }
  const size_t dense0 = (size_t)m->off[SERT_PARAM_DENSE_W];
  for (float *state : {m->s1, m->s2}) {
    if (comm_gather_pieces(c, state, off, len, m->st)) return -1;
    if (comm_broadcast(c, reinterpret_cast<char *>(state) + dense0 * el, ((size_t)m->total - dense0) * el, c->world - 1, m->st)) return -1;
  }
  SERT_CUDA(cudaStreamSynchronize(m->st));
  return 0;
}

int sert_model_table_shard_info(sert_model *m, int32_t *mode, int64_t *own_begin, int64_t *own_end, int64_t *table_floats) {
  SERT_REQUIRE(m != nullptr, "null model");
  if (mode) *mode = m->table_mode;
  const int r = m->table_comm ? m->table_comm->rank : 0;
  if (own_begin) *own_begin = m->table_comm ? m->table_lo4[r] * 4 : 0;
  if (own_end) *own_end = m->table_comm ? m->table_lo4[r + 1] * 4 : m->off[SERT_PARAM_DENSE_W];
  if (table_floats) *table_floats = m->off[SERT_PARAM_DENSE_W];
  return 0;
}

int sert_model_profile(sert_model *m, int enable) {
  SERT_REQUIRE(m, "null model");
  m->profile = enable != 0;
  return 0;
}

int sert_model_profile_read(sert_model *m, double *update_ms_total, int64_t *update_launches,
                            double *update_bytes_per_launch) {
  SERT_REQUIRE(m && update_ms_total && update_launches && update_bytes_per_launch, "null argument");
  SERT_CUDA(cudaStreamSynchronize(m->st));
  double tot = 0.0;
  for (auto &ev : m->prof_events) {
    float ms = 0.f;
    SERT_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
    tot += ms;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  *update_ms_total = tot;
  *update_launches = (int64_t)m->prof_events.size();
  *update_bytes_per_launch = m->prof_bytes;           // read+write of theta and two state arrays, f32
  m->prof_events.clear();
  return 0;
}

int sert_model_attach_dataset(sert_model *m, int split, int64_t n, const int32_t *x_dev, const int32_t *y_dev,
                              const int64_t *indptr_dev, const int32_t *indices_dev, const float *data_dev,
                              const float *w_dev) {
  SERT_REQUIRE(m, "null model");
  SERT_REQUIRE(split == SERT_SPLIT_TRAIN || split == SERT_SPLIT_VALIDATE, "unknown split");
  SERT_REQUIRE(n >= 0, "negative instance count");
  SERT_REQUIRE(n == 0 || x_dev != nullptr, "x is required");
  if (is_vs(m->cfg)) {
    // sert/models.py:933-934: 'Only one-hot vectors supported.'
    SERT_REQUIRE(n == 0 || y_dev != nullptr, "Only one-hot vectors supported.");
  } else {
    SERT_REQUIRE(n == 0 || (indptr_dev && indices_dev && data_dev), "log-linear model needs CSR labels");
  }
  Dataset &d = m->ds[split];
  d.n = n; d.x = x_dev; d.y = y_dev; d.indptr = indptr_dev; d.indices = indices_dev; d.data = data_dev; d.w = w_dev;
  return 0;
}

int sert_train_batches(sert_model *m, const int64_t *order_host, int64_t n, const int32_t *neg_dev,
                       int32_t first_slot) {
  SERT_REQUIRE(m && (order_host || n == 0), "null argument");
  SERT_REQUIRE(m->cfg.inference_only == 0, "model was created inference_only: it cannot be trained");
  SERT_REQUIRE(first_slot >= 0 && first_slot + n <= m->cfg.loss_slots, "loss slots exhausted");
  const sert_config &c = m->cfg;
  const Dataset &d = m->ds[SERT_SPLIT_TRAIN];
  for (int64_t j = 0; j < n; ++j) {
    const int64_t b = order_host[j];
    if (check_batch(m, SERT_SPLIT_TRAIN, b)) return -1;
    const int64_t r0 = b * c.batch;
    const float *w = d.w ? d.w + r0 : nullptr;
    float *loss = m->losses + first_slot + j;
    int rc;
    if (is_vs(c)) {
      const int32_t *neg = neg_dev ? neg_dev + j * (int64_t)c.batch * c.num_negatives : nullptr;
      NextBatch next;                      // table shards: the batch after this one is known here
      if (m->table_comm != nullptr && j + 1 < n) {
        if (check_batch(m, SERT_SPLIT_TRAIN, order_host[j + 1])) return -1;
        const int64_t r1 = order_host[j + 1] * c.batch;
        next.x = d.x + r1 * c.window;
        next.y = d.y + r1;
        next.neg = neg_dev ? neg_dev + (j + 1) * (int64_t)c.batch * c.num_negatives : nullptr;
        next.sampled = neg_dev == nullptr;
      }
      rc = vs_train_step(*m, d.x + r0 * c.window, d.y + r0, w, neg, loss, next);
    } else {
      rc = (m->exchange ? ll_train_step_sharded : ll_train_step)(*m, d.x + r0 * c.window, d.indptr + r0, 0,
                                                                d.indices, d.data, w, loss);
    }
    if (rc) return -1;
  }
  if (flush_pending(*m)) return -1;
  return m->exchange ? ll_reduce_losses(*m, first_slot, n) : 0;
}

int sert_eval_batches(sert_model *m, int split, const int64_t *order_host, int64_t n, const int32_t *neg_dev,
                      int32_t first_slot) {
  SERT_REQUIRE(m && (order_host || n == 0), "null argument");
  SERT_REQUIRE(first_slot >= 0 && first_slot + n <= m->cfg.loss_slots, "loss slots exhausted");
  const sert_config &c = m->cfg;
  if (flush_pending(*m)) return -1;
  for (int64_t j = 0; j < n; ++j) {
    const int64_t b = order_host[j];
    if (check_batch(m, split, b)) return -1;
    const Dataset &d = m->ds[split];
    const int64_t r0 = b * c.batch;
    float *loss = m->losses + first_slot + j;
    int rc;
    if (is_vs(c)) {
      const int32_t *neg = neg_dev ? neg_dev + j * (int64_t)c.batch * c.num_negatives : nullptr;
      rc = vs_eval_step(*m, d.x + r0 * c.window, d.y + r0, neg, loss, false);
    } else {
      rc = (m->exchange ? ll_eval_step_sharded : ll_eval_step)(*m, d.x + r0 * c.window, d.indptr + r0, 0,
                                                              d.indices, d.data, loss, false);
    }
    if (rc) return -1;
  }
  return m->exchange ? ll_reduce_losses(*m, first_slot, n) : 0;
}

int sert_losses_fetch(sert_model *m, int32_t first_slot, int64_t n, float *out_host) {
  SERT_REQUIRE(m && (out_host || n == 0), "null argument");
  SERT_REQUIRE(first_slot >= 0 && first_slot + n <= m->cfg.loss_slots, "loss slot range out of bounds");
  if (flush_pending(*m)) return -1;
  if (n > 0)
    SERT_CUDA(cudaMemcpyAsync(out_host, m->losses + first_slot, n * sizeof(float), cudaMemcpyDeviceToHost, m->st));
  SERT_CUDA(cudaStreamSynchronize(m->st));
  if (peer_barrier_check(*m)) return -1;
  for (int64_t j = 0; j < n; ++j) {
    if (!isfinite(out_host[j])) {
      // message mirrors sert/models.py:372-379
      char buf[256];
      snprintf(buf, sizeof(buf),
               "Encountered NaN or infinity (%f) during batch iteration (batch %lld/%lld).",
               (double)out_host[j], (long long)(j + 1), (long long)n);
      set_error(buf);
      return -2;
    }
  }
  return 0;
}

int sert_train_batch_host(sert_model *m, const int32_t *x_host, const int32_t *y_host, const int64_t *indptr_host,
                          const int32_t *indices_host, const float *data_host, const float *w_host,
                          const int32_t *neg_host, float *loss_host) {
  SERT_REQUIRE(m && x_host && loss_host, "null argument");
  SERT_REQUIRE(m->cfg.inference_only == 0, "model was created inference_only: it cannot be trained");
  const sert_config &c = m->cfg;
  cudaStream_t st = m->st;
  const size_t B = c.batch;
  SERT_CUDA(cudaMemcpyAsync(m->stage_x, x_host, B * c.window * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  const float *w = nullptr;
  if (w_host) {
    SERT_CUDA(cudaMemcpyAsync(m->stage_w, w_host, B * sizeof(float), cudaMemcpyHostToDevice, st));
    w = m->stage_w;
  }
  float *loss = m->losses + c.loss_slots;   // scratch slot
  if (is_vs(c)) {
    SERT_REQUIRE(y_host, "vector-space batches need one-hot labels");
    SERT_CUDA(cudaMemcpyAsync(m->stage_y, y_host, B * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    const int32_t *neg = nullptr;
    if (neg_host) {
      SERT_CUDA(cudaMemcpyAsync(m->stage_neg, neg_host, B * c.num_negatives * sizeof(int32_t),
                                cudaMemcpyHostToDevice, st));
      neg = m->stage_neg;
    }
    if (vs_train_step(*m, m->stage_x, m->stage_y, w, neg, loss)) return -1;
    if (flush_pending(*m)) return -1;
  } else {
    SERT_REQUIRE(indptr_host && indices_host && data_host, "log-linear batches need CSR labels");
    const int64_t base = indptr_host[0];
    const size_t nnz = (size_t)(indptr_host[B] - base);
    SERT_REQUIRE(nnz <= m->stage_nnz_cap, "too many labels in one host batch");
    SERT_CUDA(cudaMemcpyAsync(m->stage_indptr, indptr_host, (B + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    SERT_CUDA(cudaMemcpyAsync(m->stage_indices, indices_host + base, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    SERT_CUDA(cudaMemcpyAsync(m->stage_data, data_host + base, nnz * sizeof(float), cudaMemcpyHostToDevice, st));
    if ((m->exchange ? ll_train_step_sharded : ll_train_step)(*m, m->stage_x, m->stage_indptr, base,
                                                              m->stage_indices, m->stage_data, w, loss))
      return -1;
    if (m->exchange && ll_reduce_losses(*m, c.loss_slots, 1)) return -1;
  }
  SERT_CUDA(cudaMemcpyAsync(loss_host, loss, sizeof(float), cudaMemcpyDeviceToHost, st));
  SERT_CUDA(cudaStreamSynchronize(st));
  if (!isfinite(*loss_host)) {
    char buf[160];
    snprintf(buf, sizeof(buf), "Encountered NaN or infinity (%f) during batch iteration.", (double)*loss_host);
    set_error(buf);
    return -2;
  }
  return 0;
}

// ---- pipelined host batches ------------------------------------------------------------------------------------
static int pipe_setup(sert_model &m) {
  if (m.pipe_ready) return 0;
  SERT_CUDA(cudaStreamCreateWithFlags(&m.st_h2d, cudaStreamNonBlocking));
  SERT_CUDA(cudaStreamCreateWithFlags(&m.st_d2h, cudaStreamNonBlocking));
  for (int q = 0; q < 2; ++q) {
    SERT_CUDA(cudaEventCreateWithFlags(&m.ev_copied[q], cudaEventDisableTiming));
    SERT_CUDA(cudaEventCreateWithFlags(&m.ev_free[q], cudaEventDisableTiming));
  }
  SERT_CUDA(cudaEventCreateWithFlags(&m.ev_fused, cudaEventDisableTiming));
  for (int q = 0; q < sert_model::kPipe; ++q) SERT_CUDA(cudaEventCreateWithFlags(&m.ev_loss[q], cudaEventDisableTiming));
  SERT_CUDA(cudaMallocHost(reinterpret_cast<void **>(&m.pin_loss), sert_model::kPipe * sizeof(float)));
  m.pipe_ready = true;
  return 0;
}

// loss of ticket t: device slot -> pinned host slot on the D2H stream, once `after` has happened
static int pipe_dispatch_loss(sert_model &m, int64_t t, cudaEvent_t after) {
  const int q = (int)(t % sert_model::kPipe);
  SERT_CUDA(cudaStreamWaitEvent(m.st_d2h, after, 0));
  SERT_CUDA(cudaMemcpyAsync(m.pin_loss + q, m.pipe_losses + q, sizeof(float), cudaMemcpyDeviceToHost, m.st_d2h));
  SERT_CUDA(cudaEventRecord(m.ev_loss[q], m.st_d2h));
  return 0;
}

int sert_train_batch_host_async(sert_model *m, const int32_t *x_host, const int32_t *y_host, const float *w_host,
                                const int32_t *neg_host, int64_t *ticket_out) {
  SERT_REQUIRE(m && x_host && y_host && ticket_out, "null argument");
  SERT_REQUIRE(is_vs(m->cfg) && m->cfg.inference_only == 0, "pipelined host batches need a trainable vector-space model");
  if (pipe_setup(*m)) return -1;
  const sert_config &c = m->cfg;
  const size_t B = c.batch;
  const int64_t t = m->host_steps;
  const int set = (int)(t & 1), q = (int)(t % sert_model::kPipe);
  // the ring slot of ticket t - kPipe is reused: its loss must have reached the host
  if (t >= sert_model::kPipe && m->host_unfetched != t - sert_model::kPipe) SERT_CUDA(cudaEventSynchronize(m->ev_loss[q]));
  SERT_REQUIRE(m->host_unfetched < 0 || m->host_unfetched > t - sert_model::kPipe, "loss ring overrun");
  int32_t *sx = set ? m->stage2_x : m->stage_x, *sy = set ? m->stage2_y : m->stage_y;
  int32_t *sn = set ? m->stage2_neg : m->stage_neg;
  float *sw = set ? m->stage2_w : m->stage_w;
  // host -> device on the copy stream, into the staging set the step before last has released
  if (m->ev_free_set[set]) SERT_CUDA(cudaStreamWaitEvent(m->st_h2d, m->ev_free[set], 0));
  SERT_CUDA(cudaMemcpyAsync(sx, x_host, B * c.window * sizeof(int32_t), cudaMemcpyHostToDevice, m->st_h2d));
  SERT_CUDA(cudaMemcpyAsync(sy, y_host, B * sizeof(int32_t), cudaMemcpyHostToDevice, m->st_h2d));
  if (w_host) SERT_CUDA(cudaMemcpyAsync(sw, w_host, B * sizeof(float), cudaMemcpyHostToDevice, m->st_h2d));
  if (neg_host)
    SERT_CUDA(cudaMemcpyAsync(sn, neg_host, B * c.num_negatives * sizeof(int32_t), cudaMemcpyHostToDevice, m->st_h2d));
  SERT_CUDA(cudaEventRecord(m->ev_copied[set], m->st_h2d));
  SERT_CUDA(cudaStreamWaitEvent(m->st, m->ev_copied[set], 0));
  m->want_fused_event = true;
  const int rc = vs_train_step(*m, sx, sy, w_host ? sw : nullptr, neg_host ? sn : nullptr, m->pipe_losses + q);
  m->want_fused_event = false;
  if (rc) return -1;
  SERT_CUDA(cudaEventRecord(m->ev_free[set], m->st));
  m->ev_free_set[set] = true;
  // the loss of the ticket before this one is final behind this step's forward/backward kernel
  if (m->host_unfetched >= 0 && pipe_dispatch_loss(*m, m->host_unfetched, m->ev_fused)) return -1;
  m->host_unfetched = -1;
  if (m->pending_bank >= 0) {
    m->host_unfetched = t;                      // written by the next step's tile kernel (or sert_train_host_wait)
  } else {
    if (pipe_dispatch_loss(*m, t, m->ev_free[set])) return -1;
  }
  *ticket_out = t;
  m->host_steps = t + 1;
  return 0;
}

int sert_train_host_wait(sert_model *m, int64_t ticket, float *loss_host) {
  SERT_REQUIRE(m && loss_host, "null argument");
  SERT_REQUIRE(ticket >= 0 && ticket < m->host_steps && ticket >= m->host_steps - sert_model::kPipe,
               "ticket out of the window of the last 8 pipelined steps");
  if (ticket == m->host_unfetched) {
    if (flush_pending(*m)) return -1;
    SERT_CUDA(cudaEventRecord(m->ev_fused, m->st));
    if (pipe_dispatch_loss(*m, ticket, m->ev_fused)) return -1;
    m->host_unfetched = -1;
  }
  const int q = (int)(ticket % sert_model::kPipe);
  SERT_CUDA(cudaEventSynchronize(m->ev_loss[q]));
  *loss_host = m->pin_loss[q];
  if (!isfinite(*loss_host)) {
    char buf[160];
    snprintf(buf, sizeof(buf), "Encountered NaN or infinity (%f) during batch iteration.", (double)*loss_host);
    set_error(buf);
    return -2;
  }
  return 0;
}

int sert_vs_forward_host(sert_model *m, int split, int64_t batch_index, const int32_t *neg_dev,
                         float *out_scores_host, float *out_proj_host, float *out_ell_host) {
  SERT_REQUIRE(m && is_vs(m->cfg), "vector-space model required");
  SERT_REQUIRE(neg_dev, "parity forward needs explicit negatives");
  if (check_batch(m, split, batch_index)) return -1;
  const sert_config &c = m->cfg;
  const Dataset &d = m->ds[split];
  const int64_t r0 = batch_index * c.batch;
  if (vs_eval_step(*m, d.x + r0 * c.window, d.y + r0, neg_dev, m->losses + c.loss_slots, true)) return -1;
  const size_t B = c.batch;
  if (out_scores_host)
    SERT_CUDA(cudaMemcpyAsync(out_scores_host, m->dbg_scores, B * (c.num_negatives + 1) * sizeof(float),
                              cudaMemcpyDeviceToHost, m->st));
  if (out_proj_host)
    SERT_CUDA(cudaMemcpyAsync(out_proj_host, m->dbg_u, B * c.entity_dim * sizeof(float), cudaMemcpyDeviceToHost, m->st));
  if (out_ell_host)
    SERT_CUDA(cudaMemcpyAsync(out_ell_host, m->dbg_ell, B * sizeof(float), cudaMemcpyDeviceToHost, m->st));
  SERT_CUDA(cudaStreamSynchronize(m->st));
  return 0;
}

int sert_ll_forward_host(sert_model *m, int split, int64_t batch_index, float *out_z_host, float *out_s_host,
                         float *out_ell_host) {
  SERT_REQUIRE(m && !is_vs(m->cfg), "log-linear model required");
  if (check_batch(m, split, batch_index)) return -1;
  const sert_config &c = m->cfg;
  const Dataset &d = m->ds[split];
  const int64_t r0 = batch_index * c.batch;
  if ((m->exchange ? ll_eval_step_sharded : ll_eval_step)(*m, d.x + r0 * c.window, d.indptr + r0, 0, d.indices,
                                                          d.data, m->losses + c.loss_slots, true))
    return -1;
  const size_t B = c.batch, E = c.entities;
  // sharded: z and s are this rank's columns; the instance losses are completed over the shards
  if (m->exchange && out_ell_host && xchg(*m, SERT_XCHG_ALLREDUCE_SUM, m->dbg_ell, B)) return -1;
  if (out_z_host)
    SERT_CUDA(cudaMemcpyAsync(out_z_host, m->Z, B * c.window * E * sizeof(float), cudaMemcpyDeviceToHost, m->st));
  if (out_s_host) SERT_CUDA(cudaMemcpyAsync(out_s_host, m->S, B * E * sizeof(float), cudaMemcpyDeviceToHost, m->st));
  if (out_ell_host)
    SERT_CUDA(cudaMemcpyAsync(out_ell_host, m->dbg_ell, B * sizeof(float), cudaMemcpyDeviceToHost, m->st));
  SERT_CUDA(cudaStreamSynchronize(m->st));
  return 0;
}

int sert_predict_loglinear(sert_model *m, const int32_t *batch_host, int32_t rows, float *out_host) {
  SERT_REQUIRE(m && batch_host && out_host, "null argument");
  SERT_REQUIRE(!is_vs(m->cfg), "log-linear model required");
  const sert_config &c = m->cfg;
  SERT_REQUIRE(rows >= 0 && rows <= c.batch, "more rows than the model's batch size");
  SERT_REQUIRE(m->exchange == nullptr, "predict_fn needs the full entity axis: gather the shards first");
  if (rows == 0) return 0;
  const size_t BW = (size_t)rows * c.window, E = c.entities;
  SERT_CUDA(cudaMemcpyAsync(m->stage_x, batch_host, BW * sizeof(int32_t), cudaMemcpyHostToDevice, m->st));
  if (ll_forward(*m, m->stage_x, rows, m->st)) return -1;
  if (launch_ll_softmax_inplace(m->Z, BW, (int)E, E, m->rmax, m->rsum, m->st)) return -1;
  SERT_CUDA(cudaMemcpyAsync(out_host, m->Z, BW * E * sizeof(float), cudaMemcpyDeviceToHost, m->st));
  SERT_CUDA(cudaStreamSynchronize(m->st));
  return 0;
}

// scratch of the ranking calls: grown on demand, owned by the library
static int rank_scratch(void **buf, size_t *cap, size_t need) {
  if (*cap >= need) return 0;
  if (*buf) SERT_CUDA(cudaFree(*buf));
  *buf = nullptr; *cap = 0;
  SERT_CUDA(cudaMalloc(buf, need));
  *cap = need;
  return 0;
}

int sert_ll_rank_queries(sert_model *m, const int32_t *batch_host, int32_t rows, const int32_t *query_first_row_host,
                         const int32_t *query_terms_host, int32_t num_queries, int32_t top, int32_t *out_idx_host,
                         float *out_rel_host, float *out_term_entropy_host, float *out_entropy_host,
                         float *out_mass_host) {
  SERT_REQUIRE(m && batch_host && query_first_row_host && query_terms_host && out_idx_host && out_rel_host,
               "null argument");
  SERT_REQUIRE(!is_vs(m->cfg), "log-linear model required");
  const sert_config &c = m->cfg;
  SERT_REQUIRE(rows >= 0 && rows <= c.batch, "more rows than the model's batch size");
  SERT_REQUIRE(m->exchange == nullptr, "ranking needs the full entity axis: gather the shards first");
  SERT_REQUIRE(num_queries >= 0 && num_queries <= rows, "more queries than rows");
  if (num_queries == 0) return 0;
  const int W = c.window, E = (int)c.entities;
  std::vector<int32_t> first((size_t)num_queries);
  long long n_terms = 0;
  for (int j = 0; j < num_queries; ++j) {
    const long long r0 = query_first_row_host[j], T = query_terms_host[j];
    SERT_REQUIRE(r0 >= 0 && T >= 1 && r0 * W + T <= (long long)rows * W, "query outside the batch");
    first[j] = (int32_t)(r0 * W);          // term t of the query is row first + t of the (rows*W, E) per-term matrix
    n_terms += T;
  }
  const size_t BW = (size_t)rows * W;
  SERT_CUDA(cudaMemcpyAsync(m->stage_x, batch_host, BW * sizeof(int32_t), cudaMemcpyHostToDevice, m->st));
  if (ll_forward(*m, m->stage_x, rows, m->st)) return -1;
  const size_t need = ll_rank_scratch_bytes(num_queries, E, (int)n_terms);
  if (rank_scratch(&m->rank_scratch, &m->rank_scratch_bytes, need)) return -1;
  return ll_rank(m->Z, m->rmax, m->rsum, false, E, E, first.data(), query_terms_host, num_queries, top, m->S,
                 m->rank_scratch, m->rank_scratch_bytes, out_idx_host, out_rel_host, out_term_entropy_host,
                 out_entropy_host, out_mass_host, m->st);
}

int sert_ll_rank_distributions(const float *dist_host, int32_t num_terms, int32_t entities,
                               const int32_t *query_first_term_host, const int32_t *query_terms_host,
                               int32_t num_queries, int32_t top, int32_t *out_idx_host, float *out_rel_host,
                               float *out_term_entropy_host, float *out_entropy_host, float *out_mass_host) {
  SERT_REQUIRE(dist_host && query_first_term_host && query_terms_host && out_idx_host && out_rel_host, "null argument");
  SERT_REQUIRE(num_terms >= 1 && entities >= 1 && num_queries >= 0, "bad shape");
  int ndev = 0;
  SERT_CUDA(cudaGetDeviceCount(&ndev));
  SERT_REQUIRE(ndev > 0, "no CUDA device: libsert_b200 has no CPU fallback");
  if (num_queries == 0) return 0;
  for (int j = 0; j < num_queries; ++j)
    SERT_REQUIRE(query_first_term_host[j] >= 0 && query_terms_host[j] >= 1 &&
                 query_first_term_host[j] + query_terms_host[j] <= num_terms, "query outside the term matrix");
  float *dist = nullptr, *rel = nullptr;
  void *scratch = nullptr;
  const size_t need = ll_rank_scratch_bytes(num_queries, entities, num_terms);
  SERT_CUDA(cudaMalloc(&dist, (size_t)num_terms * entities * sizeof(float)));
  SERT_CUDA(cudaMalloc(&rel, (size_t)num_queries * entities * sizeof(float)));
  SERT_CUDA(cudaMalloc(&scratch, need));
  SERT_CUDA(cudaMemcpy(dist, dist_host, (size_t)num_terms * entities * sizeof(float), cudaMemcpyHostToDevice));
  const int rc = ll_rank(dist, nullptr, nullptr, true, entities, entities, query_first_term_host, query_terms_host,
                         num_queries, top, rel, scratch, need, out_idx_host, out_rel_host, out_term_entropy_host,
                         out_entropy_host, out_mass_host, nullptr);
  cudaFree(dist); cudaFree(rel); cudaFree(scratch);
  return rc;
}

int sert_project_queries(sert_model *m, const float *avg_host, int32_t q, float *out_host) {
  SERT_REQUIRE(m && avg_host && out_host, "null argument");
  SERT_REQUIRE(is_vs(m->cfg), "vector-space model required");
  const sert_config &c = m->cfg;
  const float *Wp = m->theta + m->off[SERT_PARAM_DENSE_W];
  const float *bp = m->theta + m->off[SERT_PARAM_DENSE_B];
  for (int32_t q0 = 0; q0 < q; q0 += c.batch) {
    const int n = std::min<int32_t>(c.batch, q - q0);
    SERT_CUDA(cudaMemcpyAsync(m->h, avg_host + (size_t)q0 * c.word_dim, (size_t)n * c.word_dim * sizeof(float),
                              cudaMemcpyHostToDevice, m->st));
    if (launch_gemm_f32(m->h, Wp, m->t, n, c.entity_dim, c.word_dim, false, false, c.word_dim, c.entity_dim,
                        c.entity_dim, EPI_BIAS_TANH, bp, 1, m->st))
      return -1;
    SERT_CUDA(cudaMemcpyAsync(out_host + (size_t)q0 * c.entity_dim, m->t, (size_t)n * c.entity_dim * sizeof(float),
                              cudaMemcpyDeviceToHost, m->st));
    SERT_CUDA(cudaStreamSynchronize(m->st));
  }
  return 0;
}

}  // extern "C"
