/*
 * sert_b200.h -- C-ABI of the B200-native SERT hot path (libsert_b200.so).
 *
 * The reference (cvangysel/SERT) has no FFI: its hot path is a Theano graph
 * reached through the Python class surface of sert/models.py and four compiled
 * callables (train_fn / test_fn / validate_fn / predict_fn).  Every entry point
 * below replaces one of those callables or the numpy/sklearn ranking code of
 * bin/query.py; the reference interface it stands in for is cited (file:line,
 * relative to the reference checkout).  The Python mirror of the class surface
 * (sert_b200/models.py, sert_b200/inference.py, sert_b200/ranking.py) binds these
 * symbols with ctypes (sert_b200/_native.py); INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in signatures.  `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every function returns 0 on success, non-zero on failure;
 *     sert_last_error() returns the message of the calling thread's last failure
 *     (the Python layer raises RuntimeError with it, matching the reference's
 *     error behaviour, sert/models.py:372-379,624-628).
 *   - "dev" pointers are device (HBM) addresses owned by the caller (the Python
 *     layer allocates them as torch tensors: torch is the HBM allocator, nothing
 *     else); "host" pointers are ordinary host memory (pinned memory makes the
 *     copies asynchronous).
 *   - one host thread per model; calls enqueue work on the model's stream and only
 *     the *_fetch / *_host / get_* calls synchronise.
 *   - all floating point is IEEE float32 (the reference aborts on float64,
 *     sert/models.py:610-628); indices are int32 on the device.
 */
#ifndef SERT_B200_H_
#define SERT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SERT_ABI_VERSION 1

#if defined(__GNUC__)
#define SERT_API __attribute__((visibility("default")))
#else
#define SERT_API
#endif

/* model kinds: bin/train.py:21-24 MODELS */
#define SERT_KIND_LOGLINEAR   0   /* sert.models.LanguageModel          (sert/models.py:804)  */
#define SERT_KIND_VECTORSPACE 1   /* sert.models.VectorSpaceLanguageModel (sert/models.py:1024) */

/* dataset splits: ModelInterface.TRAIN / VALIDATE (sert/models.py:312) */
#define SERT_SPLIT_TRAIN    0
#define SERT_SPLIT_VALIDATE 1

/* parameter tensor ids for get/set */
#define SERT_PARAM_WORD_REPR   0  /* R    (V,dw)   SparseProjectionLayer.representations, sert/models.py:167-171 */
#define SERT_PARAM_DENSE_W     1  /* W    (dw,E) log-linear | (dw,de) vector space; DenseLayer.W, :846-849,1057-1061 */
#define SERT_PARAM_DENSE_B     2  /* b    (E,) | (de,)      DenseLayer.b */
#define SERT_PARAM_ENTITY_REPR 3  /* Eemb (E,de)  "Class representations", sert/models.py:940-941 (vector space only) */
/* optimiser state slots (Adadelta: accu, delta_accu; Adam: m, v) */
#define SERT_STATE_PARAM 0
#define SERT_STATE_S1    1
#define SERT_STATE_S2    2

typedef struct sert_config {
  int32_t kind;            /* SERT_KIND_* */
  int32_t batch;           /* B: --batch_size, bin/train.py:41-42 */
  int32_t window;          /* W: data_args.window_size, bin/prepare.py:54 */
  int32_t num_negatives;   /* k: --num_negative_samples (vector space), bin/train.py:50-51 */
  int64_t vocab;           /* V */
  int64_t entities;        /* E */
  int32_t word_dim;        /* dw: --word_representation_size */
  int32_t entity_dim;      /* de: --entity_representation_size (vector space) */
  float   lambda;          /* --regularization_lambda, bin/train.py:58-59 */
  int32_t loss_slots;      /* capacity of the device-side per-batch loss buffer */
  uint64_t seed;           /* negative-sampler seed (reference: unseeded RandomStreams, sert/models.py:958-959) */
  int32_t inference_only;  /* 1: parameters + forward workspaces only (predict_fn models); training calls are refused */
  int32_t dtype_mode;      /* 0: float32 throughout -- the reference's arithmetic and the parity mode (float64 anywhere is
                              a hard error there, sert/models.py:624-628).  1: perf mode of BASELINE.json configs[1] ("bf16"):
                              the two optimiser-state arrays of every parameter (Adam m, v / Adadelta accu, delta_accu;
                              sert/models.py:820,922) are stored as bfloat16 with stochastic rounding, parameters and
                              gradients stay float32: the dense update streams 16 instead of 24 bytes per parameter */
  int64_t reserved1;       /* must be 0 */
} sert_config;

typedef struct sert_model sert_model;     /* opaque */
typedef struct sert_scorer sert_scorer;   /* opaque */
typedef struct sert_comm sert_comm;       /* opaque: one NCCL communicator (one rank = one process = one GPU) */

/* ---- library --------------------------------------------------------------------------- */
SERT_API int         sert_abi_version(void);
SERT_API const char *sert_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
SERT_API uint64_t    sert_launch_count(void);

/* ---- communicator (no reference counterpart: the reference is single-device) --------
 * The collectives of the sharded paths -- the ONE all-gather of per-shard top-k lists of row-sharded scoring and the
 * five small exchanges of an entity-sharded log-linear step -- are NCCL calls issued by the library on the stream of
 * the scorer / model.  The host ships nothing but the 128-byte unique id: rank 0 calls sert_comm_unique_id, every
 * rank receives the bytes over any channel (sert_b200/comm.py uses torch.distributed's store) and calls
 * sert_comm_init on its own device (cudaSetDevice first; collective: returns when all ranks have joined).
 * libnccl.so.2 is loaded on first use (the copy already in the process when torch is imported). */
SERT_API int sert_comm_unique_id(void *id_out, size_t capacity /* >= 128 */);
SERT_API int sert_comm_init(int32_t rank, int32_t world, const void *unique_id, sert_comm **out);
SERT_API int sert_comm_destroy(sert_comm *c);
/* rank, world, NCCL version code, collectives issued and payload bytes so far (any pointer may be NULL) */
SERT_API int sert_comm_info(sert_comm *c, int32_t *rank, int32_t *world, int32_t *nccl_version, int64_t *collectives,
                            int64_t *bytes);

/* ---- model life cycle: replaces the model constructors, sert/models.py:806-878,1026-1105 -- */
/* HBM bytes the model needs for parameters, optimiser state, gradients and workspaces. */
SERT_API int sert_model_arena_bytes(const sert_config *cfg, size_t *bytes);
/* `arena_dev` must hold sert_model_arena_bytes() bytes, 256-byte aligned; it is zero-filled by create. */
SERT_API int sert_model_create(const sert_config *cfg, void *arena_dev, size_t arena_bytes, void *stream,
                      sert_model **out);
SERT_API int sert_model_destroy(sert_model *m);

/* Parameter / optimiser-state transfer (host float32, row-major).  Replaces representations_init /
 * entity_representations_init (sert/models.py:699-713,940-941) and get_representations()/get_state()
 * (sert/models.py:670-680,797-801,943-945).  `which` = SERT_PARAM_*, `slot` = SERT_STATE_*. */
SERT_API int sert_model_set_tensor(sert_model *m, int which, int slot, const float *host, size_t count);
SERT_API int sert_model_get_tensor(sert_model *m, int which, int slot, float *host, size_t count);
/* Adam's shared step counter t (lasagne.updates.adam t_prev; sert/models.py:922). */
SERT_API int sert_model_set_step(sert_model *m, int64_t t);
SERT_API int sert_model_get_step(sert_model *m, int64_t *t);

/* State of the device-side negative sampler (sert/models.py:947-979 draws from an unseeded RandomStreams): the
 * Philox seed and the number of draws made so far.  A checkpoint that carries both resumes with the negatives an
 * uninterrupted run would have drawn. */
SERT_API int sert_model_get_sampler(sert_model *m, uint64_t *seed, uint64_t *draws);
SERT_API int sert_model_set_sampler(sert_model *m, uint64_t seed, uint64_t draws);

/* Measurement hook (no reference counterpart): when enabled, every training step brackets its dense
 * optimiser kernel with CUDA events on the model's stream -- in the two-stream vector-space step that is the
 * streaming kernel over the two tables, timed in situ (the small kernels of the second stream run beside it as
 * they do in an unprofiled step).  profile_read synchronises and returns the summed kernel time, the launch count
 * and the ALGORITHMIC bytes of one launch (24 B per parameter the launch covers: read+write of theta and the two
 * optimiser-state arrays; DESIGN.md "roofline"). */
SERT_API int sert_model_profile(sert_model *m, int enable);
/* Vector-space training step: 1 (default) = fused forward+backward kernel -- one CTA per tile of 8 instances when both
 * representation sizes are 128 (csrc/vs_tile.cu), else one warp per pair of instances for sizes up to 384
 * (csrc/vs_warp.cu); 2 = always the warp kernel; 0 = one kernel per stage (any shape; what the fused kernels are
 * tested against). */
SERT_API int sert_model_set_fused(sert_model *m, int enable);
/* Vector-space training step, fused tile kernel: word ids that occur very often in every batch ("the", "of", ...).
 * Thousands of gradient-row additions per step on one row serialise in the L2 slices that own its four lines
 * (tools/red_probe.cu), so additions to these rows are spread over 16 private copies and folded into the gradient
 * row by a small kernel behind the step.  Result-neutral up to float summation order.  n <= 32; n = 0 turns it
 * off.  The Python model picks the ids from the training set's word counts (sert_b200/models.py). */
SERT_API int sert_model_set_hot_words(sert_model *m, const int32_t *ids_host, int32_t n);
/* Vector-space training step: 1 (default) = the two small dense-gradient kernels (gW = h^T.da, gb = colsum(da))
 * run on a second stream concurrently with the Adam stream over the tables; 0 = everything on one stream. */
SERT_API int sert_model_set_overlap(sert_model *m, int enable);
/* Log-linear word x entity GEMMs (sert/models.py:846-849 and its two gradients): 1 (default) = tcgen05 tensor
 * cores on bf16x3-split operands whenever the output has >= 32 tiles of 128x256, 0 = fp32 FMA tiles always. */
SERT_API int sert_model_set_tensor_cores(sert_model *m, int enable);
SERT_API int sert_model_profile_read(sert_model *m, double *update_ms_total, int64_t *update_launches,
                                     double *update_bytes_per_launch);

/* ---- entity-sharded log-linear training (no reference counterpart: the reference is single-device; SURVEY.md
 * 8(e) "column-parallel softmax") ----------------------------------------------------------------------
 * A model created with cfg.entities = E_loc owns the columns [entity_begin, entity_begin+E_loc) of the dense
 * layer W (dw,E) / b (E,) (sert/models.py:846-849) and of the logits; the word table R is replicated and gets
 * the identical update on every rank.  CSR label indices stay GLOBAL entity ids.  The five exchange points of a
 * step (row statistics of the two softmaxes, two per-row sums of the backward, the partial dX) call `fn` with a
 * device buffer; the host performs the collective in place, ordered on the model's stream (torch.distributed /
 * NCCL in sert_b200/sharding.py).  Per-batch losses are completed by one more all-reduce over the loss slots.
 *   SERT_XCHG_ALLREDUCE_SUM: buf[0..count) <- sum over ranks.
 *   SERT_XCHG_ALLGATHER:     buf holds world blocks of `count` floats; rank r filled block r; all blocks are
 *                            gathered on every rank.
 * `fn` returns 0 on success. */
#define SERT_XCHG_ALLREDUCE_SUM 0
#define SERT_XCHG_ALLGATHER     1
typedef int (*sert_exchange_fn)(void *ctx, int32_t op, float *buf_dev, size_t count);
SERT_API int sert_model_set_entity_shard(sert_model *m, int32_t rank, int32_t world, int64_t entity_begin,
                                         int64_t entities_total, sert_exchange_fn fn, void *ctx);

/* The same sharding with the five exchanges issued by the library as NCCL collectives on the model's stream
 * (rank / world come from the communicator); the callback form above remains for gloo / single-device tests. */
SERT_API int sert_model_set_entity_shard_comm(sert_model *m, sert_comm *comm, int64_t entity_begin,
                                              int64_t entities_total);

/* ---- table-sharded vector-space training ("shard rows of R and Eemb (+Adam state) across ranks") --
 * One model at the global batch, `world` <= 8 ranks of one NVLink domain.  Every rank is fed the SAME batches (and
 * draws the same negatives: same sert_config.seed) and computes the whole step's gradient, but streams the Adam
 * update only over its own contiguous piece of the two representation tables, 1/world of the 24 B/parameter dense
 * update (the last rank also updates the projection matrix and bias).  The new parameters then reach every rank:
 *   peer_stores = 0: in place, by grouped ncclBroadcast of the pieces behind the update kernels;
 *   peer_stores = 1: by the update kernels' own stores into the next of two parameter buffers of every rank, mapped
 *                    with CUDA IPC over NVLink (the update is the exchange); the 64 partial sums of theta^2 of the
 *                    loss travel through a barrier kernel over the same peer mappings (csrc/peer_sync.cu) behind which
 *                    the buffers swap -- no NCCL call on the step (SERT_TABLE_SHARD_BARRIER=nccl: one ncclAllReduce).
 *   peer_stores = 2: "instance shards": as 1, and every rank runs the forward / backward of its own slice of the
 *                    batch's instances only (fused tile kernel shapes: entity_dim 128, word_dim <= 384), adding each
 *                    gradient row (red.global.add.v4.f32) and its touched stamp into the gradient arena of the rank
 *                    that updates the row, over NVLink for the other ranks' rows; its share of the projection's
 *                    gradient goes to the last rank the same way.  Two barrier kernels per step (gradients landed /
 *                    parameters landed); a new parameter row is sent only to the ranks whose instances of the next
 *                    batch read it.
 * The reference trains on one device (sert/models.py:520-560 builds one update function); there is no counterpart.
 * Collective: every rank calls it with its communicator; parameters and optimiser step are taken from rank 0.
 * comm = NULL detaches.  The optimiser state of a rank is current only inside its piece:
 * sert_model_gather_table_state makes it whole everywhere (checkpoints). */
SERT_API int sert_model_set_table_shard_comm(sert_model *m, sert_comm *comm, int32_t peer_stores);
SERT_API int sert_model_gather_table_state(sert_model *m);
/* mode: 0 none, 1 broadcast, 2 peer stores, 3 instance shards; [own_begin, own_end): this rank's float range of the tables' table_floats */
SERT_API int sert_model_table_shard_info(sert_model *m, int32_t *mode, int64_t *own_begin, int64_t *own_end,
                                         int64_t *table_floats);

/* Host only (no device needed): the cut sert_model_set_table_shard_comm makes for this configuration over `world`
 * ranks; every array has world + 1 entries.  Rank r updates entity rows [entity_row_bounds[r], [r + 1]), word rows
 * [word_row_bounds[r], [r + 1]) = floats [float_bounds[r], [r + 1]) of the parameter arrays (pieces end on rows whose
 * index is a multiple of 4), and with instance shards runs instances [instance_bounds[r], [r + 1]) of a batch. */
SERT_API int sert_table_shard_plan(const sert_config *cfg, int32_t world, int64_t *entity_row_bounds,
                                   int64_t *word_row_bounds, int64_t *float_bounds, int32_t *instance_bounds);

/* ---- device-resident data set: replaces the theano.shared X/Y/W variables, sert/models.py:470-480 -- */
/* x_dev (N,W) int32; labels either one-hot y_dev (N,) int32 (vector space, bin/train.py:186-245) or CSR
 * (indptr_dev int64 (N+1), indices_dev int32, data_dev f32; bin/prepare.py:593-597); w_dev (N,) f32 or NULL
 * (validation has no weights).  Pointers are borrowed until the model is destroyed. */
SERT_API int sert_model_attach_dataset(sert_model *m, int split, int64_t n, const int32_t *x_dev,
                              const int32_t *y_dev, const int64_t *indptr_dev, const int32_t *indices_dev,
                              const float *data_dev, const float *w_dev);

/* ---- train_fn(batch_index): sert/models.py:581-588, driven by _iterate_batches :351-399 ------- */
/* Runs n training batches in the given order (host int64 batch indices; the shuffled order of
 * sert/models.py:363-367).  neg_dev: (n,B,k) int32 device negatives or NULL = sample on device
 * (sert/models.py:947-979).  Per-batch train losses go to loss slots [first_slot, first_slot+n). */
SERT_API int sert_train_batches(sert_model *m, const int64_t *order_host, int64_t n, const int32_t *neg_dev,
                       int32_t first_slot);
/* test_fn / validate_fn (sert/models.py:593-608): eval loss of n batches of `split`. */
SERT_API int sert_eval_batches(sert_model *m, int split, const int64_t *order_host, int64_t n,
                      const int32_t *neg_dev, int32_t first_slot);
/* Synchronises, copies loss slots to the host and fails with the reference's message if any is NaN/Inf
 * (sert/models.py:372-379). */
SERT_API int sert_losses_fetch(sert_model *m, int32_t first_slot, int64_t n, float *out_host);

/* End-to-end variant of train_fn for callers that stream batches from the host: copies one batch
 * (x (B,W) int32, labels, w (B,), negatives (B,k) or NULL) host->device, runs the step, reads the loss back. */
SERT_API int sert_train_batch_host(sert_model *m, const int32_t *x_host, const int32_t *y_host,
                          const int64_t *indptr_host, const int32_t *indices_host, const float *data_host,
                          const float *w_host, const int32_t *neg_host, float *loss_host);

/* Pipelined form of the above for the vector-space model: enqueues host->device copies (own copy stream, two
 * staging sets) and the step, and returns a ticket without waiting for the device, so the host prepares batch n+1
 * while batch n runs.  x/y/w/neg must stay valid until the step has started (pinned memory for real overlap).
 * sert_train_host_wait blocks until the loss of `ticket` (one of the last 8 issued) is in host memory, stores it in
 * *loss_host and fails with the reference's NaN/Inf message (sert/models.py:372-379) like the synchronous call.
 * Same results as sert_train_batch_host: only the host/device synchronisation moves. */
SERT_API int sert_train_batch_host_async(sert_model *m, const int32_t *x_host, const int32_t *y_host,
                                const float *w_host, const int32_t *neg_host, int64_t *ticket_out);
SERT_API int sert_train_host_wait(sert_model *m, int64_t ticket, float *loss_host);

/* ---- run files: the step behind scoring (cvangysel-common trec_utils.write_run, trec_utils.py:531-580) ---- */
/* Formats n_lines lines "<subject> Q0 <object> <rank> <relevance> <model_name>\n" into `out` (host memory, `capacity`
 * bytes; 64 bytes + the three strings per line always suffice).  Subjects and objects are UTF-8 blobs with n+1 byte
 * offsets; every line names its subject, object, rank and relevance.  The relevance is printed like Python's
 * '{0}'.format(float) (repr: shortest round-trip digits, exponent notation outside 1e-4 <= |v| < 1e16), so the bytes
 * equal write_run's.  Pure host code (no device needed).  Returns the bytes written, -1 on a null argument, -2 when
 * `out` is too small. */
SERT_API int64_t sert_format_run(const char *subject_blob, const int64_t *subject_off, const char *object_blob,
                        const int64_t *object_off, const int32_t *line_subject, const int32_t *line_object,
                        const int32_t *line_rank, const double *line_relevance, int64_t n_lines,
                        const char *model_name, char *out, int64_t capacity);

/* ---- parity hooks (no reference counterpart; expose the graph's intermediate tensors) ---------- */
/* vector space: forward of one attached batch; out_scores_host (B,1+k) = [u.E[y], u.E[n_j]] logits,
 * out_proj_host (B,de) = clipped tanh projection u, out_ell_host (B,) instance losses. Any may be NULL. */
SERT_API int sert_vs_forward_host(sert_model *m, int split, int64_t batch_index, const int32_t *neg_dev,
                         float *out_scores_host, float *out_proj_host, float *out_ell_host);
/* log-linear: z (B*W,E) per-word logits, s (B,E) joint logits, ell (B,). Any may be NULL. */
SERT_API int sert_ll_forward_host(sert_model *m, int split, int64_t batch_index, float *out_z_host,
                         float *out_s_host, float *out_ell_host);

/* ---- predict_fn ------------------------------------------------------------------------------ */
/* log-linear predict_fn(batch, mask) -> (rows,W,E) per-word softmax (sert/models.py:880-890; the mask is
 * accepted and unused there).  rows <= B. Host in, host out. */
SERT_API int sert_predict_loglinear(sert_model *m, const int32_t *batch_host, int32_t rows, float *out_host);
/* LogLinearCallback.process on the device (bin/query.py:204-233; aggregate_distribution(mode='product'),
 * sert/inference.py:170-183).  `batch_host` is the WordBatcher's (rows, W) index array; query j occupies rows from
 * query_first_row[j] with its query_terms[j] tokens laid out row-major (sert/inference.py:99-111).  For every query:
 * per-term softmax over all E entities, sum over its terms of log p (exact zeros skipped, as np.ma.log(p).filled(0)),
 * exp, renormalise, and the full descending order.  Returns the first `top` (<= E; E = rank everything, as the
 * reference does) entity ids and relevances per query, the normalised entropy of every term's distribution in query
 * order (compute_normalised_entropy, bin/query.py:370-376; NULL to skip), the normalised entropy of each query's final
 * distribution and its mass after renormalisation (the reference's isclose(sum, 1) check).  Ties order by lower
 * entity id.  Host in / host out; nothing of size (rows, W, E) leaves the device. */
SERT_API int sert_ll_rank_queries(sert_model *m, const int32_t *batch_host, int32_t rows,
                                  const int32_t *query_first_row_host, const int32_t *query_terms_host,
                                  int32_t num_queries, int32_t top, int32_t *out_idx_host, float *out_rel_host,
                                  float *out_term_entropy_host, float *out_entropy_host, float *out_mass_host);
/* The same ranking from given per-term distributions (num_terms, entities) -- the callback's own input contract
 * (bin/query.py:204: `distribution`): query j owns rows query_first_term[j] .. +query_terms[j].  No model needed. */
SERT_API int sert_ll_rank_distributions(const float *dist_host, int32_t num_terms, int32_t entities,
                                        const int32_t *query_first_term_host, const int32_t *query_terms_host,
                                        int32_t num_queries, int32_t top, int32_t *out_idx_host, float *out_rel_host,
                                        float *out_term_entropy_host, float *out_entropy_host, float *out_mass_host);
/* vector-space predict_fn(avg) -> tanh(avg.W+b), no clip (sert/models.py:1107-1118), batched over q rows. */
SERT_API int sert_project_queries(sert_model *m, const float *avg_host, int32_t q, float *out_host);

/* ---- entity scoring: replaces VectorSpaceCallback's sklearn kNN / cdist, bin/query.py:241-318 -- */
/* Builds a scorer over `rows` entity vectors of dimension d (host float32, row-major).  normalise!=0
 * L2-normalises rows first (bin/query.py:270-274).  The rows are this rank's shard
 * [row_begin, row_begin+rows) of the global matrix.  arena as for models. */
SERT_API int sert_scorer_arena_bytes(int64_t rows, int32_t d, int32_t max_queries, int32_t max_k, size_t *bytes);
SERT_API int sert_scorer_create(const float *entities_host, int64_t rows, int32_t d, int64_t row_begin,
                       int32_t normalise, int32_t max_queries, int32_t max_k, void *arena_dev,
                       size_t arena_bytes, void *stream, sert_scorer **out);
SERT_API int sert_scorer_destroy(sert_scorer *s);
/* Scoring arithmetic.  1 (default) = tcgen05 tensor cores, coarse-then-exact: one bf16 GEMM scores every row, each
 * query keeps every row within a rigorous rounding-error margin of its k-th best (|q - bf16(q)| max|e| +
 * |bf16(q)| max|e - bf16(e)|, see score.cu), and the survivors are re-scored in fp32, so the returned top k is the
 * exact fp32 top k at a third of the tensor work.  The sweep is ONE GEMM launch over the shard: a strided row sample
 * seeds every query's threshold first, and one kernel selects, re-scores and sorts afterwards; a query whose list
 * comes up short or holds too many near-ties sends the call to the chunked sweeps (3, then 2) automatically.
 * 3 = the same coarse arithmetic in growing chunks with a prune after each (no threshold seed).  2 = tensor cores on
 * a 3-term bf16 split of both operands (fp32-class scores throughout) + fp32 re-scoring.  0 = fp32 FMA tiles on CUDA
 * cores. */
SERT_API int sert_scorer_set_mode(sert_scorer *s, int32_t mode);
/* Host counters since creation: top-k calls answered by the seeded one-launch sweep / calls that fell back to the
 * chunked sweeps (measurement and tests; no reference counterpart). */
SERT_API int sert_scorer_stats(sert_scorer *s, int64_t *seeded_sweeps, int64_t *fallback_sweeps);
/* The threshold-seeding plan of a top-k call on this shard (tests, diagnostics): the sample is `groups` groups of
 * `group_rows` consecutive rows, taken from every `tile_stride`-th tile of 256 rows; tau is seeded from the
 * `rank`-th largest group maximum; about `expected_survivors` rows per query pass.  group_rows = 0: no seed (the
 * shard fits a candidate list, or no sample small enough exists). */
SERT_API int sert_scorer_plan(sert_scorer *s, int32_t k, int32_t *group_rows, int32_t *groups, int32_t *rank,
                              int64_t *tile_stride, double *expected_survivors);
/* Row-sharded scoring: this scorer holds rows [row_begin, row_begin+rows) of the global matrix and
 * `comm` joins the other shards.  Every top-k call then runs the local sweep, ONE ncclAllGather of the packed
 * (row id, score)[q,k] lists (q*k*8 bytes per rank) and a k-way merge on the scorer's stream, and returns the top k of
 * the GLOBAL matrix on every rank -- the same list a single device returns (keys carry global row ids; ties order by
 * row id).  All ranks must issue the same calls.  comm = NULL detaches. */
SERT_API int sert_scorer_set_comm(sert_scorer *s, sert_comm *comm);
/* Top-k by inner product of q query vectors (host f32 (q,d); normalise_q!=0 L2-normalises them,
 * bin/query.py:333-336) against the shard.  Outputs (q,k) global row ids and float32 inner products,
 * sorted by score descending (ties: lower row id first). */
SERT_API int sert_scorer_topk_host(sert_scorer *s, const float *queries_host, int32_t q, int32_t normalise_q,
                          int32_t k, int32_t *out_idx_host, float *out_score_host);
/* Dense scores (q, rows) of every query against every row of the shard, host in / host out: the
 * "all entities" mode of VectorSpaceCallback.query (scipy cdist over all E, bin/query.py:310-318). */
SERT_API int sert_scorer_scores_host(sert_scorer *s, const float *queries_host, int32_t q, int32_t normalise_q,
                                     float *out_host);
/* Same, device in / device out, asynchronous on the scorer's stream (used under NCCL sharding). */
SERT_API int sert_scorer_topk_dev(sert_scorer *s, const float *queries_dev, int32_t q, int32_t normalise_q,
                         int32_t k, int32_t *out_idx_dev, float *out_score_dev);
/* k-way merge of `parts` gathered (q,k) candidate lists laid out [part][q][k] into one (q,k) list. */
SERT_API int sert_topk_merge_dev(const int32_t *idx_dev, const float *score_dev, int32_t parts, int32_t q, int32_t k,
                        int32_t *out_idx_dev, float *out_score_dev, void *stream);

/* ---- test hook -------------------------------------------------------------------------------- */
/* C (m,n) = A (m,k) . B (n,k)^T (+ bias (n,)) through the tcgen05/TMEM/TMA GEMM with `terms` bf16 split terms per
 * operand (1; 3 = [hi|hi|mid] x [hi|mid|hi]; 2 = pair operands [hi|mid], the log-linear path); host in / host out.  Used by tests/test_gpu_gemm_tc.py only. */
SERT_API int sert_debug_gemm_tc(const float *a_host, const float *b_host, int m, int n, int k, int terms,
                                const float *bias_host, float *c_host);
/* the same product with B given transposed, bt_host (k, n) row-major: the N-major B operand of the pair path */
SERT_API int sert_debug_gemm_tc_bn(const float *a_host, const float *bt_host, int m, int n, int k, float *c_host);
/* Raw throughput of the tcgen05 kernel on zero operands (mode 0: store epilogue into one aliased row, mode 1:
 * top-k filter that rejects everything); mean launch time in ms.  Used by tools/gemm_bench.py only. */
SERT_API int sert_debug_gemm_tc_bench(int m, int n, int kt, int reps, int mode, float *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* SERT_B200_H_ */
