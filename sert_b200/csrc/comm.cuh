// Communicator handle shared by score.cu (sharded scoring), sert_abi.cu (entity-sharded log-linear step) and
// the table-sharded vector-space step (sert_abi.cu: sert_model_set_table_shard_comm); see comm.cu.
#pragma once

#include "common.cuh"

struct sert_comm {
  int rank = 0, world = 1, device = 0;
  void *nccl = nullptr;          // ncclComm_t (null when world == 1)
  uint64_t collectives = 0;      // issued so far
  uint64_t bytes = 0;            // payload bytes this rank contributed / received (diagnostic)
};

namespace sert {

// Every call enqueues on `st`; nothing synchronises.
// recv holds world blocks of bytes_per_rank; in-place when send == recv + rank * bytes_per_rank.
int comm_all_gather(sert_comm *c, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t st);
int comm_all_reduce_sum_f32(sert_comm *c, float *buf, size_t count, cudaStream_t st);
int comm_all_reduce_sum_f64(sert_comm *c, double *buf, size_t count, cudaStream_t st);
// rank r's bytes [off[r], off[r] + len[r]) of `base` reach every rank, in place (grouped broadcasts)
int comm_gather_pieces(sert_comm *c, void *base, const size_t *off, const size_t *len, cudaStream_t st);
int comm_broadcast(sert_comm *c, void *buf, size_t bytes, int root, cudaStream_t st);
// CUDA IPC: peers[r] = local address of rank r's cudaMalloc'ed buffer (peers[rank] = local); collective, synchronises
int comm_map_peers(sert_comm *c, void *local, void **peers, cudaStream_t st);
int comm_unmap_peers(sert_comm *c, void **peers);
int comm_reduce_scatter_sum_f32(sert_comm *c, const float *send, float *recv, size_t count_per_rank, cudaStream_t st);
int comm_all_to_all_v(sert_comm *c, const void *send, const size_t *send_off, const size_t *send_bytes, void *recv,
                      const size_t *recv_off, const size_t *recv_bytes, cudaStream_t st);

}  // namespace sert
