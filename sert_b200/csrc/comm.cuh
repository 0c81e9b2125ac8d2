// Communicator handle shared by score.cu (sharded scoring), sert_abi.cu (entity-sharded log-linear step) and
// vs_shard.cu (row-sharded vector-space step); see comm.cu.
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
int comm_reduce_scatter_sum_f32(sert_comm *c, const float *send, float *recv, size_t count_per_rank, cudaStream_t st);
int comm_all_to_all_v(sert_comm *c, const void *send, const size_t *send_off, const size_t *send_bytes, void *recv,
                      const size_t *recv_off, const size_t *recv_bytes, cudaStream_t st);

}  // namespace sert
