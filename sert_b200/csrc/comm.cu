// NCCL communicator of libsert_b200: the collectives of the two sharded paths run INSIDE the library, on the
// stream of the scorer / model that issues them (sert_comm_init(rank, world, ncclUniqueId*)):
//   * row-sharded entity scoring: ONE ncclAllGather of the per-shard (row id, score)[Q,k] lists (score.cu),
//   * entity-sharded log-linear training: the five small exchanges of a step (sert_abi.cu: xchg).
// The reference is single-device and has no counterpart.  The host only ships the 128-byte unique id between the
// ranks (torch.distributed's store in sert_b200/comm.py; any out-of-band channel works).
//
// libnccl.so.2 is resolved at run time with dlopen: a process that already imported torch gets the NCCL torch
// loaded (same soname), so exactly one NCCL lives in the process; a process that never shards needs no NCCL at all.
#include "comm.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

namespace sert {

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                                cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi g_api;
std::once_flag g_once;
std::string g_load_error;

template <typename F>
bool bind(F &fn, const char *name) {
  fn = reinterpret_cast<F>(dlsym(g_api.handle, name));
  if (fn == nullptr) g_load_error = std::string("libnccl.so.2 lacks ") + name;
  return fn != nullptr;
}

void load_nccl() {
  g_api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (g_api.handle == nullptr) {
    g_load_error = std::string("cannot load libnccl.so.2: ") + dlerror();
    return;
  }
  bool ok = bind(g_api.GetUniqueId, "ncclGetUniqueId") && bind(g_api.CommInitRank, "ncclCommInitRank") &&
            bind(g_api.CommDestroy, "ncclCommDestroy") && bind(g_api.AllGather, "ncclAllGather") &&
            bind(g_api.AllReduce, "ncclAllReduce") && bind(g_api.ReduceScatter, "ncclReduceScatter") &&
            bind(g_api.Broadcast, "ncclBroadcast") && bind(g_api.GroupStart, "ncclGroupStart") && bind(g_api.GroupEnd, "ncclGroupEnd") &&
            bind(g_api.Send, "ncclSend") && bind(g_api.Recv, "ncclRecv") &&
            bind(g_api.GetErrorString, "ncclGetErrorString") && bind(g_api.GetVersion, "ncclGetVersion");
  if (!ok) g_api.handle = nullptr;
}

int api() {
  std::call_once(g_once, load_nccl);
  if (g_api.handle == nullptr) {
    set_error(g_load_error);
    return -1;
  }
  return 0;
}

#define SERT_NCCL(expr)                                                                        \
  do {                                                                                         \
    ncclResult_t _r = (expr);                                                                  \
    if (_r != ncclSuccess) {                                                                   \
      ::sert::set_error(std::string(#expr) + ": " + g_api.GetErrorString(_r));                 \
      return -1;                                                                               \
    }                                                                                          \
  } while (0)

}  // namespace

int comm_all_gather(sert_comm *c, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t st) {
  SERT_REQUIRE(c != nullptr, "null communicator");
  if (c->world == 1) {
    if (send != recv) SERT_CUDA(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  SERT_NCCL(g_api.AllGather(send, recv, bytes_per_rank, ncclInt8, static_cast<ncclComm_t>(c->nccl), st));
  ++c->collectives;
  c->bytes += bytes_per_rank * (size_t)c->world;
  return 0;
}

int comm_all_reduce_sum_f32(sert_comm *c, float *buf, size_t count, cudaStream_t st) {
  SERT_REQUIRE(c != nullptr, "null communicator");
  if (c->world == 1) return 0;
  SERT_NCCL(g_api.AllReduce(buf, buf, count, ncclFloat32, ncclSum, static_cast<ncclComm_t>(c->nccl), st));
  ++c->collectives;
  c->bytes += count * 4;
  return 0;
}

int comm_all_reduce_sum_f64(sert_comm *c, double *buf, size_t count, cudaStream_t st) {
  SERT_REQUIRE(c != nullptr, "null communicator");
  if (c->world == 1) return 0;
  SERT_NCCL(g_api.AllReduce(buf, buf, count, ncclFloat64, ncclSum, static_cast<ncclComm_t>(c->nccl), st));
  ++c->collectives;
  c->bytes += count * 8;
  return 0;
}

// In-place all-gather of unequal pieces: rank r's bytes [off[r], off[r] + len[r]) of `base` reach every rank (one
// grouped ncclBroadcast per piece: NCCL fuses the group into one launch).
int comm_gather_pieces(sert_comm *c, void *base, const size_t *off, const size_t *len, cudaStream_t st) {
  SERT_REQUIRE(c != nullptr, "null communicator");
  if (c->world == 1) return 0;
  ncclComm_t comm = static_cast<ncclComm_t>(c->nccl);
  SERT_NCCL(g_api.GroupStart());
  for (int r = 0; r < c->world; ++r) {
    if (len[r] == 0) continue;
    char *piece = static_cast<char *>(base) + off[r];
    SERT_NCCL(g_api.Broadcast(piece, piece, len[r], ncclInt8, r, comm, st));
    c->bytes += len[r];
  }
  SERT_NCCL(g_api.GroupEnd());
  ++c->collectives;
  return 0;
}

int comm_broadcast(sert_comm *c, void *buf, size_t bytes, int root, cudaStream_t st) {
  SERT_REQUIRE(c != nullptr && root >= 0 && root < c->world, "bad broadcast");
  if (c->world == 1 || bytes == 0) return 0;
  SERT_NCCL(g_api.Broadcast(buf, buf, bytes, ncclInt8, root, static_cast<ncclComm_t>(c->nccl), st));
  ++c->collectives;
  c->bytes += bytes;
  return 0;
}

// Peer mappings of a cudaMalloc'ed buffer of every rank (CUDA IPC; the ranks are processes of one node whose GPUs
// reach each other over NVLink): peers[r] = this process's address of rank r's buffer, peers[rank] = local.  The
// 64-byte handles travel through one ncclAllGather.  Collective: every rank of the communicator calls it.
int comm_map_peers(sert_comm *c, void *local, void **peers, cudaStream_t st) {
  SERT_REQUIRE(c != nullptr && local != nullptr && peers != nullptr, "null argument");
  for (int r = 0; r < c->world; ++r) peers[r] = nullptr;
  peers[c->rank] = local;
  if (c->world == 1) return 0;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
  cudaIpcMemHandle_t mine;
  SERT_CUDA(cudaIpcGetMemHandle(&mine, local));
  char *dev = nullptr;
  SERT_CUDA(cudaMalloc(&dev, (size_t)c->world * sizeof(mine)));
  std::vector<cudaIpcMemHandle_t> all((size_t)c->world);
  int rc = 0;
  do {
    if (cudaMemcpyAsync(dev + (size_t)c->rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        comm_all_gather(c, dev + (size_t)c->rank * sizeof(mine), dev, sizeof(mine), st) != 0 ||
        cudaMemcpyAsync(all.data(), dev, all.size() * sizeof(mine), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
      rc = -1;
      break;
    }
    for (int r = 0; r < c->world && rc == 0; ++r) {
      if (r == c->rank) continue;
      const cudaError_t e = cudaIpcOpenMemHandle(&peers[r], all[(size_t)r], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        set_error(std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
        (void)cudaGetLastError();
        rc = -1;
      }
    }
  } while (false);
  cudaFree(dev);
  if (rc != 0) {
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && peers[r] != nullptr) { cudaIpcCloseMemHandle(peers[r]); peers[r] = nullptr; }
  }
  return rc;
}

int comm_unmap_peers(sert_comm *c, void **peers) {
  if (c == nullptr || peers == nullptr) return 0;
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && peers[r] != nullptr) { cudaIpcCloseMemHandle(peers[r]); peers[r] = nullptr; }
  return 0;
}

int comm_reduce_scatter_sum_f32(sert_comm *c, const float *send, float *recv, size_t count_per_rank, cudaStream_t st) {
  SERT_REQUIRE(c != nullptr, "null communicator");
  if (c->world == 1) {
    if (send != recv) SERT_CUDA(cudaMemcpyAsync(recv, send, count_per_rank * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  SERT_NCCL(g_api.ReduceScatter(send, recv, count_per_rank, ncclFloat32, ncclSum, static_cast<ncclComm_t>(c->nccl), st));
  ++c->collectives;
  c->bytes += count_per_rank * 4 * (size_t)c->world;
  return 0;
}

// Variable-size all-to-all of bytes: rank r sends send[send_off[p] .. +send_bytes[p]) to rank p and receives
// recv_bytes[p] bytes from p at recv_off[p] (one grouped ncclSend/ncclRecv per peer).
int comm_all_to_all_v(sert_comm *c, const void *send, const size_t *send_off, const size_t *send_bytes, void *recv,
                      const size_t *recv_off, const size_t *recv_bytes, cudaStream_t st) {
  SERT_REQUIRE(c != nullptr, "null communicator");
  if (c->world == 1) {
    if (send_bytes[0])
      SERT_CUDA(cudaMemcpyAsync(static_cast<char *>(recv) + recv_off[0], static_cast<const char *>(send) + send_off[0],
                                send_bytes[0], cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  ncclComm_t comm = static_cast<ncclComm_t>(c->nccl);
  SERT_NCCL(g_api.GroupStart());
  for (int p = 0; p < c->world; ++p) {
    if (send_bytes[p])
      SERT_NCCL(g_api.Send(static_cast<const char *>(send) + send_off[p], send_bytes[p], ncclInt8, p, comm, st));
    if (recv_bytes[p])
      SERT_NCCL(g_api.Recv(static_cast<char *>(recv) + recv_off[p], recv_bytes[p], ncclInt8, p, comm, st));
    c->bytes += send_bytes[p];
  }
  SERT_NCCL(g_api.GroupEnd());
  ++c->collectives;
  return 0;
}

}  // namespace sert

using namespace sert;

extern "C" {

int sert_comm_unique_id(void *id_out, size_t capacity) {
  SERT_REQUIRE(id_out != nullptr && capacity >= sizeof(ncclUniqueId), "the unique id needs 128 bytes");
  if (api()) return -1;
  ncclUniqueId id;
  SERT_NCCL(g_api.GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

int sert_comm_init(int32_t rank, int32_t world, const void *unique_id, sert_comm **out) {
  SERT_REQUIRE(out != nullptr && world >= 1 && rank >= 0 && rank < world, "bad rank / world");
  sert_comm *c = new sert_comm();
  c->rank = rank;
  c->world = world;
  SERT_CUDA(cudaGetDevice(&c->device));
  if (world > 1) {
    if (unique_id == nullptr) { delete c; set_error("null unique id"); return -1; }
    if (api()) { delete c; return -1; }
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_t comm = nullptr;
    ncclResult_t r = g_api.CommInitRank(&comm, world, id, rank);
    if (r != ncclSuccess) {
      delete c;
      set_error(std::string("ncclCommInitRank: ") + g_api.GetErrorString(r));
      return -1;
    }
    c->nccl = comm;
  }
  *out = c;
  return 0;
}

int sert_comm_destroy(sert_comm *c) {
  if (c == nullptr) return 0;
  if (c->nccl != nullptr && g_api.CommDestroy != nullptr) g_api.CommDestroy(static_cast<ncclComm_t>(c->nccl));
  delete c;
  return 0;
}

int sert_comm_info(sert_comm *c, int32_t *rank, int32_t *world, int32_t *nccl_version, int64_t *collectives,
                   int64_t *bytes) {
  SERT_REQUIRE(c != nullptr, "null communicator");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  if (nccl_version) {
    int v = 0;
    if (c->world > 1 && g_api.GetVersion) g_api.GetVersion(&v);
    *nccl_version = v;
  }
  if (collectives) *collectives = (int64_t)c->collectives;
  if (bytes) *bytes = (int64_t)c->bytes;
  return 0;
}

}  // extern "C"
