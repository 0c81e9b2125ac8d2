// Barrier + small all-reduce between the ranks of a table-shard group, over NVLink peer memory (no NCCL on the step).
//
// Every rank owns one PeerSyncBlock (cudaMalloc'ed by the library, mapped by every other rank with CUDA IPC:
// comm_map_peers).  One launch of peer_barrier_kernel on rank r at epoch e
//   1. stores r's partial accumulators (data-loss sum, 64 partial sums of theta^2) into inbox[e & 1][r] of EVERY
//      rank's block,
//   2. fences at system scope and release-stores e into flag[r] of every rank's block,
//   3. waits until all `world` flags of its OWN block have reached e (acquire loads), and
//   4. replaces the local accumulators by the sum of the world inbox rows, added in rank order on every rank, so that
//      all ranks hold bit-identical sums (one model: every rank reports the same loss).
// The launch sits on the model's stream behind the step's update kernels, whose NVLink stores into the other ranks'
// parameter buffers have therefore completed before step 2 signals; a rank leaves the kernel only after every rank
// has signalled, i.e. after every store into ITS buffers has landed.  Inbox rows are double-banked by epoch parity:
// a rank can be at most one barrier ahead of the slowest (it cannot pass barrier e+1 before the slowest has signalled
// e+1, which that one does after it has read its epoch-e inbox).
// A rank that waits longer than kTimeoutNs (a peer died) raises *error (page-locked host memory) and leaves; the host
// turns it into an error at the next fetch.
#include "kernels.cuh"

namespace sert {

namespace {

constexpr unsigned long long kTimeoutNs = 20ull * 1000 * 1000 * 1000;

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double *p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(512) peer_barrier_kernel(PeerBarrierArgs a) {
  const int tid = threadIdx.x;
  const int bank = (int)(a.epoch & 1u);
  if (a.acc != nullptr) {
    for (int t = tid; t < a.world * kAccSlots; t += blockDim.x) {
      const int p = t / kAccSlots, s = t % kAccSlots;
      if (s >= a.acc_first) a.blk[p]->inbox[bank][a.rank][s] = a.acc[s];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (tid < a.world) {
    st_release_sys(&a.blk[tid]->flag[a.rank][0], a.epoch);
    const uint32_t *mine = &a.blk[a.rank]->flag[tid][0];
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - a.epoch) < 0) {
      if (global_timer_ns() - t0 > kTimeoutNs) {
        *a.error = 1u;
        break;
      }
    }
  }
  __syncthreads();
  if (a.acc != nullptr && tid >= a.acc_first && tid < kAccSlots) {
    double v = 0.0;
    for (int r = 0; r < a.world; ++r) v += ld_volatile_f64(&a.blk[a.rank]->inbox[bank][r][tid]);
    a.acc[tid] = v;
  }
}

}  // namespace

int launch_peer_barrier(const PeerBarrierArgs &a, cudaStream_t st) {
  SERT_REQUIRE(a.world >= 1 && a.world <= kMaxPeers + 1 && a.rank >= 0 && a.rank < a.world, "bad rank / world");
  SERT_REQUIRE(a.error != nullptr, "null error word");
  for (int r = 0; r < a.world; ++r) SERT_REQUIRE(a.blk[r] != nullptr, "unmapped peer block");
  peer_barrier_kernel<<<1, 512, 0, st>>>(a);
  SERT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sert
