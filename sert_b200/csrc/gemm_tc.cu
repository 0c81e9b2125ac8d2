// tcgen05 (5th-gen tensor core) GEMM with TMEM accumulators and TMA operand staging; see gemm_tc.cuh.
//
// Kernel anatomy (one persistent CTA per SM, 192 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor 2D loads of the A (128 x 64) and B (256 x 64) bf16 tiles
//            into a 4-stage 128B-swizzled shared-memory ring, completion via mbarrier expect_tx
//   warp 1   allocates all 512 TMEM columns (two 128 x 256 fp32 accumulators), then one elected lane
//            issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16) x 4 per stage and
//            tcgen05.commit's the stage's "empty" barrier / the accumulator's "full" barrier
//   warps 2-9 epilogue (two per 32-lane TMEM quarter, 128 columns each): tcgen05.ld 32x32b.x32, then fp32 stores (+bias)
//            or the running top-k filter; double-buffered against the next tile's MMAs
#include <cuda.h>
#include <stdlib.h>

#include "gemm_tc.cuh"
#include "topk_keys.cuh"

namespace sert {

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr uint32_t B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
#ifndef SERT_TC_EPI_WARPS
#define SERT_TC_EPI_WARPS 16
#endif
constexpr int EPI_WARPS = SERT_TC_EPI_WARPS;   // EPI_WARPS/4 warps per TMEM lane quarter, each owns a slice of the tile's columns
constexpr int COLS_PER_WARP = BN / (EPI_WARPS / 4);
static_assert(EPI_WARPS % 4 == 0 && COLS_PER_WARP % 32 == 0, "epilogue warps split the columns in 32-column chunks");
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t TMEM_COLS = 512;
// top-k epilogue: survivors of one tile that a WARP parks in shared memory (keys compacted across its 32 rows with
// ballots; 8-byte key + 2-byte (owner lane, ordinal within the owner's row)), two buffers: this tile's and the
// previous one's, whose list slots are being reserved
#ifndef SERT_TC_TSTORE
#define SERT_TC_TSTORE 1          // store epilogue: transposed 128-byte row stores (0: one row per lane)
#endif
#ifndef SERT_TC_GROUP
#define SERT_TC_GROUP 16
#endif
constexpr int GW = SERT_TC_GROUP;         // columns of the accumulator the top-k epilogue examines at a time (8 or 16)
static_assert(GW == 8 || GW == 16, "tcgen05.ld x8 / x16");
constexpr int WSTASH = 96;
constexpr uint32_t STASH_KEYS_BYTES = 2 * EPI_WARPS * WSTASH * 8;
constexpr uint32_t STASH_BYTES = STASH_KEYS_BYTES + 2 * EPI_WARPS * WSTASH * 2;
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES + 256 /*barriers*/ + STASH_BYTES;

struct alignas(64) TcMap {
  unsigned char bytes[128];
};

struct KernelArgs {
  int M;
  long long n_begin, n_end;
  int num_kb;          // Kt / 64
  int k_slices;        // split-K: the K blocks are cut into this many slices, each a tile of its own (TC_EPI_STORE
                       // with epi.accumulate: partial sums meet in C through float4 reductions)
  int b_nmajor;        // 1 (pair operands, single-CTA tiles): B lies N-major in global memory -- rows = K, columns = N, the
                       // hi block behind map_b and the mid block behind map_b2 (boxes of 64 N x 64 K); the MMA reads it
                       // through an MN-major descriptor.  Lets gWd = X^T . dZ of the log-linear model consume the same
                       // split dZ rows as dX = dZ . Wd^T, so no transposed copy of dZ is ever written.
  int tile_n;          // columns of C per n-tile: BN, or less (a multiple of 16, store epilogue) to cut a narrow C into
                       // n-tiles of EQUAL width -- dX of the log-linear model is 300 columns: 256 + 44 left the CTAs of the
                       // narrow tiles four times faster than their neighbours, and the A blocks both need (8 GB of split
                       // dZ) came from HBM twice; 160 + 140 keeps the pairs in step, so the second read hits L2
  int n_fast;          // 1: tiles are numbered n-fastest inside a K slice (cta_tile)
  int pair_kp;         // > 0: "pair" operands [hi | mid] (two blocks of pair_kp columns, launch_gemm_tc_pair); num_kb
                       // then counts 64-column blocks of ONE term, and each takes two ring stages: (A_hi, B_hi) and
                       // (A_mid, B_mid), from which the MMA warp forms hi.hi + hi.mid + mid.hi -- the three products of
                       // the bf16x3 split from 96 KB of operands instead of the 144 KB of the [hi|hi|mid] x [hi|mid|hi]
                       // layout
  TcEpilogue epi;
};

// Tile order of one CTA.  Classic: tiles blockIdx.x, +grid, ... of the m-fastest numbering.  B-stationary (BSTAT): the
// CTA owns whole n-tiles (blockIdx.x, +grid, ...) and sweeps every m-tile of each, so that the B operand of an n-tile
// is loaded ONCE and stays in shared memory: at K = 256 a 128 x 256 tile needs 192 KB of operands for 2048 cycles of
// MMA, more than L2 can feed all SMs (measured 8.8 TB/s of L2 reads, 6200 cycles per tile); keeping B resident
// leaves the 64 KB A tile.  Needs num_kb <= STAGES (B of one n-tile fits in the B slots of the ring).
// The CTAs start their sweeps at different m-tiles (blockIdx.x apart): 148 SMs asking L2 for the same A tile at the
// same moment queue up on the slices that hold its lines (measured: 10.3 k cycles per tile unstaggered).
// seq = position inside the n-tile's sweep (0 = first, num_m_tiles - 1 = last).
// bid / gdim: the CTA's place among the tile walkers -- blockIdx.x / gridDim.x, or the CLUSTER's index / count when two
// CTAs walk the tiles as a pair (num_m_tiles then counts pairs of m-tiles).
template <bool BSTAT>
__device__ __forceinline__ bool cta_tile(int it, int bid, int gdim, int num_m_tiles, int num_n_tiles, int k_slices,
                                         int step_m, int step_n, int &mt, int &nt, int &seq, int &slice,
                                         bool n_fast = false) {
  if (!BSTAT && n_fast) {
    // n-fastest numbering inside a K slice (store epilogue with a handful of n-tiles and a huge A operand -- dX of the
    // log-linear model): the n-tiles that share an A tile are neighbours in the tile order, so the CTAs working on
    // them run side by side and the second read of the A blocks hits L2.  step_m / step_n are then the steps of the
    // (n, slice * m-tiles + m) numbering.
    int fast, slow;
    if (it == 0) {
      fast = bid % num_n_tiles;
      slow = bid / num_n_tiles;
    } else {
      fast = nt - slice * num_n_tiles + step_m;
      slow = slice * num_m_tiles + mt + step_n;
      if (fast >= num_n_tiles) { fast -= num_n_tiles; ++slow; }
    }
    seq = 0;
    slice = slow / num_m_tiles;
    mt = slow - slice * num_m_tiles;
    nt = slice * num_n_tiles + fast;
    return slice < k_slices;
  }
  slice = 0;
  if (BSTAT) {
    if (it == 0) {
      nt = bid; seq = 0;
      mt = bid % num_m_tiles;
    } else {
      if (++seq == num_m_tiles) { seq = 0; nt += gdim; }
      if (++mt == num_m_tiles) mt = 0;
    }
    return nt < num_n_tiles;
  }
  seq = 0;
  // Incremental: every warp of the CTA walks its tiles with this function, and two 64-bit divisions per tile
  // (~150 instructions) were a fifth of an epilogue warp's work on a tile without survivors.  `mt` and `nt` carry
  // the previous tile's coordinates (nt counts n-tiles across all K slices); a step adds gridDim.x tiles.
  if (it == 0) {
    mt = bid % num_m_tiles;
    nt = bid / num_m_tiles;
  } else {
    mt += step_m;
    nt += step_n;
    if (mt >= num_m_tiles) { mt -= num_m_tiles; ++nt; }
  }
  if (k_slices == 1) return nt < num_n_tiles;
  slice = nt / num_n_tiles;
  return slice < k_slices;
}
// the n-tile inside its K slice (cta_tile leaves the running count over all slices in `nt`)
__device__ __forceinline__ int slice_nt(int nt, int num_n_tiles, int slice) { return nt - slice * num_n_tiles; }
// K blocks [kb0, kb0 + n) of slice `slice` (the last slice takes the remainder)
__device__ __forceinline__ void slice_range(int num_kb, int k_slices, int slice, int &kb0, int &n) {
  const int per = (num_kb + k_slices - 1) / k_slices;
  kb0 = slice * per;
  n = min(per, num_kb - kb0);
}

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const void *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// the same load, delivered to the same shared-memory offset (and mbarrier offset) of every CTA in `cta_mask`
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t dst, const void *map, uint32_t bar, int c0, int c1,
                                                      uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_alloc(uint32_t smem_result, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- two-SM MMA (cta_group::2): the CTA pair of a cluster multiplies ONE 256 x N tile; CTA r holds rows [128 r, +128)
// of A and rows [N/2 r, +N/2) of B in its own shared memory and the matching 128 rows of the accumulator in its own TMEM;
// the leader (rank 0) issues the MMAs, both CTAs' TMA loads complete on the LEADER's barrier.
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const void *map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void tc_alloc_cg2(uint32_t smem_result, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc_cg2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_cg2(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ... on the mbarrier at this offset of every CTA in `cta_mask` (the peer's producer waits for both MMA warps)
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tc_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <int N>
__device__ __forceinline__ void tc_ld_group(uint32_t taddr, uint32_t (&r)[N]) {
  if constexpr (N == 8) tc_ld_32x8(taddr, r);
  else tc_ld_32x16(taddr, r);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major bf16 tile stored as rows of 64 elements (128 bytes) with the
// 128-byte swizzle TMA wrote: start address >> 4 [0,14), LBO = 1 (unused for swizzled K-major) [16,30),
// SBO = 1024 bytes (8 rows x 128 B) >> 4 [32,46), descriptor version 1 [46,48), layout SWIZZLE_128B = 2 [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// The same for an MN-major operand (tile stored [K rows][64 MN elements] in 128-byte swizzled rows, 64 K rows = 8 KB per
// group of 64 MN elements, as four TMA boxes of 64 x 64 leave it): canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in
// 16-byte units -- LBO = distance between groups of 64 MN elements (8192 B), SBO = distance between groups of 8 K rows
// (1024 B).  One MMA step (K = 16) advances the start address by 16 rows = 2048 B.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t IDESC_B_MN_MAJOR = 1u << 16;   // instruction descriptor bit 16: B is MN-major
// Instruction descriptor (kind::f16): D fp32 (bits 4-5 = 1), A/B bf16 (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29).
constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- the kernel -----------------------------------------------------------------------------------
// MODE = the epilogue (TcEpilogueMode): one kernel per epilogue keeps each kernel's code small enough for the
// instruction caches (the epilogues are long unrolled register code; a warp only ever runs one of them).
// CL = 2 or 4: the CTAs run in clusters that share an n-tile and take CL adjacent m-tiles.  Each CTA loads its own A
// tile and 1/CL of the B tile, multicast into every CTA's shared memory (TMA .multicast::cluster), so a cluster pulls
// CL*16 + 32 KB per K block through L2 instead of CL*48 KB (-33 % at CL = 2, -50 % at CL = 4); a stage's slot is free
// for the next load once ALL the cluster's MMA warps have retired their reads of it (tcgen05.commit multicast on the
// empty barriers).
// CG2 (with CL = 2, store epilogue): the pair runs as ONE two-SM MMA (cta_group::2, M = 256) instead of two single-SM MMAs
// over a multicast B tile.  Each CTA then stages only its own 128 rows of A and its own half of B -- 32 instead of 48 KB
// per ring step, so the 192 KB ring holds SIX steps (three blocks of pair operands in flight instead of two), and the
// shared-memory write traffic of the B tile halves.
template <bool BSTAT, int MODE, int CL = 1, bool CG2 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ TcMap map_a, const __grid_constant__ TcMap map_b, const __grid_constant__ TcMap map_b2,
               const KernelArgs args) {
  static_assert(!CG2 || (CL == 2 && !BSTAT && MODE == TC_EPI_STORE), "the two-SM MMA serves the clustered store kernels");
  constexpr int STAGES = CG2 ? 6 : sert::STAGES;
  constexpr uint32_t B_STAGE_BYTES = CG2 ? sert::B_STAGE_BYTES / 2 : sert::B_STAGE_BYTES;
  constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static_assert(STAGES * STAGE_BYTES == sert::STAGES * sert::STAGE_BYTES, "the ring keeps its 192 KB");
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024-byte alignment
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
  // barrier layout (8 bytes each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_ptr_smem = bars + 8u * (2 * STAGES + 4);
  const uint32_t bfull_bar = bars + 8u * (2 * STAGES + 5);    // BSTAT: the n-tile's B blocks have landed
  const uint32_t bempty_bar = bars + 8u * (2 * STAGES + 6);   // BSTAT: every MMA of the n-tile has retired
  const uint32_t stash_base = bars + 256u;                    // top-k epilogue: STASH keys per epilogue thread

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m_tiles = (args.M + BM - 1) / BM;
  const int TN = MODE == TC_EPI_STORE ? args.tile_n : BN;               // columns of C per n-tile
  const int num_n_tiles = (int)((args.n_end - args.n_begin + TN - 1) / TN);
  const uint32_t cta_rank = CL > 1 ? cluster_cta_rank() : 0u;
  const int bid = (int)blockIdx.x / CL, gdim = (int)gridDim.x / CL;
  const int walk_m_tiles = (num_m_tiles + CL - 1) / CL;                  // m-tiles, or pairs of m-tiles
  const bool n_fast = !BSTAT && args.n_fast != 0;
  const int walk_fast = n_fast ? num_n_tiles : walk_m_tiles;
  const int step_m = gdim % walk_fast, step_n = gdim / walk_fast;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_b2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CG2 ? 1 : CL);  // one arrival per MMA warp that reads the slot (CG2: the leader's commit)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), CG2 ? 2 * EPI_WARPS : EPI_WARPS);   // one arrive per epilogue warp (CG2: of both CTAs, at the leader)
    }
    mbar_init(bfull_bar, 1);
    mbar_init(bempty_bar, 1);
    fence_barrier_init();
  }
  if (CG2) {
    __syncthreads();
    cluster_sync_all();                      // both CTAs are resident before the pair allocates its tensor memory
    if (warp == 1) tc_alloc_cg2(tmem_ptr_smem, TMEM_COLS);
  } else if (warp == 1) {
    tc_alloc(tmem_ptr_smem, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();            // the peers' barriers exist before anything is multicast at them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, bphase = 0;
      int mt, nt, seq, slice;
      for (int it = 0; cta_tile<BSTAT>(it, bid, gdim, walk_m_tiles, num_n_tiles, args.k_slices, step_m, step_n, mt, nt, seq, slice, n_fast); ++it) {
        const int m0 = (mt * CL + (int)cta_rank) * BM;
        const int n0 = (int)(args.n_begin + (long long)slice_nt(nt, num_n_tiles, slice) * TN * args.epi.tile_stride);
        int kb0, nkb;
        slice_range(args.num_kb, args.k_slices, slice, kb0, nkb);
        if (BSTAT && seq == 0) {
          mbar_wait(bempty_bar, bphase ^ 1u);            // the previous n-tile's MMAs no longer read the B slots
          mbar_expect_tx(bfull_bar, (uint32_t)args.num_kb * B_STAGE_BYTES);
          for (int kb = 0; kb < args.num_kb; ++kb)
            tma_load_2d(smem_b + kb * B_STAGE_BYTES, &map_b, bfull_bar, kb * BK, n0);
          bphase ^= 1u;
        }
        // ring steps: one per K block, or two (the hi and the mid columns of the block) for pair operands
        const int per_kb = args.pair_kp > 0 ? 2 : 1;
        for (int step = kb0 * per_kb; step < (kb0 + nkb) * per_kb; ++step) {
          const int kc = args.pair_kp > 0 ? (step >> 1) * BK + (step & 1) * args.pair_kp : step * BK;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (CG2) {
            // both CTAs' bytes complete on the leader's barrier; each CTA stages its own A rows and its half of the
            // B rows the MMA multiplies (N/2 rows from n0 + rank * N/2; the box is 128 rows, the rest is ignored)
            const uint32_t lead_full = mapa_cluster(full_bar(stage), 0u);
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2u * STAGE_BYTES);
            long long n_left = args.n_end - n0;
            if (n_left > TN) n_left = TN;
            const int n_mma = n_left < BN ? (int)((n_left + 15) / 16 * 16) : BN;
            tma_load_2d_cg2(smem_a + stage * A_STAGE_BYTES, &map_a, lead_full, kc, m0);
            tma_load_2d_cg2(smem_b + stage * B_STAGE_BYTES, &map_b, lead_full, kc, n0 + (int)cta_rank * (n_mma / 2));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            continue;
          }
          mbar_expect_tx(full_bar(stage), BSTAT ? A_STAGE_BYTES : STAGE_BYTES);
          tma_load_2d(smem_a + stage * A_STAGE_BYTES, &map_a, full_bar(stage), kc, m0);
          if (CL > 1) {
            // this CTA's share of the B tile (map_b's box is BN / CL rows), into every CTA of the cluster
            tma_load_2d_multicast(smem_b + stage * B_STAGE_BYTES + cta_rank * (B_STAGE_BYTES / CL), &map_b, full_bar(stage),
                                  kc, n0 + (int)cta_rank * (BN / CL), (uint16_t)((1u << CL) - 1u));
          } else if (!BSTAT && args.b_nmajor) {
            // N-major B: K rows [block * 64, +64) of the hi (even step) or mid (odd step) block, four boxes of 64 columns
            const TcMap *mb = (step & 1) ? &map_b2 : &map_b;
#pragma unroll
            for (int g = 0; g < BN / 64; ++g)
              tma_load_2d(smem_b + stage * B_STAGE_BYTES + g * (64 * BK * 2), mb, full_bar(stage), n0 + g * 64, (step >> 1) * BK);
          } else if (!BSTAT) {
            tma_load_2d(smem_b + stage * B_STAGE_BYTES, &map_b, full_bar(stage), kc, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0 && (!CG2 || cta_rank == 0)) {
      constexpr uint32_t idesc_full = make_idesc(BM, BN);
      auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t accum) {
        if (CG2) tc_mma_bf16_cg2(d, ad, bd, id, accum);
        else tc_mma_bf16(d, ad, bd, id, accum);
      };
      // the ring slot is free once the MMAs reading it have retired: in every CTA whose loads land there
      auto release_stage = [&](int st) {
        if (CG2) tc_commit_cg2(empty_bar(st), (uint16_t)0x3);
        else if (CL > 1) tc_commit_multicast(empty_bar(st), (uint16_t)((1u << CL) - 1u));
        else tc_commit(empty_bar(st));
      };
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0, bphase = 0;
      int mt, nt, seq, slice;
      for (int it = 0; cta_tile<BSTAT>(it, bid, gdim, walk_m_tiles, num_n_tiles, args.k_slices, step_m, step_n, mt, nt, seq, slice, n_fast); ++it) {
        int kb0, nkb;
        slice_range(args.num_kb, args.k_slices, slice, kb0, nkb);
        if (BSTAT && seq == 0) {
          mbar_wait(bfull_bar, bphase);                  // this n-tile's B blocks have landed
          bphase ^= 1u;
        }
        // Store epilogue: a ragged last n-tile only multiplies the columns that exist (N rounded up to 16) -- dX of the
        // log-linear model has 320 columns, and a full-width second tile spent 37 % of the GEMM's MMA time on padding
        uint32_t idesc = idesc_full;
        if (MODE == TC_EPI_STORE) {
          long long n_left = args.n_end - (args.n_begin + (long long)slice_nt(nt, num_n_tiles, slice) * TN * args.epi.tile_stride);
          if (n_left > TN) n_left = TN;
          if (n_left < BN) idesc = make_idesc(BM, (int)((n_left + 15) / 16 * 16));
        }
        if (CG2) idesc = (idesc & ~(0x1fu << 24)) | ((uint32_t)(2 * BM >> 4) << 24);   // M = 256 over the pair
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
        if (!BSTAT && args.pair_kp > 0) {
          // pair operands: stage s holds (A_hi, B_hi), stage s + 1 (A_mid, B_mid) of one 64-column block (s is even:
          // every block takes two steps of the 4-stage ring)
          static_assert(STAGES % 2 == 0, "pair operands take the ring's stages two at a time");
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(full_bar(stage), phase);
            mbar_wait(full_bar(stage + 1), phase);
            tc_fence_after();
            const uint32_t a_hi = smem_a + stage * A_STAGE_BYTES, a_mid = a_hi + A_STAGE_BYTES;
            const uint32_t b_hi = smem_b + stage * B_STAGE_BYTES, b_mid = b_hi + B_STAGE_BYTES;
            if (args.b_nmajor) {
              const uint32_t idn = idesc | IDESC_B_MN_MAJOR;
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                mma(d_tmem, make_smem_desc(a_hi + k * UMMA_K * 2), make_smem_desc_mn(b_hi + k * UMMA_K * 128), idn,
                            (kb | k) != 0 ? 1u : 0u);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                mma(d_tmem, make_smem_desc(a_hi + k * UMMA_K * 2), make_smem_desc_mn(b_mid + k * UMMA_K * 128), idn, 1u);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                mma(d_tmem, make_smem_desc(a_mid + k * UMMA_K * 2), make_smem_desc_mn(b_hi + k * UMMA_K * 128), idn, 1u);
            } else {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              mma(d_tmem, make_smem_desc(a_hi + k * UMMA_K * 2), make_smem_desc(b_hi + k * UMMA_K * 2), idesc,
                          (kb | k) != 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              mma(d_tmem, make_smem_desc(a_hi + k * UMMA_K * 2), make_smem_desc(b_mid + k * UMMA_K * 2), idesc, 1u);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              mma(d_tmem, make_smem_desc(a_mid + k * UMMA_K * 2), make_smem_desc(b_hi + k * UMMA_K * 2), idesc, 1u);
            }
            release_stage(stage);
            release_stage(stage + 1);
            stage += 2;
            if (stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        } else
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);             // TMA bytes have landed
          tc_fence_after();
          const uint32_t a_addr = smem_a + stage * A_STAGE_BYTES;
          const uint32_t b_addr = smem_b + (BSTAT ? kb : stage) * B_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advancing K by 16 bf16 = 32 bytes inside the 128-byte swizzled row
            mma(d_tmem, make_smem_desc(a_addr + k * UMMA_K * 2), make_smem_desc(b_addr + k * UMMA_K * 2),
                        idesc, (kb | k) != 0 ? 1u : 0u);
          }
          release_stage(stage);                          // frees the smem slot once those MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (CG2) tc_commit_cg2(tfull_bar(acc), (uint16_t)0x3);   // both CTAs' epilogues read their halves
        else tc_commit(tfull_bar(acc));                  // accumulator complete -> epilogue
        if (BSTAT && seq == num_m_tiles - 1) tc_commit(bempty_bar);   // ... and the B slots are free once it is
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ================= epilogue (warps 2..9) =================
    const int quarter = warp & 3;                        // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int col_lo = ((warp - 2) >> 2) * COLS_PER_WARP;   // ... and this warp handles columns [col_lo, +COLS_PER_WARP)
    constexpr int CHUNKS = COLS_PER_WARP / 32;
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    const TcEpilogue &ep = args.epi;
    // top-k epilogue: the keys parked by the PREVIOUS tile wait for their list slots (pend_pos is the result of an
    // atomicAdd issued one tile ago: its ~1 us round trip runs under this tile's accumulator read).  The warp's
    // keys sit compacted in shared memory; entry i belongs to row (pend_m0 + quarter*32 + owner) and goes to slot
    // pend_pos[owner] + ordinal of that row's list, so the copy-out is one coalesced loop over the entries.
    int pend_wtotal = 0, pend_pos = 0, pend_m0 = 0, stash_buf = 0;
    const uint32_t lane_lt = (1u << lane) - 1u;
    auto stash_keys = [&](int buf) { return stash_base + (uint32_t)((buf * EPI_WARPS + (warp - 2)) * WSTASH) * 8u; };
    auto stash_meta = [&](int buf) {
      return stash_base + STASH_KEYS_BYTES + (uint32_t)((buf * EPI_WARPS + (warp - 2)) * WSTASH) * 2u;
    };
    auto flush_pending = [&]() {
      if (pend_wtotal > 0) {                              // warp-uniform
        const uint32_t keys = stash_keys(stash_buf ^ 1), meta = stash_meta(stash_buf ^ 1);
        for (int i0 = 0; i0 < pend_wtotal; i0 += 32) {
          const int i = i0 + lane;
          unsigned long long key = 0ull;
          uint32_t mt16 = 0u;
          if (i < pend_wtotal) {
            asm volatile("ld.shared.b64 %0, [%1];" : "=l"(key) : "r"(keys + (uint32_t)i * 8u) : "memory");
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(mt16) : "r"(meta + (uint32_t)i * 2u) : "memory");
          }
          const int owner = (int)(mt16 >> 11), ord = (int)(mt16 & 0x7ffu);
          const int pos = __shfl_sync(0xffffffffu, pend_pos, owner) + ord;
          if (i < pend_wtotal) {
            if (pos < ep.cap) ep.cand[(size_t)(pend_m0 + quarter * 32 + owner) * ep.cap + pos] = key;
            else *ep.overflow = 1;
          }
        }
      }
      pend_wtotal = 0;
    };
    int mt, nt, seq, slice;
    for (int it = 0; cta_tile<BSTAT>(it, bid, gdim, walk_m_tiles, num_n_tiles, args.k_slices, step_m, step_n, mt, nt, seq, slice, n_fast); ++it) {
      const int m0 = (mt * CL + (int)cta_rank) * BM;
      const int nt_in = slice_nt(nt, num_n_tiles, slice);
      const long long n0 = args.n_begin + (long long)nt_in * TN * args.epi.tile_stride;
      // store epilogue: the tile's columns end at tile_end (tile_n may be narrower than the 256 accumulator columns)
      const long long tile_end = MODE == TC_EPI_STORE && n0 + TN < args.n_end ? n0 + TN : args.n_end;
      const int gm = m0 + row;
      const bool row_ok = gm < args.M;
      // Rows are swept in increasing id order and tau only moves between launches, so a later row that
      // ties tau's score has a larger id (= a smaller key): `score > tau_score` is the exact key test.
      float tau_score = INFINITY;
      if (MODE == TC_EPI_TOPK && row_ok) {
        const unsigned long long tau_key = ep.tau[gm];
        tau_score = tau_key == 0ull ? -INFINITY : key_score(tau_key);
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t)acc * BN + ((uint32_t)(quarter * 32) << 16);
      if (MODE == TC_EPI_STORE) {
        float st_max = -INFINITY, st_sum = 0.f;             // running softmax statistics of this lane's 64 columns
#pragma unroll 1
        for (int c0 = col_lo; c0 < col_lo + COLS_PER_WARP; c0 += 32) {
          uint32_t v[32];
          tc_ld_32x32(t_row + (uint32_t)c0, v);
          tc_wait_ld();
          const long long gn0 = n0 + c0;
          // Whole chunks of a 16-byte-aligned C go out TRANSPOSED inside groups of 8 lanes (warp-uniform condition):
          // with one row per lane, a 16-byte store instruction touches 32 different 128-byte lines (32 L1 wavefronts:
          // ~8 k cycles of store issue per 128 x 256 tile, the whole MMA time of the log-linear projection, whose
          // epilogue writes the 8 GB logit matrix); after an 8 x 8 transpose of 16-byte pieces, 8 lanes write one
          // row's 128 contiguous bytes and an instruction touches 4 lines.
          const bool t_store = SERT_TC_TSTORE && gn0 + 32 <= tile_end && ep.extra_row < 0 && (ep.ldc & 3) == 0 &&
                               (reinterpret_cast<uintptr_t>(ep.C) & 15) == 0;
          if (t_store) {
            // ---- whole chunk, every lane takes part (rows beyond M carry garbage that is never stored): no per-element
            // or per-lane predicates anywhere, so the shuffles below sit in warp-uniform code
            float w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = __uint_as_float(v[j]);
            if (ep.bias != nullptr) {
              if ((reinterpret_cast<uintptr_t>(ep.bias + gn0) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {              // broadcast 16-byte loads: every lane adds the same bias
                  const float4 bv = __ldg(reinterpret_cast<const float4 *>(ep.bias + gn0 + j));
                  w[j] += bv.x; w[j + 1] += bv.y; w[j + 2] += bv.z; w[j + 3] += bv.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) w[j] += __ldg(ep.bias + gn0 + j);
              }
            }
            if (ep.row_stats != nullptr) {
              float mx = w[0];
#pragma unroll
              for (int j = 1; j < 32; ++j) mx = fmaxf(mx, w[j]);
              const float m_new = fmaxf(st_max, mx);
              // exp(w - m) = 2^(w log2e - m log2e): one FFMA and one MUFU per value
              const float kLog2e = 1.4426950408889634f;
              const float ml = m_new * kLog2e;
              float add = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) add += fast_exp2(fmaf(w[j], kLog2e, -ml));
              st_sum = st_sum * __expf(st_max - m_new) + add;
              st_max = m_new;
            }
            float4 f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = make_float4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
#pragma unroll
            for (int mk = 1; mk < 8; mk <<= 1) {
              const bool upper = (lane & mk) != 0;
#pragma unroll
              for (int sl = 0; sl < 8; ++sl) {
                if (sl & mk) continue;
                const float4 send = upper ? f[sl] : f[sl | mk];
                float4 recv;
                recv.x = __shfl_xor_sync(0xffffffffu, send.x, mk);
                recv.y = __shfl_xor_sync(0xffffffffu, send.y, mk);
                recv.z = __shfl_xor_sync(0xffffffffu, send.z, mk);
                recv.w = __shfl_xor_sync(0xffffffffu, send.w, mk);
                if (upper) f[sl] = recv; else f[sl | mk] = recv;
              }
            }
            // lane 8g + j now holds, in f[i], columns [4j, 4j + 4) of the row lane 8g + i read from TMEM
            const int r0 = m0 + quarter * 32 + (lane & ~7);
            float *base = ep.C + (long long)r0 * ep.ldc + gn0 + 4 * (lane & 7);
            const int rows_here = args.M - r0;               // the same for the 8 lanes of a group
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (i < rows_here) {
                if (ep.accumulate) red_add_f4(base + (long long)i * ep.ldc, f[i]);
                else *reinterpret_cast<float4 *>(base + (long long)i * ep.ldc) = f[i];
              }
            }
          } else if (row_ok && gn0 < tile_end) {
            const int nv = (int)(tile_end - gn0 < 32 ? tile_end - gn0 : 32);
            float w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = __uint_as_float(v[j]);
            if (ep.bias != nullptr) {
              if (nv == 32 && (reinterpret_cast<uintptr_t>(ep.bias + gn0) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {              // broadcast 16-byte loads: every lane adds the same bias
                  const float4 bv = __ldg(reinterpret_cast<const float4 *>(ep.bias + gn0 + j));
                  w[j] += bv.x; w[j + 1] += bv.y; w[j + 2] += bv.z; w[j + 3] += bv.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < nv) w[j] += __ldg(ep.bias + gn0 + j);
              }
            }
            if (ep.row_stats != nullptr) {
              float mx = -INFINITY;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nv) mx = fmaxf(mx, w[j]);
              const float m_new = fmaxf(st_max, mx);
              float add = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nv) add += __expf(w[j] - m_new);
              st_sum = st_sum * __expf(st_max - m_new) + add;
              st_max = m_new;
            }
            float *dst = gm == ep.extra_row ? ep.extra_dst + gn0 : ep.C + (long long)gm * ep.ldc + gn0;
            if (nv == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 o = make_float4(w[j], w[j + 1], w[j + 2], w[j + 3]);
                if (ep.accumulate) red_add_f4(dst + j, o);
                else *reinterpret_cast<float4 *>(dst + j) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nv) {
                  if (ep.accumulate) atomicAdd(dst + j, w[j]);
                  else dst[j] = w[j];
                }
            }
          }
        }
        if (ep.row_stats != nullptr && row_ok && n0 + col_lo < args.n_end)
          ep.row_stats[(size_t)gm * ep.stats_ld + (size_t)((n0 - args.n_begin) + col_lo) / COLS_PER_WARP] =
              make_float2(st_max, st_sum);
      } else if (MODE == TC_EPI_GROUPMAX) {
        // ---- maxima over groups of 8 or 64 columns (full tiles only): the sample the scoring sweep seeds its
        // thresholds from (score.cu: seed_tau_kernel)
        float mx = -INFINITY;
#pragma unroll 1
        for (int ci = 0; ci < CHUNKS; ++ci) {
          uint32_t v[32];
          tc_ld_32x32(t_row + (uint32_t)(col_lo + ci * 32), v);
          tc_wait_ld();
          float m4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a = fmaxf(fmaxf(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                                  fmaxf(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])));
            const float b = fmaxf(fmaxf(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                                  fmaxf(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
            m4[j] = fmaxf(a, b);
          }
          float *dst = ep.gmax + (size_t)gm * ep.gmax_ld + (size_t)nt_in * (BN / ep.group) + (col_lo + ci * 32) / ep.group;
          if (ep.group == 8) {
            if (row_ok) *reinterpret_cast<float4 *>(dst) = make_float4(m4[0], m4[1], m4[2], m4[3]);
          } else if (ep.group == 16) {
            if (row_ok) *reinterpret_cast<float2 *>(dst) = make_float2(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          } else if (ep.group == 32) {
            if (row_ok) *dst = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          } else {
            mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
          }
        }
        if (ep.group == COLS_PER_WARP && row_ok)
          ep.gmax[(size_t)gm * ep.gmax_ld + (size_t)nt_in * (BN / COLS_PER_WARP) + col_lo / COLS_PER_WARP] = mx;
      } else {
        // ---- running top-k filter.  The accumulator is read ONCE, eight columns at a time (tcgen05.ld x8, the next
        // group in flight while the current one is examined), in a ROLLED loop: the epilogue's code must stay in the
        // instruction caches -- a fully unrolled two-pass version measured 6x slower ("no instruction" stalls), and the
        // 32-column form of this pass, with its four quarter bodies unrolled, still spent a quarter of its issue
        // slots there (profiles/ncu_gemm_tc_r2c_*).  Per group: a max tree and ONE ballot decide whether any of the
        // warp's 32 rows has a survivor (usually not); if so every lane counts its survivors and remembers the last
        // one's column; when no row has two (a row's survivor then IS its group maximum) they are compacted into the
        // warp's shared-memory stash with one more ballot, else the group is scanned column by column.
        int total = 0;                                                   // this lane's survivors in the tile
        int wcount = 0;                                                  // the warp's (uniform)
        uint32_t group_any = 0;                                          // warp-uniform: 8-column groups with a survivor
        const uint32_t my_keys = stash_keys(stash_buf), my_meta = stash_meta(stash_buf);
        const unsigned int col_base = (unsigned int)(n0 + col_lo + ep.row_offset);
        const long long left_all = args.n_end - (n0 + col_lo);           // valid columns of this warp's slice
        auto examine = [&](const uint32_t (&v)[GW], int g) {
          float mx = __uint_as_float(v[0]);
#pragma unroll
          for (int j = 1; j < GW; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
          if (__ballot_sync(0xffffffffu, mx > tau_score) == 0u) return;          // nothing here (the common case)
          const int left = (int)(left_all - g * GW < GW ? left_all - g * GW : GW);   // < GW only in the shard's last tile
          int cnt = 0, idx = 0;
#pragma unroll
          for (int j = 0; j < GW; ++j) {
            const bool hit = __uint_as_float(v[j]) > tau_score && j < left;
            cnt += hit ? 1 : 0;
            idx = hit ? j : idx;
          }
          const uint32_t h1 = __ballot_sync(0xffffffffu, cnt >= 1);
          if (h1 == 0u) return;                                                  // only padding columns beat tau
          group_any |= 1u << g;
          if (left == GW && __ballot_sync(0xffffffffu, cnt >= 2) == 0u) {
            if (cnt != 0) {
              const int slot = wcount + __popc(h1 & lane_lt);
              if (slot < WSTASH) {
                const unsigned long long key = make_key(mx, col_base + (unsigned int)(g * GW + idx));
                asm volatile("st.shared.b64 [%0], %1;" ::"r"(my_keys + (uint32_t)slot * 8u), "l"(key) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(my_meta + (uint32_t)slot * 2u),
                             "h"((unsigned short)((lane << 11) | (total & 0x7ff)))
                             : "memory");
              }
              ++total;
            }
            wcount += __popc(h1);
            return;
          }
#pragma unroll
          for (int j = 0; j < GW; ++j) {
            const bool hit = __uint_as_float(v[j]) > tau_score && j < left;
            const uint32_t hm = __ballot_sync(0xffffffffu, hit);
            if (hm == 0u) continue;
            if (hit) {
              const int slot = wcount + __popc(hm & lane_lt);
              if (slot < WSTASH) {
                const unsigned long long key = make_key(__uint_as_float(v[j]), col_base + (unsigned int)(g * GW + j));
                asm volatile("st.shared.b64 [%0], %1;" ::"r"(my_keys + (uint32_t)slot * 8u), "l"(key) : "memory");
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(my_meta + (uint32_t)slot * 2u),
                             "h"((unsigned short)((lane << 11) | (total & 0x7ff)))
                             : "memory");
              }
              ++total;
            }
            wcount += __popc(hm);
          }
        };
        {
          constexpr int GROUPS = COLS_PER_WARP / GW;
          static_assert(GROUPS % 2 == 0, "the group loop is unrolled by two");
          const uint32_t tbase = t_row + (uint32_t)col_lo;
          uint32_t va[GW], vb[GW];
          tc_ld_group<GW>(tbase, va);
#pragma unroll 1
          for (int g = 0; g < GROUPS; g += 2) {
            tc_wait_ld();
            tc_ld_group<GW>(tbase + (uint32_t)((g + 1) * GW), vb);
            examine(va, g);
            tc_wait_ld();
            if (g + 2 < GROUPS) tc_ld_group<GW>(tbase + (uint32_t)((g + 2) * GW), va);
            examine(vb, g + 1);
          }
        }
        uint32_t chunk_any = 0;                                          // 32-column chunks with a survivor (two-pass path)
#pragma unroll
        for (int ci = 0; ci < CHUNKS; ++ci)
          chunk_any |= ((group_any >> ((32 / GW) * ci)) & ((1u << (32 / GW)) - 1u)) ? (1u << ci) : 0u;
        if (wcount <= WSTASH) {
          // Everything this warp keeps sits in shared memory: hand the accumulator back to the MMA warp, copy out
          // the previous tile's keys (their slots were reserved a tile ago), reserve this tile's slots.
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
          flush_pending();                                // previous tile's keys: buffer stash_buf ^ 1
          if (total > 0) pend_pos = atomicAdd(ep.count + gm, total);
          pend_wtotal = wcount;
          pend_m0 = m0;
          stash_buf ^= 1;                                 // this tile's keys now sit in the "previous" buffer
          continue;
        }
        // cold tau: finish the previous tile's keys first, then the two-pass path below
        flush_pending();
        // Pass 2 (cold tau: the first chunks of a sweep): ONE atomic per thread reserves its slots, then TMEM is
        // read again.  tcgen05.ld is warp-collective (.sync.aligned), so the chunk loop is warp-uniform; only the
        // per-lane key stores diverge.
        if (chunk_any != 0u) {
          int pos = 0;
          unsigned long long *list = ep.cand + (size_t)(row_ok ? gm : 0) * ep.cap;
          if (total > 0) {
            pos = atomicAdd(ep.count + gm, total);
            if (pos + total > ep.cap) *ep.overflow = 1;
          }
#pragma unroll 1
          for (int ci = 0; ci < CHUNKS; ++ci) {
            if (!((chunk_any >> ci) & 1u)) continue;                    // warp-uniform skip
            uint32_t v[32];
            tc_ld_32x32(t_row + (uint32_t)(col_lo + ci * 32), v);
            tc_wait_ld();
            if (total == 0) continue;                                   // this lane has nothing to write
            const long long left = args.n_end - (n0 + col_lo + ci * 32);
            const int nv = left >= 32 ? 32 : (left <= 0 ? 0 : (int)left);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nv && __uint_as_float(v[j]) > tau_score) {
                if (pos < ep.cap)
                  list[pos] = make_key(__uint_as_float(v[j]),
                                       (unsigned int)(n0 + col_lo + ci * 32 + j + ep.row_offset));
                ++pos;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG2) mbar_arrive_cluster(mapa_cluster(tempty_bar(acc), 0u));   // the leader's MMA warp waits for both CTAs
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    flush_pending();
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();            // no CTA leaves while a peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if (CG2) tc_dealloc_cg2(tmem_base, TMEM_COLS);
    else tc_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn *out) {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SERT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    SERT_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  *out = fn;
  return 0;
}

// 2-D bf16 tensor (rows, Kt) row-major with row stride ld elements; box = (64 elements of K) x box_rows, 128-byte
// swizzle, zero OOB fill.
int make_map(const __nv_bfloat16 *base, long long rows, long long Kt, long long ld, int box_rows, TcMap *out) {
  static_assert(sizeof(CUtensorMap) <= sizeof(TcMap), "CUtensorMap does not fit");
  EncodeTiledFn encode;
  if (get_encode_fn(&encode)) return -1;
  SERT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor-map base must be 16-byte aligned");
  SERT_REQUIRE(Kt % BK == 0, "K must be padded to a multiple of 64");
  const cuuint64_t gdim[2] = {(cuuint64_t)Kt, (cuuint64_t)std::max<long long>(rows, 1)};
  SERT_REQUIRE(ld >= Kt && ld % 8 == 0, "row stride must cover K and be a multiple of 16 bytes");
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(__nv_bfloat16)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = encode(reinterpret_cast<CUtensorMap *>(out->bytes), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                            const_cast<__nv_bfloat16 *>(base), gdim, gstride, box, estride,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SERT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
  return 0;
}

}  // namespace

int launch_gemm_tc(const __nv_bfloat16 *A, int M, const __nv_bfloat16 *B, long long N_total, long long n_begin,
                   long long n_end, int Kt, const TcEpilogue &epi, cudaStream_t st) {
  return launch_gemm_tc_ld(A, Kt, M, B, Kt, N_total, n_begin, n_end, Kt, epi, st);
}

static int launch_gemm_tc_impl(const __nv_bfloat16 *A, long long lda, int M, const __nv_bfloat16 *B, long long ldb,
                               long long N_total, long long n_begin, long long n_end, int Kt, int pair_kp,
                               const TcEpilogue &epi, cudaStream_t st, long long bn_rows = -1, long long bn_cols = 0);

int launch_gemm_tc_ld(const __nv_bfloat16 *A, long long lda, int M, const __nv_bfloat16 *B, long long ldb,
                      long long N_total, long long n_begin, long long n_end, int Kt, const TcEpilogue &epi,
                      cudaStream_t st) {
  return launch_gemm_tc_impl(A, lda, M, B, ldb, N_total, n_begin, n_end, Kt, 0, epi, st);
}

int launch_gemm_tc_pair(const __nv_bfloat16 *A, int M, const __nv_bfloat16 *B, long long N_total, long long n_begin,
                        long long n_end, int Kp, const TcEpilogue &epi, cudaStream_t st) {
  SERT_REQUIRE(epi.mode == TC_EPI_STORE, "pair operands serve the store epilogue");
  return launch_gemm_tc_impl(A, 2ll * Kp, M, B, 2ll * Kp, N_total, n_begin, n_end, 2 * Kp, Kp, epi, st);
}

int launch_gemm_tc_pair_bn(const __nv_bfloat16 *A, int M, const __nv_bfloat16 *Bn, long long k_rows, long long n_pad,
                           long long N_total, long long n_begin, long long n_end, int Kp, const TcEpilogue &epi,
                           cudaStream_t st) {
  SERT_REQUIRE(epi.mode == TC_EPI_STORE, "pair operands serve the store epilogue");
  SERT_REQUIRE(n_pad % 64 == 0 && N_total <= n_pad && k_rows <= Kp, "N-major B: blocks of n_pad columns, K rows within Kp");
  return launch_gemm_tc_impl(A, 2ll * Kp, M, Bn, 2 * n_pad, N_total, n_begin, n_end, 2 * Kp, Kp, epi, st, k_rows, n_pad);
}

// Kt = columns of the operands the tensor maps cover; pair_kp > 0: [hi | mid] operands of 2 * pair_kp columns;
// bn_rows >= 0: B is N-major, bn_rows K rows of [hi | mid] blocks of bn_cols columns (row stride ldb)
static int launch_gemm_tc_impl(const __nv_bfloat16 *A, long long lda, int M, const __nv_bfloat16 *B, long long ldb,
                               long long N_total, long long n_begin, long long n_end, int Kt, int pair_kp,
                               const TcEpilogue &epi, cudaStream_t st, long long bn_rows, long long bn_cols) {
  if (M == 0 || n_end <= n_begin) return 0;
  SERT_REQUIRE(pair_kp == 0 || (pair_kp % BK == 0 && Kt == 2 * pair_kp), "pair operands: two blocks of Kp columns");
  SERT_REQUIRE(n_begin >= 0 && n_end <= N_total && N_total < (1ll << 31), "bad column range");
  SERT_REQUIRE(Kt > 0 && Kt % BK == 0, "K must be a positive multiple of 64");
  const bool b_nmajor = bn_rows >= 0;
  SERT_REQUIRE(!b_nmajor || pair_kp > 0, "N-major B needs pair operands");
  TcMap ma, mb, mb2;
  if (make_map(A, M, Kt, lda, BM, &ma)) return -1;
  if (b_nmajor) {
    // boxes of 64 N columns x 64 K rows over the hi block (columns [0, bn_cols)) and the mid block behind it
    if (make_map(B, bn_rows, bn_cols, ldb, 64, &mb)) return -1;
    if (make_map(B + bn_cols, bn_rows, bn_cols, ldb, 64, &mb2)) return -1;
  } else {
    if (make_map(B, N_total, Kt, ldb, BN, &mb)) return -1;
    mb2 = mb;
  }
  static std::atomic<uint64_t> configured{0};
  if (first_use_on_device(configured)) {
    SERT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false, TC_EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SERT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false, TC_EPI_TOPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SERT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false, TC_EPI_GROUPMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SERT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<true, TC_EPI_TOPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SERT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false, TC_EPI_STORE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SERT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false, TC_EPI_STORE, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SERT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false, TC_EPI_STORE, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  }
  int dev = 0, sms = kNumSMs;
  SERT_CUDA(cudaGetDevice(&dev));
  SERT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  KernelArgs args;
  args.M = M;
  args.n_begin = n_begin;
  args.n_end = n_end;
  args.num_kb = (pair_kp > 0 ? pair_kp : Kt) / BK;
  args.k_slices = 1;
  args.pair_kp = pair_kp;
  args.n_fast = 0;
  args.tile_n = BN;
  args.b_nmajor = b_nmajor ? 1 : 0;
  args.epi = epi;
  const long long m_tiles = (M + BM - 1) / BM;
  long long n_tiles = (n_end - n_begin + BN - 1) / BN;
  {
    // a narrow, ragged C (2-4 n-tiles) over many m-tiles: n-tiles of equal width (KernelArgs::tile_n)
    const char *tn = getenv("SERT_GEMM_EQUAL_TILES");
    if (epi.mode == TC_EPI_STORE && epi.row_stats == nullptr && epi.tile_stride == 1 && n_tiles >= 2 && n_tiles <= 4 &&
        m_tiles >= 8 && (n_end - n_begin) % BN != 0 && !(tn != nullptr && tn[0] == '0')) {
      const long long even = ((n_end - n_begin + n_tiles - 1) / n_tiles + 15) / 16 * 16;
      if ((n_end - n_begin + even - 1) / even == n_tiles) args.tile_n = (int)even;
    }
  }
  const long long tiles = m_tiles * n_tiles;
  // B-stationary schedule (top-k epilogue; SERT_GEMM_BSTAT=0 turns it off): K fits the ring's B slots, every CTA gets
  // an n-tile, enough m-tiles to amortise the load of B (one bubble per n-tile), and whole n-tiles balance at least as
  // well as single tiles would.  It cuts the operand traffic of a tile from 192 KB to 64 KB at K = 256.  Round 1
  // measured it 2-3 % SLOWER (8.86 vs 8.62 ms at BASELINE configs[3]) because the top-k epilogue paced the tiles;
  // with the round-2 epilogue it is 4 % faster (4.60 vs 4.79 ms, same box, same run), so it is on by default.
  const long long per_cta = (n_tiles + sms - 1) / sms;
  const char *bstat_env = getenv("SERT_GEMM_BSTAT");
  const bool bstat = pair_kp == 0 && args.num_kb <= STAGES && m_tiles >= 8 && n_tiles >= sms &&
                     per_cta * sms * 10 <= n_tiles * 12 && !(bstat_env != nullptr && bstat_env[0] == '0');
  if (bstat && epi.mode == TC_EPI_TOPK) {
    gemm_tc_kernel<true, TC_EPI_TOPK><<<(int)std::min<long long>(n_tiles, sms), NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, mb2, args);
    SERT_LAUNCH_CHECK();
    return 0;
  }
  long long work = tiles;
  if (epi.mode == TC_EPI_STORE && epi.accumulate && epi.bias == nullptr) {
    // split-K: few output tiles and a very deep K (dX = dZ . Wd^T of the log-linear model: 160 tiles, K = 600 k) leave
    // the chip in two unequal waves; slices of the K range become tiles of their own and meet in C by reduction
    int slices = 1;
    const int kb_weight = pair_kp > 0 ? 3 : 1;           // a pair block carries three blocks' worth of MMAs
    while (tiles * slices < 4ll * sms && args.num_kb * kb_weight / (slices * 2) >= 32) slices *= 2;
    const int per = (args.num_kb + slices - 1) / slices;
    args.k_slices = (args.num_kb + per - 1) / per;       // no empty slice
    work = tiles * args.k_slices;
  }
  {
    // few n-tiles over a deep, tall A (dX = dZ . Wd^T: 2 n-tiles, A = 8 GB): keep the n-tiles of an A tile together
    const char *nf = getenv("SERT_GEMM_NFAST");
    if (epi.mode == TC_EPI_STORE && n_tiles >= 2 && n_tiles <= 4 && m_tiles >= 8 && !(nf != nullptr && nf[0] == '0'))
      args.n_fast = 1;
  }
  const int grid = (int)std::min<long long>(work, sms);
  // Clusters of 2 or 4 CTAs with a multicast B tile (store epilogue): a third / half less operand traffic through L2
  // for the GEMMs whose K is too deep for the B-stationary schedule (the log-linear projection and its gradients).
  // Needs enough m-tiles that grouping them wastes little (a ragged count pads the last group).
  // Measured at BASELINE configs[4] (one box, one session): step 29.9 ms without clusters, 27.9 ms with pairs, 32.7 ms
  // with clusters of four (the lockstep of four CTAs costs more than the traffic saves), so pairs are the default.
  // SERT_GEMM_CLUSTER = 0: off, 4: clusters of four where the m-tiles allow.
  const char *cl_env = getenv("SERT_GEMM_CLUSTER");
  int cl = 1;
  if (epi.mode == TC_EPI_STORE && m_tiles >= 8 && !b_nmajor && !(cl_env != nullptr && cl_env[0] == '0')) {
    if (cl_env != nullptr && cl_env[0] == '4' && (m_tiles % 4 == 0 || m_tiles >= 64) && sms % 4 == 0) cl = 4;
    else if (m_tiles % 2 == 0 || m_tiles >= 32) cl = 2;
  }
  if (cl > 1) {
    TcMap mb_part;
    if (make_map(B, N_total, Kt, ldb, BN / cl, &mb_part)) return -1;
    const long long group_work = ((m_tiles + cl - 1) / cl) * n_tiles * args.k_slices;
    const int clusters = (int)std::min<long long>(group_work, sms / cl);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl * clusters);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static const char *cg2_env = getenv("SERT_GEMM_CG2");
    const bool cg2 = cl == 2 && !(cg2_env != nullptr && cg2_env[0] == '0');   // two-SM MMA for the pairs (SERT_GEMM_CG2=0: two single-SM MMAs over a multicast B tile)
    if (cg2) SERT_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<false, TC_EPI_STORE, 2, true>, ma, mb_part, mb_part, args));
    else if (cl == 4) SERT_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<false, TC_EPI_STORE, 4>, ma, mb_part, mb_part, args));
    else SERT_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<false, TC_EPI_STORE, 2>, ma, mb_part, mb_part, args));
    count_launch();
    return 0;
  }
  if (epi.mode == TC_EPI_TOPK) gemm_tc_kernel<false, TC_EPI_TOPK><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, mb2, args);
  else if (epi.mode == TC_EPI_GROUPMAX) gemm_tc_kernel<false, TC_EPI_GROUPMAX><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, mb2, args);
  else gemm_tc_kernel<false, TC_EPI_STORE><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, mb2, args);
  SERT_LAUNCH_CHECK();
  return 0;
}

// ---- fp32 -> bf16 split operand (see gemm_tc.cuh) --------------------------------------------------
__global__ void __launch_bounds__(256) split_bf16_kernel(const float *__restrict__ src, long long rows, int K,
                                                         long long ld_src, int Kp, int terms, int role,
                                                         __nv_bfloat16 *__restrict__ dst) {
  const long long total = rows * Kp;
  const long long ld_dst = (long long)terms * Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Kp;
    const int k = (int)(i - r * Kp);
    const float x = k < K ? src[r * ld_src + k] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    __nv_bfloat16 *d = dst + r * ld_dst + k;
    if (terms == 1) {
      d[0] = hi;
    } else {
      const __nv_bfloat16 mid = __float2bfloat16_rn(x - __bfloat162float(hi));
      // terms 3: A'' = [hi | hi | mid],  B'' = [hi | mid | hi];  terms 2 (pair operands): [hi | mid]
      d[0] = hi;
      if (terms == 2) {
        d[Kp] = mid;
      } else {
        d[Kp] = role == SPLIT_A ? hi : mid;
        d[2 * (long long)Kp] = role == SPLIT_A ? mid : hi;
      }
    }
  }
}

int launch_split_bf16(const float *src, long long rows, int K, long long ld_src, int terms, SplitRole role,
                      __nv_bfloat16 *dst, cudaStream_t st) {
  SERT_REQUIRE(terms >= 1 && terms <= 3, "split terms must be 1, 2 (pair operands) or 3");
  if (rows == 0) return 0;
  const int Kp = tc_padded_k(K);
  const long long total = rows * Kp;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)kNumSMs * 16);
  split_bf16_kernel<<<blocks, 256, 0, st>>>(src, rows, K, ld_src, Kp, terms, (int)role, dst);
  SERT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sert

// ---- test hook: C = A . B^T through the tcgen05 path, host in / host out ----------------------------
extern "C" SERT_API int sert_debug_gemm_tc(const float *a_host, const float *b_host, int m, int n, int k, int terms,
                                           const float *bias_host, float *c_host) {
  using namespace sert;
  SERT_REQUIRE(a_host && b_host && c_host && m > 0 && n > 0 && k > 0, "bad argument");
  const int Kp = tc_padded_k(k);
  const int Kt = terms * Kp;
  float *dA = nullptr, *dB = nullptr, *dC = nullptr, *dbias = nullptr;
  __nv_bfloat16 *sA = nullptr, *sB = nullptr;
  SERT_CUDA(cudaMalloc(&dA, (size_t)m * k * 4));
  SERT_CUDA(cudaMalloc(&dB, (size_t)n * k * 4));
  SERT_CUDA(cudaMalloc(&dC, (size_t)m * n * 4));
  SERT_CUDA(cudaMalloc(&sA, (size_t)m * Kt * 2));
  SERT_CUDA(cudaMalloc(&sB, (size_t)n * Kt * 2));
  SERT_CUDA(cudaMemcpy(dA, a_host, (size_t)m * k * 4, cudaMemcpyHostToDevice));
  SERT_CUDA(cudaMemcpy(dB, b_host, (size_t)n * k * 4, cudaMemcpyHostToDevice));
  SERT_CUDA(cudaMemset(dC, 0xff, (size_t)m * n * 4));
  if (bias_host) {
    SERT_CUDA(cudaMalloc(&dbias, (size_t)n * 4));
    SERT_CUDA(cudaMemcpy(dbias, bias_host, (size_t)n * 4, cudaMemcpyHostToDevice));
  }
  int rc = launch_split_bf16(dA, m, k, k, terms, SPLIT_A, sA, nullptr);
  if (!rc) rc = launch_split_bf16(dB, n, k, k, terms, SPLIT_B, sB, nullptr);
  TcEpilogue ep;
  ep.mode = TC_EPI_STORE;
  ep.C = dC;
  ep.ldc = n;
  ep.bias = dbias;
  if (!rc) rc = terms == 2 ? launch_gemm_tc_pair(sA, m, sB, n, 0, n, Kp, ep, nullptr)
                           : launch_gemm_tc(sA, m, sB, n, 0, n, Kt, ep, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (!rc && e != cudaSuccess) {
    set_error(std::string("gemm_tc: ") + cudaGetErrorString(e));
    rc = -1;
  }
  if (!rc) SERT_CUDA(cudaMemcpy(c_host, dC, (size_t)m * n * 4, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(sA); cudaFree(sB); cudaFree(dbias);
  return rc;
}

// ---- test hook: C = A . Bt with Bt given (k, n) row-major -- the N-major B operand of the pair path ------------------
extern "C" SERT_API int sert_debug_gemm_tc_bn(const float *a_host, const float *bt_host, int m, int n, int k, float *c_host) {
  using namespace sert;
  SERT_REQUIRE(a_host && bt_host && c_host && m > 0 && n > 0 && k > 0, "bad argument");
  const int Kp = tc_padded_k(k), Np = tc_padded_k(n);
  float *dA = nullptr, *dB = nullptr, *dC = nullptr;
  __nv_bfloat16 *sA = nullptr, *sB = nullptr;
  SERT_CUDA(cudaMalloc(&dA, (size_t)m * k * 4));
  SERT_CUDA(cudaMalloc(&dB, (size_t)n * k * 4));
  SERT_CUDA(cudaMalloc(&dC, (size_t)m * n * 4));
  SERT_CUDA(cudaMalloc(&sA, (size_t)m * 2 * Kp * 2));
  SERT_CUDA(cudaMalloc(&sB, (size_t)k * 2 * Np * 2));
  SERT_CUDA(cudaMemcpy(dA, a_host, (size_t)m * k * 4, cudaMemcpyHostToDevice));
  SERT_CUDA(cudaMemcpy(dB, bt_host, (size_t)n * k * 4, cudaMemcpyHostToDevice));
  SERT_CUDA(cudaMemset(dC, 0xff, (size_t)m * n * 4));
  int rc = launch_split_bf16(dA, m, k, k, 2, SPLIT_A, sA, nullptr);
  if (!rc) rc = launch_split_bf16(dB, k, n, n, 2, SPLIT_B, sB, nullptr);     // rows = K, [hi | mid] blocks of Np columns
  TcEpilogue ep;
  ep.mode = TC_EPI_STORE;
  ep.C = dC;
  ep.ldc = n;
  if (!rc) rc = launch_gemm_tc_pair_bn(sA, m, sB, k, Np, n, 0, n, Kp, ep, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (!rc && e != cudaSuccess) {
    set_error(std::string("gemm_tc (N-major B): ") + cudaGetErrorString(e));
    rc = -1;
  }
  if (!rc) SERT_CUDA(cudaMemcpy(c_host, dC, (size_t)m * n * 4, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(sA); cudaFree(sB);
  return rc;
}

// ---- measurement hook: raw throughput of the tcgen05 kernel on zero operands ---------------------------
// mode 0: fp32 store epilogue into ONE aliased row (ldc = 0: the C traffic stays in L2), mode 1: top-k filter
// with thresholds that reject everything (pass 1 only).  Returns the mean launch time in milliseconds.
extern "C" SERT_API int sert_debug_gemm_tc_bench(int m, int n, int kt, int reps, int mode, float *ms_out) {
  using namespace sert;
  SERT_REQUIRE(m > 0 && n > 0 && kt > 0 && kt % 64 == 0 && reps > 0 && ms_out, "bad argument");
  __nv_bfloat16 *A = nullptr, *B = nullptr;
  float *C = nullptr;
  unsigned long long *tau = nullptr, *cand = nullptr;
  int *count = nullptr, *overflow = nullptr;
  SERT_CUDA(cudaMalloc(&A, (size_t)m * kt * 2));
  SERT_CUDA(cudaMalloc(&B, (size_t)n * kt * 2));
  SERT_CUDA(cudaMalloc(&C, (size_t)n * 4 + 1024));
  SERT_CUDA(cudaMalloc(&tau, (size_t)m * 8));
  SERT_CUDA(cudaMalloc(&cand, (size_t)m * 64 * 8));
  SERT_CUDA(cudaMalloc(&count, (size_t)m * 4));
  SERT_CUDA(cudaMalloc(&overflow, 4));
  SERT_CUDA(cudaMemset(A, 0, (size_t)m * kt * 2));
  SERT_CUDA(cudaMemset(B, 0, (size_t)n * kt * 2));
  SERT_CUDA(cudaMemset(tau, 0xff, (size_t)m * 8));      // threshold decodes to NaN: no score compares greater
  SERT_CUDA(cudaMemset(count, 0, (size_t)m * 4));
  TcEpilogue ep;
  if (mode == 0) {
    ep.mode = TC_EPI_STORE; ep.C = C; ep.ldc = 0;
  } else {
    ep.mode = TC_EPI_TOPK; ep.tau = tau; ep.count = count; ep.cand = cand; ep.cap = 64; ep.overflow = overflow;
  }
  cudaEvent_t e0, e1;
  SERT_CUDA(cudaEventCreate(&e0));
  SERT_CUDA(cudaEventCreate(&e1));
  int rc = 0;
  for (int i = 0; i < 2 && !rc; ++i) rc = launch_gemm_tc(A, m, B, n, 0, n, kt, ep, nullptr);
  SERT_CUDA(cudaEventRecord(e0));
  for (int i = 0; i < reps && !rc; ++i) rc = launch_gemm_tc(A, m, B, n, 0, n, kt, ep, nullptr);
  SERT_CUDA(cudaEventRecord(e1));
  cudaError_t e = cudaEventSynchronize(e1);
  if (!rc && e != cudaSuccess) { set_error(cudaGetErrorString(e)); rc = -1; }
  float ms = 0.f;
  if (!rc) { SERT_CUDA(cudaEventElapsedTime(&ms, e0, e1)); *ms_out = ms / reps; }
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(tau); cudaFree(cand); cudaFree(count); cudaFree(overflow);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

// ---- transposing fp32 -> bf16 split: dst (C, terms * R64) <- src (R, C) ----------------------------------
namespace sert {

__global__ void __launch_bounds__(256) split_bf16_t_kernel(const float *__restrict__ src, long long R, long long C,
                                                           long long ld_src, long long Rp, int terms, int role,
                                                           __nv_bfloat16 *__restrict__ dst) {
  __shared__ float tile[32][33];
  const long long c_tiles = (C + 31) / 32;
  const long long r_tiles = Rp / 32;
  for (long long t = blockIdx.x; t < c_tiles * r_tiles; t += gridDim.x) {
    const long long r0 = (t / c_tiles) * 32, c0 = (t % c_tiles) * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
      const long long r = r0 + j, c = c0 + threadIdx.x;
      tile[j][threadIdx.x] = (r < R && c < C) ? src[r * ld_src + c] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
      const long long c = c0 + j, r = r0 + threadIdx.x;
      if (c < C) {
        const float x = tile[threadIdx.x][j];
        const __nv_bfloat16 hi = __float2bfloat16_rn(x);
        __nv_bfloat16 *d = dst + c * (terms * Rp) + r;
        if (terms == 1) {
          d[0] = hi;
        } else {
          const __nv_bfloat16 mid = __float2bfloat16_rn(x - __bfloat162float(hi));
          d[0] = hi;
          if (terms == 2) {
            d[Rp] = mid;
          } else {
            d[Rp] = role == SPLIT_A ? hi : mid;
            d[2 * Rp] = role == SPLIT_A ? mid : hi;
          }
        }
      }
    }
    __syncthreads();
  }
}

int launch_split_bf16_t(const float *src, long long R, long long C, long long ld_src, int terms, SplitRole role,
                        __nv_bfloat16 *dst, cudaStream_t st) {
  SERT_REQUIRE(terms >= 1 && terms <= 3, "split terms must be 1, 2 (pair operands) or 3");
  if (R == 0 || C == 0) return 0;
  const long long Rp = tc_padded_k((int)R);
  const long long tiles = ((C + 31) / 32) * (Rp / 32);
  const int blocks = (int)std::min<long long>(tiles, (long long)kNumSMs * 32);
  split_bf16_t_kernel<<<blocks, dim3(32, 8), 0, st>>>(src, R, C, ld_src, Rp, terms, (int)role, dst);
  SERT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sert
