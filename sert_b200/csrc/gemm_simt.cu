// Small float32 GEMMs on CUDA cores for the vector-space projection layer
// (sert/models.py:1057-1061: tanh(h.Wp + bp), (B,dw)x(dw,de)) and its two gradients.
// These are ~0.13 GFLOP each at B=4096, d=128 -- three orders of magnitude below the
// dense-update traffic of the same step -- so exact fp32 FMA is used instead of tensor cores
// (the tensor-core kernels live in gemm_tc.cu for the word x entity and query x entity GEMMs).
#include "kernels.cuh"

namespace sert {

constexpr int BM = 64, BN = 64, BK = 16;

// C[M,N] = epi(op(A) * op(B)).  A_T: A stored (K,M) row-major (lda = M-stride of k rows);
// B_T: B stored (N,K) row-major.  256 threads, each computes a 4x4 micro-tile.
template <bool A_T, bool B_T, int EPI>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float *__restrict__ A,
                                                       const float *__restrict__ Bm, float *__restrict__ C,
                                                       int M, int N, int K, int lda, int ldb, int ldc,
                                                       const float *__restrict__ bias, int k_per_split) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads -> 64 x 64 outputs
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // ---- stage A tile (BM x BK) ----
#pragma unroll
    for (int l = 0; l < (BM * BK) / 256; ++l) {
      const int e = tid + l * 256;
      int m, k;
      if (A_T) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < k_end) v = A_T ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
      As[k][m] = v;
    }
    // ---- stage B tile (BK x BN) ----
#pragma unroll
    for (int l = 0; l < (BN * BK) / 256; ++l) {
      const int e = tid + l * 256;
      int n, k;
      if (B_T) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < k_end) v = B_T ? Bm[(size_t)gn * ldb + gk] : Bm[(size_t)gk * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
      const float a_[4] = {av.x, av.y, av.z, av.w};
      const float b_[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a_[i], b_[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      float *dst = C + (size_t)gm * ldc + gn;
      if (EPI == EPI_BIAS_TANH) {
        *dst = tanhf(v + bias[gn]);
      } else if (EPI == EPI_BIAS) {
        *dst = v + bias[gn];
      } else if (EPI == EPI_ATOMIC_ADD) {
        atomicAdd(dst, v);
      } else {
        *dst = v;
      }
    }
  }
}

template <bool A_T, bool B_T>
static int dispatch_epi(const float *A, const float *Bm, float *C, int M, int N, int K, int lda, int ldb,
                        int ldc, GemmEpilogue epi, const float *bias, int split_k, cudaStream_t st) {
  const int kps = (int)align_up((size_t)cdiv(K, split_k), BK);
  dim3 grid(cdiv(N, BN), cdiv(M, BM), cdiv(K, kps));
  switch (epi) {
    case EPI_STORE:
      gemm_f32_kernel<A_T, B_T, EPI_STORE><<<grid, 256, 0, st>>>(A, Bm, C, M, N, K, lda, ldb, ldc, bias, kps);
      break;
    case EPI_BIAS_TANH:
      gemm_f32_kernel<A_T, B_T, EPI_BIAS_TANH><<<grid, 256, 0, st>>>(A, Bm, C, M, N, K, lda, ldb, ldc, bias, kps);
      break;
    case EPI_BIAS:
      gemm_f32_kernel<A_T, B_T, EPI_BIAS><<<grid, 256, 0, st>>>(A, Bm, C, M, N, K, lda, ldb, ldc, bias, kps);
      break;
    case EPI_ATOMIC_ADD:
      gemm_f32_kernel<A_T, B_T, EPI_ATOMIC_ADD><<<grid, 256, 0, st>>>(A, Bm, C, M, N, K, lda, ldb, ldc, bias, kps);
      break;
  }
  SERT_LAUNCH_CHECK();
  return 0;
}

int launch_gemm_f32(const float *A, const float *Bm, float *C, int M, int N, int K, bool a_t, bool b_t,
                    int lda, int ldb, int ldc, GemmEpilogue epi, const float *bias, int split_k,
                    cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  SERT_REQUIRE(split_k >= 1, "split_k must be positive");
  SERT_REQUIRE(split_k == 1 || epi == EPI_ATOMIC_ADD, "split-K needs the atomic epilogue");
  if (!a_t && !b_t) return dispatch_epi<false, false>(A, Bm, C, M, N, K, lda, ldb, ldc, epi, bias, split_k, st);
  if (!a_t && b_t) return dispatch_epi<false, true>(A, Bm, C, M, N, K, lda, ldb, ldc, epi, bias, split_k, st);
  if (a_t && !b_t) return dispatch_epi<true, false>(A, Bm, C, M, N, K, lda, ldb, ldc, epi, bias, split_k, st);
  return dispatch_epi<true, true>(A, Bm, C, M, N, K, lda, ldb, ldc, epi, bias, split_k, st);
}

// out[n] += sum_m A[m,n]; blockDim (32, 8): 32 columns x 8 row lanes, rows strided over gridDim.y
__global__ void colsum_atomic_kernel(const float *__restrict__ A, float *__restrict__ out, int M, int N) {
  __shared__ float part[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (n < N)
    for (int m = blockIdx.y * 8 + threadIdx.y; m < M; m += gridDim.y * 8) s += A[(size_t)m * N + n];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float tot = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) tot += part[r][threadIdx.x];
    atomicAdd(out + n, tot);
  }
}

int launch_colsum_atomic(const float *A, float *out, int M, int N, cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  dim3 grid(cdiv(N, 32), min(64, cdiv(M, 64)));
  colsum_atomic_kernel<<<grid, dim3(32, 8), 0, st>>>(A, out, M, N);
  SERT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sert
