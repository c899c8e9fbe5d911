// FP64 tensor-core GEMM used by the blocked LU (trailing update, recursive panel, TRSM):
//     C[M x N] -= A[M x K] * B[K x N]          (all row-major, arbitrary leading dims)
//
// This is the "one true dense contraction" of the Boltzmann solve (reference:
// LAPACK dgesv behind np.linalg.solve, boltzmann.py:283).  tcgen05 has no FP64 kind, so
// the FP64 tensor pipe is driven with warp-level mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4;
// ptxas splits every larger f64 shape into it on sm_100a).  Operand tiles are staged
// through shared memory with a 4-stage cp.async pipeline; accumulators are pre-loaded
// with C so the epilogue is a plain store.
#include "wg_common.cuh"
#include "wg_kernels.h"

namespace wg {

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES, bool VEC, int MINB = 1>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32, MINB)
dgemm_sub_kernel(double* __restrict__ C, const double* __restrict__ A, const double* __restrict__ B,
                 int M, int N, int K, long long ldc, long long lda, long long ldb, long long sC, long long sA,
                 long long sB) {
  // blockIdx.z = system of a strided batch (independent systems in lock step; strides 0 for one system)
  C += (long long)blockIdx.z * sC;
  A += (long long)blockIdx.z * sA;
  B += (long long)blockIdx.z * sB;
  constexpr int NT = WARPS_M * WARPS_N * 32;
  constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
  constexpr int MI = WM / 8, NJ = WN / 8;
  constexpr int LDA_S = BK + 4, LDB_S = BN + 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* As = reinterpret_cast<double*>(smem_raw);               // [STAGES][BM][LDA_S]
  double* Bs = As + (size_t)STAGES * BM * LDA_S;                  // [STAGES][BK][LDB_S]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int KT = (K + BK - 1) / BK;

  auto load_tile = [&](int kt, int st) {
    const int k0 = kt * BK;
    double* as = As + (size_t)st * BM * LDA_S;
    double* bs = Bs + (size_t)st * BK * LDB_S;
    if (VEC) {
      constexpr int ACH = BM * BK / 2;
#pragma unroll
      for (int c = tid; c < ACH; c += NT) {
        const int r = c / (BK / 2), kc = (c % (BK / 2)) * 2;
        const bool ok = (m0 + r < M) && (k0 + kc < K);
        const double* src = ok ? A + (long long)(m0 + r) * lda + k0 + kc : A;
        cp_async16(as + r * LDA_S + kc, src, ok);
      }
      constexpr int BCH = BK * BN / 2;
#pragma unroll
      for (int c = tid; c < BCH; c += NT) {
        const int r = c / (BN / 2), nc = (c % (BN / 2)) * 2;
        const bool ok = (k0 + r < K) && (n0 + nc < N);
        const double* src = ok ? B + (long long)(k0 + r) * ldb + n0 + nc : B;
        cp_async16(bs + r * LDB_S + nc, src, ok);
      }
    } else {
      for (int c = tid; c < BM * BK; c += NT) {
        const int r = c / BK, kc = c % BK;
        const bool ok = (m0 + r < M) && (k0 + kc < K);
        const double* src = ok ? A + (long long)(m0 + r) * lda + k0 + kc : A;
        cp_async8(as + r * LDA_S + kc, src, ok);
      }
      for (int c = tid; c < BK * BN; c += NT) {
        const int r = c / BN, nc = c % BN;
        const bool ok = (k0 + r < K) && (n0 + nc < N);
        const double* src = ok ? B + (long long)(k0 + r) * ldb + n0 + nc : B;
        cp_async8(bs + r * LDB_S + nc, src, ok);
      }
    }
  };

  // start the operand pipeline first, then fetch C while the copies are in flight
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_tile(s, s);
    cp_async_commit();
  }

  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int r = m0 + wm * WM + i * 8 + g;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = n0 + wn * WN + j * 8 + 2 * t;
      double c0 = 0.0, c1 = 0.0;
      if (r < M) {
        const double* p = C + (long long)r * ldc + c;
        if (VEC) {
          if (c < N) {  // N even on the VEC path
            const double2 v = *reinterpret_cast<const double2*>(p);
            c0 = v.x;
            c1 = v.y;
          }
        } else {
          if (c < N) c0 = p[0];
          if (c + 1 < N) c1 = p[1];
        }
      }
      acc[i][j][0] = c0;
      acc[i][j][1] = c1;
    }
  }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      if (nk < KT) load_tile(nk, nk % STAGES);
      cp_async_commit();
    }
    const double* as = As + (size_t)(kt % STAGES) * BM * LDA_S + (wm * WM + g) * LDA_S + t;
    const double* bs = Bs + (size_t)(kt % STAGES) * BK * LDB_S + t * LDB_S + wn * WN + g;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double af[MI], bf[NJ];
#pragma unroll
      for (int i = 0; i < MI; ++i) af[i] = -as[i * 8 * LDA_S + kk * 4];
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[j] = bs[kk * 4 * LDB_S + j * 8];
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int r = m0 + wm * WM + i * 8 + g;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = n0 + wn * WN + j * 8 + 2 * t;
      double* p = C + (long long)r * ldc + c;
      if (VEC) {
        if (c < N) *reinterpret_cast<double2*>(p) = make_double2(acc[i][j][0], acc[i][j][1]);
      } else {
        if (c < N) p[0] = acc[i][j][0];
        if (c + 1 < N) p[1] = acc[i][j][1];
      }
    }
  }
}

template <int BM, int BN, int BK, int WARPS_M, int WARPS_N, int STAGES, int MINB = 1>
static int launch_cfg(double* C, const double* A, const double* B, int M, int N, int K, long long ldc,
                      long long lda, long long ldb, bool vec, cudaStream_t st, const GemmBatch& gb) {
  constexpr int NT = WARPS_M * WARPS_N * 32;
  constexpr size_t SMEM = (size_t)STAGES * (BM * (BK + 4) + BK * (BN + 4)) * sizeof(double);
  dim3 grid(wg_cdiv(N, BN), wg_cdiv(M, BM), gb.count);
  if (vec) {
    auto kern = dgemm_sub_kernel<BM, BN, BK, WARPS_M, WARPS_N, STAGES, true, MINB>;
    if (wg_func_configured(reinterpret_cast<const void*>(kern), SMEM) < (long long)SMEM)
      WG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    WG_CUDA(launch_prio(kern, grid, dim3(NT), SMEM, st, C, A, B, M, N, K, ldc, lda, ldb, gb.sC, gb.sA, gb.sB));
  } else {
    auto kern = dgemm_sub_kernel<BM, BN, BK, WARPS_M, WARPS_N, STAGES, false, MINB>;
    if (wg_func_configured(reinterpret_cast<const void*>(kern), SMEM) < (long long)SMEM)
      WG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    WG_CUDA(launch_prio(kern, grid, dim3(NT), SMEM, st, C, A, B, M, N, K, ldc, lda, ldb, gb.sC, gb.sA, gb.sB));
  }
  WG_LAUNCH_CHECK();
  return WG_OK;
}

static int g_gemm_variant = 1;  // 1: 128x64 tiles, two CTAs per SM (default); 0: 128x128, one CTA per SM
void gemm_set_variant(int v) { g_gemm_variant = v; }

int dgemm_sub(double* C, const double* A, const double* B, int M, int N, int K, long long ldc,
              long long lda, long long ldb, cudaStream_t st, GemmBatch gb) {
  if (M <= 0 || N <= 0 || K <= 0 || gb.count <= 0) return WG_OK;
  const bool vec = ((ldc | lda | ldb | gb.sC | gb.sA | gb.sB) % 2 == 0) && (N % 2 == 0) && (K % 2 == 0) &&
                   ((reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(A) |
                     reinterpret_cast<uintptr_t>(B)) % 16 == 0);
  // every variant accumulates k in the same order, so the choice of tile never changes a result bit
  if (N > 64 && g_gemm_variant == 1) return launch_cfg<128, 64, 16, 2, 2, 3, 2>(C, A, B, M, N, K, ldc, lda, ldb, vec, st, gb);
  if (N > 64 && g_gemm_variant == 2) return launch_cfg<128, 128, 32, 2, 4, 3>(C, A, B, M, N, K, ldc, lda, ldb, vec, st, gb);
  if (N > 64) return launch_cfg<128, 128, 16, 2, 4, 4>(C, A, B, M, N, K, ldc, lda, ldb, vec, st, gb);
  if (N > 32) return launch_cfg<128, 64, 16, 4, 2, 4>(C, A, B, M, N, K, ldc, lda, ldb, vec, st, gb);
  return launch_cfg<128, 32, 16, 8, 1, 4>(C, A, B, M, N, K, ldc, lda, ldb, vec, st, gb);
}

}  // namespace wg
