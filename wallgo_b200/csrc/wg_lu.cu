// Blocked right-looking FP64 LU with row partial pivoting, row-major storage, and the
// triangular solves.  Replaces LAPACK dgesv behind np.linalg.solve
// (/root/reference/src/WallGo/boltzmann.py:283, :548).  Same pivot rule as dgetf2
// (first row of maximal |a| in the column; reciprocal scaling), so P*A = L*U holds with
// LAPACK-style ipiv (0-based here).
//
// Structure of one top-level panel [k, k+NB):
//   * recursive column-halving inside the panel; every flop that is not a leaf goes through
//     the DMMA GEMM (wg_gemm.cu);
//   * leaves (16 or 8 columns) are factored by ONE thread-block cluster that keeps the whole
//     tall-skinny block register-resident (one or more rows per thread), finds pivots with
//     warp redux + a DSMEM exchange and one hardware cluster barrier per column, and never
//     moves rows while factoring (implicit pivoting); at the end it writes rows to their
//     pivoted positions and moves the rest of those rows inside the panel;
//   * the permutation of a whole panel is composed once on the device and applied to the
//     columns left and right of the panel by two fully parallel kernels (gather / write);
//   * U12 = L11^-1 A12 by recursive TRSM (GEMM + register-resident 32-wide base);
//   * trailing update A22 -= L21 U12 on the FP64 tensor pipe.
// No host synchronisation anywhere: pivots, permutations and `info` stay on the device.
#include <limits.h>
#include <stdarg.h>

#include <vector>

#include "wg_common.cuh"
#include "wg_kernels.h"

namespace wg {

constexpr int PT = 512;       // threads per leaf-panel CTA
constexpr int PNW = PT / 32;  // warps per leaf-panel CTA
constexpr int MAXCL = 16;     // largest cluster
constexpr int MAXEXT = 248;   // widest "rest of the panel" a leaf has to move (NB - 8)
constexpr int MAXMOVE = 32;   // most rows a leaf can move per CTA (2 * 16)

static int g_nb = 256;
static int g_use_graph = 0;
void lu_set_tuning(int nb, int use_graph) {
  if (nb >= 32 && nb <= 256 && nb % 16 == 0) g_nb = nb;
  g_use_graph = use_graph;
}

struct PanelArgs {
  double* A;
  long long lda;
  int n;
  int c0, w;      // leaf columns [c0, c0+w)
  int pc0, pc1;   // enclosing top-level panel columns
  int* ipiv;
  int* info;
  int* grpSrcTop;  // [n] : row (position at leaf start) that ends at position c0+i
  int* grpBDst;    // [n] : final position of the row that started at c0+i, or -1 if it stayed in the top w
  int vec;
};

__device__ __forceinline__ void argmax_redux(long long key, int pos, long long& mkey, int& mpos) {
  const unsigned full = 0xffffffffu;
  const int hi = (int)(key >> 32);
  const unsigned lo = (unsigned)key;
  const int mhi = __reduce_max_sync(full, hi);
  const unsigned mlo = __reduce_max_sync(full, hi == mhi ? lo : 0u);
  const bool cand = (hi == mhi) && (lo == mlo);
  mpos = __reduce_min_sync(full, cand ? pos : INT_MAX);
  mkey = ((long long)mhi << 32) | (long long)mlo;
}

template <int W, int R>
__global__ void __launch_bounds__(PT, 1) lu_panel_leaf_kernel(const PanelArgs p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int SL = W + 2;
  double* warpSlot = reinterpret_cast<double*>(smraw);     // [PNW][SL]
  double* cluSlot = warpSlot + PNW * SL;                   // [2][MAXCL][SL]
  double* moveBuf = cluSlot + 2 * MAXCL * SL;              // [MAXMOVE][MAXEXT]
  int* moveDst = reinterpret_cast<int*>(moveBuf + MAXMOVE * MAXEXT);  // [MAXMOVE]
  int* moveCount = moveDst + MAXMOVE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned rank = cluster_ctarank(), CL = cluster_nctarank();
  const int c0 = p.c0, w = p.w;
  const int m = p.n - c0;
  const long long lda = p.lda;
  double* __restrict__ A = p.A;

  double a[R][W];
  int pos[R], spos[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int lr = k * (int)(CL * PT) + (int)rank * PT + tid;
    const bool valid = lr < m;
    pos[k] = valid ? c0 + lr : -1;
    spos[k] = pos[k];
    const double* src = A + (long long)(c0 + lr) * lda + c0;
    if (p.vec) {
#pragma unroll
      for (int c = 0; c < W; c += 2) {
        double2 v = make_double2(0.0, 0.0);
        if (valid && c < w) v = *reinterpret_cast<const double2*>(src + c);
        a[k][c] = v.x;
        a[k][c + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < W; ++c) a[k][c] = (valid && c < w) ? src[c] : 0.0;
    }
  }
  if (tid == 0) *moveCount = 0;

#pragma unroll
  for (int j = 0; j < W; ++j) {
    if (j < w) {  // uniform
      const int dj = c0 + j;
      const int par = j & 1;
      // ---- thread-local and warp-level candidate
      long long bk = -1;
      int bp = INT_MAX;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const bool act = pos[k] >= dj;
        const long long key = act ? __double_as_longlong(fabs(a[k][j])) : -1LL;
        const int pp = act ? pos[k] : INT_MAX;
        if (key > bk || (key == bk && pp < bp)) {
          bk = key;
          bp = pp;
        }
      }
      long long wkey;
      int wpos;
      argmax_redux(bk, bp, wkey, wpos);
      {
        const bool owner = (wpos == INT_MAX) ? (lane == 0) : (bk == wkey && bp == wpos);
        if (owner) {
          double* slot = warpSlot + warp * SL;
#pragma unroll
          for (int k = 0; k < R; ++k)
            if (pos[k] == wpos) {
#pragma unroll
              for (int c = 0; c < W; ++c) slot[c] = a[k][c];
            }
          reinterpret_cast<long long*>(slot)[W] = wkey;
          reinterpret_cast<long long*>(slot)[W + 1] = (long long)wpos;
        }
      }
      __syncthreads();
      // ---- CTA-level winner (every warp recomputes it), pushed to every CTA of the cluster
      {
        const double* s = warpSlot + (lane & (PNW - 1)) * SL;
        const long long k2 = reinterpret_cast<const long long*>(s)[W];
        const int p2 = (int)reinterpret_cast<const long long*>(s)[W + 1];
        long long ckey;
        int cpos;
        argmax_redux(k2, p2, ckey, cpos);
        const unsigned bal = __ballot_sync(0xffffffffu, k2 == ckey && p2 == cpos);
        const int ww = __ffs(bal) - 1;
        const unsigned long long* srcw = reinterpret_cast<const unsigned long long*>(warpSlot + ww * SL);
        const uint32_t base = smem_u32(cluSlot + (par * MAXCL + (int)rank) * SL);
        for (int u = tid; u < SL * (int)CL; u += PT) {
          const int dest = u / SL, e = u - dest * SL;
          dsmem_st_u64(dsmem_map(base + 8u * e, dest), srcw[e]);
        }
      }
      cluster_sync_all();
      // ---- cluster-level winner = pivot row
      long long mkey;
      int mpos;
      int wc;
      {
        const double* s = cluSlot + (par * MAXCL + (lane & ((int)CL - 1))) * SL;
        const long long k3 = reinterpret_cast<const long long*>(s)[W];
        const int p3 = (int)reinterpret_cast<const long long*>(s)[W + 1];
        argmax_redux(k3, p3, mkey, mpos);
        const unsigned bal = __ballot_sync(0xffffffffu, k3 == mkey && p3 == mpos);
        wc = __ffs(bal) - 1;
      }
      const double* urow = cluSlot + (par * MAXCL + wc) * SL;
      const bool sing = (mkey <= 0);
      const double rinv = sing ? 0.0 : 1.0 / urow[j];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (pos[k] == mpos)
          pos[k] = dj;
        else if (pos[k] == dj)
          pos[k] = mpos;
        if (pos[k] > dj && !sing) {
          const double l = a[k][j] * rinv;
          a[k][j] = l;
#pragma unroll
          for (int c = j + 1; c < W; ++c) a[k][c] = fma(-l, urow[c], a[k][c]);
        }
      }
      if (rank == 0 && tid == 0) {
        p.ipiv[dj] = mpos;
        if (sing && *p.info == 0) *p.info = dj + 1;
      }
    }
  }

  // ---- rows go to their pivoted positions (these w columns live in registers)
#pragma unroll
  for (int k = 0; k < R; ++k) {
    if (pos[k] < 0) continue;
    double* dst = A + (long long)pos[k] * lda + c0;
    if (p.vec) {
#pragma unroll
      for (int c = 0; c < W; c += 2)
        if (c < w) *reinterpret_cast<double2*>(dst + c) = make_double2(a[k][c], a[k][c + 1]);
    } else {
#pragma unroll
      for (int c = 0; c < W; ++c)
        if (c < w) dst[c] = a[k][c];
    }
    if (pos[k] < c0 + w) p.grpSrcTop[pos[k]] = spos[k];
    if (spos[k] < c0 + w) p.grpBDst[spos[k]] = (pos[k] >= c0 + w) ? pos[k] : -1;
  }

  // ---- move the rest of the moved rows inside the enclosing panel (read all, barrier, write all)
  const int nLeft = c0 - p.pc0;
  const int ext = nLeft + (p.pc1 - (c0 + w));
  if (ext > 0) {  // uniform
    __syncthreads();  // moveCount = 0 visible
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const bool moved = pos[k] >= 0 && pos[k] != spos[k];
      unsigned mask = __ballot_sync(0xffffffffu, moved);
      while (mask) {
        const int L = __ffs(mask) - 1;
        mask &= mask - 1;
        const int sp = __shfl_sync(0xffffffffu, spos[k], L);
        const int dp = __shfl_sync(0xffffffffu, pos[k], L);
        int slot = 0;
        if (lane == 0) slot = atomicAdd(moveCount, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        const double* srow = A + (long long)sp * lda;
        for (int e = lane; e < ext; e += 32) {
          const int col = e < nLeft ? p.pc0 + e : c0 + w + (e - nLeft);
          moveBuf[slot * MAXEXT + e] = srow[col];
        }
        if (lane == 0) moveDst[slot] = dp;
      }
    }
    cluster_sync_all();
    const int cnt = *moveCount;
    for (int s = warp; s < cnt; s += PNW) {
      double* drow = A + (long long)moveDst[s] * lda;
      for (int e = lane; e < ext; e += 32) {
        const int col = e < nLeft ? p.pc0 + e : c0 + w + (e - nLeft);
        drow[col] = moveBuf[s * MAXEXT + e];
      }
    }
  }
}

template <int W, int R>
static size_t leaf_smem() {
  return (size_t)(PNW * (W + 2) + 2 * MAXCL * (W + 2) + MAXMOVE * MAXEXT) * sizeof(double) +
         (MAXMOVE + 4) * sizeof(int);
}

template <int W, int R>
static int launch_leaf(const PanelArgs& pa, int CL, cudaStream_t st) {
  auto kern = lu_panel_leaf_kernel<W, R>;
  const size_t smem = leaf_smem<W, R>();
  static bool configured = false;
  if (!configured) {
    WG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, 1, 1);
  cfg.blockDim = dim3(PT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  WG_CUDA(cudaLaunchKernelEx(&cfg, kern, pa));
  ++g_wg_launches;
  return WG_OK;
}

// ------------------------------------------------------------------ panel permutation: compose
// Composes the leaf permutations of panel [k, k+nb) into
//   srcTop[k+i]  : row (position before the panel) whose content ends at position k+i
//   bDst/bSrc[k+t], t < nbot : position bDst receives the content that was at row bSrc (a top row)
__global__ void lu_perm_compose_kernel(int n, int k, int nb, int wb, const int* __restrict__ grpSrcTop,
                                       const int* __restrict__ grpBDst, int* __restrict__ posContent,
                                       int* __restrict__ srcTop, int* __restrict__ bDst,
                                       int* __restrict__ bSrc, int* __restrict__ nbotOut) {
  __shared__ int topC[256];
  __shared__ int touched[256];
  __shared__ int ntouched;
  const int tid = threadIdx.x;
  if (tid < nb) topC[tid] = k + tid;
  if (tid == 0) ntouched = 0;
  __syncthreads();
  for (int gs = k; gs < k + nb; gs += wb) {
    const int gw = min(wb, k + nb - gs);
    int vnew = 0, vdisp = 0, d = -1;
    if (tid < gw) {
      const int s = grpSrcTop[gs + tid];
      if (s < k + nb) {
        vnew = topC[s - k];
      } else {
        const int pc = posContent[s];
        vnew = pc >= 0 ? pc : s;
      }
      vdisp = topC[gs + tid - k];
      d = grpBDst[gs + tid];
    }
    __syncthreads();
    if (tid < gw) {
      topC[gs + tid - k] = vnew;
      if (d >= 0) {
        if (d < k + nb) {
          topC[d - k] = vdisp;
        } else {
          if (posContent[d] < 0) touched[atomicAdd(&ntouched, 1)] = d;
          posContent[d] = vdisp;
        }
      }
    }
    __syncthreads();
  }
  if (tid < nb) srcTop[k + tid] = topC[tid];
  const int nt = ntouched;
  if (tid < nt) {
    const int d = touched[tid];
    bDst[k + tid] = d;
    bSrc[k + tid] = posContent[d];
    posContent[d] = -1;  // leave the table clean for the next panel
  }
  if (tid == 0) *nbotOut = nt;
}

// ------------------------------------------------------------------ panel permutation: apply
// columns outside [k, k+nb): c' in [0, n-nb) -> col = c' < k ? c' : c' + nb
__global__ void lu_perm_gather_kernel(const double* __restrict__ A, long long lda, int n, int k, int nb,
                                      const int* __restrict__ srcTop, const int* __restrict__ bSrc,
                                      const int* __restrict__ nbotPtr, double* __restrict__ S,
                                      double* __restrict__ S2, int vec) {
  const int i = blockIdx.y;
  const int ncol = n - nb;
  const int nbot = *nbotPtr;
  const int src1 = srcTop[k + i];
  const bool do1 = src1 != k + i;
  const bool do2 = i < nbot;
  if (!do1 && !do2) return;
  const int src2 = do2 ? bSrc[k + i] : 0;
  const double* r1 = A + (long long)src1 * lda;
  const double* r2 = A + (long long)src2 * lda;
  double* s1 = S + (long long)i * ncol;
  double* s2 = S2 + (long long)i * ncol;
  if (vec) {
    for (int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x); c < ncol; c += 2 * gridDim.x * blockDim.x) {
      const int col = c < k ? c : c + nb;
      if (do1) *reinterpret_cast<double2*>(s1 + c) = *reinterpret_cast<const double2*>(r1 + col);
      if (do2) *reinterpret_cast<double2*>(s2 + c) = *reinterpret_cast<const double2*>(r2 + col);
    }
  } else {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncol; c += gridDim.x * blockDim.x) {
      const int col = c < k ? c : c + nb;
      if (do1) s1[c] = r1[col];
      if (do2) s2[c] = r2[col];
    }
  }
}

__global__ void lu_perm_write_kernel(double* __restrict__ A, long long lda, int n, int k, int nb,
                                     const int* __restrict__ srcTop, const int* __restrict__ bDst,
                                     const int* __restrict__ nbotPtr, const double* __restrict__ S,
                                     const double* __restrict__ S2, int vec) {
  const int i = blockIdx.y;
  const int ncol = n - nb;
  const int nbot = *nbotPtr;
  const bool do1 = srcTop[k + i] != k + i;
  const bool do2 = i < nbot;
  if (!do1 && !do2) return;
  const int dst2 = do2 ? bDst[k + i] : 0;
  double* r1 = A + (long long)(k + i) * lda;
  double* r2 = A + (long long)dst2 * lda;
  const double* s1 = S + (long long)i * ncol;
  const double* s2 = S2 + (long long)i * ncol;
  if (vec) {
    for (int c = 2 * (blockIdx.x * blockDim.x + threadIdx.x); c < ncol; c += 2 * gridDim.x * blockDim.x) {
      const int col = c < k ? c : c + nb;
      if (do1) *reinterpret_cast<double2*>(r1 + col) = *reinterpret_cast<const double2*>(s1 + c);
      if (do2) *reinterpret_cast<double2*>(r2 + col) = *reinterpret_cast<const double2*>(s2 + c);
    }
  } else {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncol; c += gridDim.x * blockDim.x) {
      const int col = c < k ? c : c + nb;
      if (do1) r1[col] = s1[c];
      if (do2) r2[col] = s2[c];
    }
  }
}

// ------------------------------------------------------------------ TRSM base (unit lower, in place)
template <int W>
__global__ void __launch_bounds__(128) lu_trsm_base_kernel(const double* __restrict__ L, long long ldl,
                                                          double* __restrict__ B, long long ldb, int w,
                                                          int ncols) {
  __shared__ double Ls[W][W + 1];
  for (int e = threadIdx.x; e < W * W; e += blockDim.x) {
    const int r = e / W, c = e % W;
    Ls[r][c] = (r < w && c < r) ? L[(long long)r * ldl + c] : 0.0;
  }
  __syncthreads();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  double x[W];
#pragma unroll
  for (int i = 0; i < W; ++i) x[i] = i < w ? B[(long long)i * ldb + col] : 0.0;
#pragma unroll
  for (int i = 1; i < W; ++i) {
    double s = x[i];
#pragma unroll
    for (int kk = 0; kk < i; ++kk) s = fma(-Ls[i][kk], x[kk], s);
    x[i] = s;
  }
#pragma unroll
  for (int i = 0; i < W; ++i)
    if (i < w) B[(long long)i * ldb + col] = x[i];
}

static int trsm_rec(const double* L, long long ldl, double* B, long long ldb, int w, int ncols,
                    cudaStream_t st) {
  if (w <= 0 || ncols <= 0) return WG_OK;
  if (w <= 32) {
    const int grid = wg_cdiv(ncols, 128);
    if (w <= 8)
      lu_trsm_base_kernel<8><<<grid, 128, 0, st>>>(L, ldl, B, ldb, w, ncols);
    else if (w <= 16)
      lu_trsm_base_kernel<16><<<grid, 128, 0, st>>>(L, ldl, B, ldb, w, ncols);
    else
      lu_trsm_base_kernel<32><<<grid, 128, 0, st>>>(L, ldl, B, ldb, w, ncols);
    WG_LAUNCH_CHECK();
    return WG_OK;
  }
  const int w1 = 32 * (((w + 31) / 32 + 1) / 2);
  int rc = trsm_rec(L, ldl, B, ldb, w1, ncols, st);
  if (rc) return rc;
  rc = dgemm_sub(B + (long long)w1 * ldb, L + (long long)w1 * ldl, B, w - w1, ncols, w1, ldb, ldl, ldb, st);
  if (rc) return rc;
  return trsm_rec(L + (long long)w1 * ldl + w1, ldl, B + (long long)w1 * ldb, ldb, w - w1, ncols, st);
}

// ------------------------------------------------------------------ solve kernels
__global__ void lu_rhs_perm_kernel(double* __restrict__ b, const int* __restrict__ srcTop,
                                   const int* __restrict__ bDst, const int* __restrict__ bSrc,
                                   const int* __restrict__ nbot, int n, int NB) {
  const int tid = threadIdx.x;
  for (int k = 0, pnl = 0; k < n; k += NB, ++pnl) {
    const int nb = min(NB, n - k);
    const int nbt = nbot[pnl];
    double v = 0.0;
    int dst = -1;
    if (tid < nb) {
      v = b[srcTop[k + tid]];
      dst = k + tid;
    } else if (tid - nb < nbt) {
      v = b[bSrc[k + tid - nb]];
      dst = bDst[k + tid - nb];
    }
    __syncthreads();
    if (dst >= 0) b[dst] = v;
    __syncthreads();
  }
}

// y[r0+i] -= sum_{c in [c0,c1)} Mx[r0+i][c] * x[c]     (one warp per row)
__global__ void lu_gemv_sub_kernel(const double* __restrict__ Mx, long long lda, const double* x, double* y,
                                   int r0, int nr, int c0, int c1) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nr) return;
  const double* row = Mx + (long long)(r0 + warp) * lda;
  double s = 0.0;
  for (int c = c0 + lane; c < c1; c += 32) s = fma(row[c], x[c], s);
  s = warp_sum(s);
  if (lane == 0) y[r0 + warp] -= s;
}

__global__ void __launch_bounds__(256) lu_trsv_lower_block_kernel(const double* __restrict__ LU, long long lda,
                                                                   double* __restrict__ b, int kb, int bs) {
  __shared__ double y[256];
  __shared__ double Ts[32][33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < bs) y[tid] = b[kb + tid];
  __syncthreads();
  for (int sb = 0; sb < bs; sb += 32) {
    const int sw = min(32, bs - sb);
    for (int e = tid; e < 32 * 32; e += 256) {
      const int r = e >> 5, c = e & 31;
      Ts[r][c] = (r < sw && c < r) ? LU[(long long)(kb + sb + r) * lda + kb + sb + c] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double yi = lane < sw ? y[sb + lane] : 0.0;
      for (int r = 0; r < sw; ++r) {
        const double yr = __shfl_sync(0xffffffffu, yi, r);
        if (lane > r) yi = fma(-Ts[lane][r], yr, yi);
      }
      if (lane < sw) y[sb + lane] = yi;
    }
    __syncthreads();
    for (int i = sb + sw + tid; i < bs; i += 256) {
      const double* row = LU + (long long)(kb + i) * lda + kb + sb;
      double s = y[i];
      for (int c = 0; c < sw; ++c) s = fma(-row[c], y[sb + c], s);
      y[i] = s;
    }
    __syncthreads();
  }
  if (tid < bs) b[kb + tid] = y[tid];
}

__global__ void __launch_bounds__(256) lu_trsv_upper_block_kernel(const double* __restrict__ LU, long long lda,
                                                                   double* __restrict__ b, int kb, int bs) {
  __shared__ double y[256];
  __shared__ double Ts[32][33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < bs) y[tid] = b[kb + tid];
  __syncthreads();
  const int nsb = (bs + 31) / 32;
  for (int q = nsb - 1; q >= 0; --q) {
    const int sb = q * 32;
    const int sw = min(32, bs - sb);
    for (int e = tid; e < 32 * 32; e += 256) {
      const int r = e >> 5, c = e & 31;
      Ts[r][c] = (r < sw && c < sw && c >= r) ? LU[(long long)(kb + sb + r) * lda + kb + sb + c] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double yi = lane < sw ? y[sb + lane] : 0.0;
      for (int r = sw - 1; r >= 0; --r) {
        if (lane == r) yi = yi / Ts[r][r];
        const double yr = __shfl_sync(0xffffffffu, yi, r);
        if (lane < r) yi = fma(-Ts[lane][r], yr, yi);
      }
      if (lane < sw) y[sb + lane] = yi;
    }
    __syncthreads();
    for (int i = tid; i < sb; i += 256) {
      const double* row = LU + (long long)(kb + i) * lda + kb + sb;
      double s = y[i];
      for (int c = 0; c < sw; ++c) s = fma(-row[c], y[sb + c], s);
      y[i] = s;
    }
    __syncthreads();
  }
  if (tid < bs) b[kb + tid] = y[tid];
}

// ------------------------------------------------------------------ host-side plan / scheduler
struct LuPlan {
  int n = 0;
  int nb = 256;
  int maxCL = 16;
  int *grpSrcTop = nullptr, *grpBDst = nullptr;
  int *permSrcTop = nullptr, *permBDst = nullptr, *permBSrc = nullptr, *permNBot = nullptr;
  int* posContent = nullptr;
  double *S = nullptr, *S2 = nullptr;
  // CUDA-graph cache of the factorisation for one set of buffers
  cudaGraphExec_t exec = nullptr;
  long long graphLaunches = 0;
  double* gA = nullptr;
  long long glda = 0;
  int* gipiv = nullptr;
  int* ginfo = nullptr;
};

template <int W, int R>
static int probe_max_cluster() {
  auto kern = lu_panel_leaf_kernel<W, R>;
  const size_t smem = leaf_smem<W, R>();
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return 8;
  for (int cl = 16; cl >= 2; cl >>= 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl, 1, 1);
    cfg.blockDim = dim3(PT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) == cudaSuccess && nclusters > 0) return cl;
  }
  (void)cudaGetLastError();
  return 1;
}

int lu_plan_create(int n, LuPlan** out) {
  WG_REQUIRE(n > 0 && out != nullptr, "lu_plan_create: bad arguments");
  LuPlan* p = new LuPlan();
  p->n = n;
  p->nb = g_nb;
  p->maxCL = probe_max_cluster<16, 2>();
  const int npanels = wg_cdiv(n, 32) + 1;
  WG_CUDA(cudaMalloc(&p->grpSrcTop, sizeof(int) * n));
  WG_CUDA(cudaMalloc(&p->grpBDst, sizeof(int) * n));
  WG_CUDA(cudaMalloc(&p->permSrcTop, sizeof(int) * n));
  WG_CUDA(cudaMalloc(&p->permBDst, sizeof(int) * n));
  WG_CUDA(cudaMalloc(&p->permBSrc, sizeof(int) * n));
  WG_CUDA(cudaMalloc(&p->permNBot, sizeof(int) * npanels));
  WG_CUDA(cudaMalloc(&p->posContent, sizeof(int) * n));
  WG_CUDA(cudaMemset(p->posContent, 0xff, sizeof(int) * n));
  WG_CUDA(cudaMemset(p->permNBot, 0, sizeof(int) * npanels));
  const size_t sbytes = sizeof(double) * (size_t)256 * (size_t)n;
  WG_CUDA(cudaMalloc(&p->S, sbytes));
  WG_CUDA(cudaMalloc(&p->S2, sbytes));
  *out = p;
  return WG_OK;
}

void lu_plan_destroy(LuPlan* p) {
  if (!p) return;
  if (p->exec) cudaGraphExecDestroy(p->exec);
  cudaFree(p->grpSrcTop);
  cudaFree(p->grpBDst);
  cudaFree(p->permSrcTop);
  cudaFree(p->permBDst);
  cudaFree(p->permBSrc);
  cudaFree(p->permNBot);
  cudaFree(p->posContent);
  cudaFree(p->S);
  cudaFree(p->S2);
  delete p;
}

struct FactorCtx {
  LuPlan* p;
  double* A;
  long long lda;
  int* ipiv;
  int* info;
  cudaStream_t st;
  int wb;  // leaf width for the current panel
  int vecA;
};

static int leaf(FactorCtx& c, int c0, int w, int pc0, int pc1) {
  const int m = c.p->n - c0;
  PanelArgs pa;
  pa.A = c.A;
  pa.lda = c.lda;
  pa.n = c.p->n;
  pa.c0 = c0;
  pa.w = w;
  pa.pc0 = pc0;
  pa.pc1 = pc1;
  pa.ipiv = c.ipiv;
  pa.info = c.info;
  pa.grpSrcTop = c.p->grpSrcTop;
  pa.grpBDst = c.p->grpBDst;
  pa.vec = c.vecA && (c0 % 2 == 0) && (w % 2 == 0);
  auto pick_cl = [&](int rows_per_cta) {
    int cl = 1;
    while (cl * rows_per_cta < m) cl <<= 1;
    return cl;
  };
  if (c.wb == 16) {
    if (m <= PT) return launch_leaf<16, 1>(pa, 1, c.st);
    return launch_leaf<16, 2>(pa, pick_cl(PT * 2), c.st);
  }
  if (m <= PT) return launch_leaf<8, 1>(pa, 1, c.st);
  if (m <= c.p->maxCL * PT * 2) return launch_leaf<8, 2>(pa, pick_cl(PT * 2), c.st);
  return launch_leaf<8, 4>(pa, pick_cl(PT * 4), c.st);
}

static int panel_rec(FactorCtx& c, int c0, int w, int pc0, int pc1) {
  if (w <= c.wb) return leaf(c, c0, w, pc0, pc1);
  const int units = wg_cdiv(w, c.wb);
  const int w1 = c.wb * ((units + 1) / 2);
  const int n = c.p->n;
  int rc = panel_rec(c, c0, w1, pc0, pc1);
  if (rc) return rc;
  double* A = c.A;
  const long long lda = c.lda;
  rc = trsm_rec(A + (long long)c0 * lda + c0, lda, A + (long long)c0 * lda + c0 + w1, lda, w1, w - w1, c.st);
  if (rc) return rc;
  rc = dgemm_sub(A + (long long)(c0 + w1) * lda + c0 + w1, A + (long long)(c0 + w1) * lda + c0,
                 A + (long long)c0 * lda + c0 + w1, n - c0 - w1, w - w1, w1, lda, lda, lda, c.st);
  if (rc) return rc;
  return panel_rec(c, c0 + w1, w - w1, pc0, pc1);
}

static int factor_enqueue(LuPlan* p, double* A, long long lda, int* ipiv, int* info, cudaStream_t st) {
  const int n = p->n, NB = p->nb;
  FactorCtx c{p, A, lda, ipiv, info, st, 16, 0};
  c.vecA = (lda % 2 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
  WG_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
  for (int k = 0, pnl = 0; k < n; k += NB, ++pnl) {
    const int nb = NB < n - k ? NB : n - k;
    const int m = n - k;
    if (m <= p->maxCL * PT * 2)
      c.wb = 16;
    else if (m <= p->maxCL * PT * 4)
      c.wb = 8;
    else {
      wg_set_error("lu_factor: n=%d exceeds the leaf-panel capacity (%d rows) of this build", n,
                   p->maxCL * PT * 4);
      return WG_ERR_UNSUPPORTED;
    }
    int rc = panel_rec(c, k, nb, k, k + nb);
    if (rc) return rc;
    lu_perm_compose_kernel<<<1, 256, 0, st>>>(n, k, nb, c.wb, p->grpSrcTop, p->grpBDst, p->posContent,
                                              p->permSrcTop, p->permBDst, p->permBSrc, p->permNBot + pnl);
    WG_LAUNCH_CHECK();
    if (n - nb > 0) {
      const int vec = c.vecA && (k % 2 == 0) && (nb % 2 == 0) && (n % 2 == 0);
      dim3 grid(wg_cdiv(n - nb, 2048), nb);
      lu_perm_gather_kernel<<<grid, 256, 0, st>>>(A, lda, n, k, nb, p->permSrcTop, p->permBSrc,
                                                  p->permNBot + pnl, p->S, p->S2, vec);
      WG_LAUNCH_CHECK();
      lu_perm_write_kernel<<<grid, 256, 0, st>>>(A, lda, n, k, nb, p->permSrcTop, p->permBDst,
                                                 p->permNBot + pnl, p->S, p->S2, vec);
      WG_LAUNCH_CHECK();
    }
    const int rest = n - k - nb;
    if (rest > 0) {
      rc = trsm_rec(A + (long long)k * lda + k, lda, A + (long long)k * lda + k + nb, lda, nb, rest, st);
      if (rc) return rc;
      rc = dgemm_sub(A + (long long)(k + nb) * lda + k + nb, A + (long long)(k + nb) * lda + k,
                     A + (long long)k * lda + k + nb, rest, rest, nb, lda, lda, lda, st);
      if (rc) return rc;
    }
  }
  return WG_OK;
}

int lu_factor(LuPlan* p, double* A, long long lda, int* ipiv, int* info, cudaStream_t st) {
  WG_REQUIRE(p && A && ipiv && info && lda >= p->n, "lu_factor: bad arguments");
  if (!g_use_graph) return factor_enqueue(p, A, lda, ipiv, info, st);
  if (!p->exec || p->gA != A || p->glda != lda || p->gipiv != ipiv || p->ginfo != info) {
    if (p->exec) {
      cudaGraphExecDestroy(p->exec);
      p->exec = nullptr;
    }
    cudaStream_t cs;
    WG_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    WG_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    const long long before = g_wg_launches;
    int rc = factor_enqueue(p, A, lda, ipiv, info, cs);
    p->graphLaunches = g_wg_launches - before;
    g_wg_launches = before;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(cs, &graph);
    cudaStreamDestroy(cs);
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    WG_CUDA(e);
    WG_CUDA(cudaGraphInstantiate(&p->exec, graph, 0));
    cudaGraphDestroy(graph);
    p->gA = A;
    p->glda = lda;
    p->gipiv = ipiv;
    p->ginfo = info;
  }
  WG_CUDA(cudaGraphLaunch(p->exec, st));
  g_wg_launches += p->graphLaunches;
  return WG_OK;
}

int lu_solve(LuPlan* p, const double* LU, long long lda, double* b, cudaStream_t st) {
  WG_REQUIRE(p && LU && b && lda >= p->n, "lu_solve: bad arguments");
  const int n = p->n, NB = p->nb;
  lu_rhs_perm_kernel<<<1, 1024, 0, st>>>(b, p->permSrcTop, p->permBDst, p->permBSrc, p->permNBot, n, NB);
  WG_LAUNCH_CHECK();
  const int BS = 256;
  for (int kb = 0; kb < n; kb += BS) {
    const int bs = BS < n - kb ? BS : n - kb;
    if (kb > 0) {
      lu_gemv_sub_kernel<<<wg_cdiv(bs, 8), 256, 0, st>>>(LU, lda, b, b, kb, bs, 0, kb);
      WG_LAUNCH_CHECK();
    }
    lu_trsv_lower_block_kernel<<<1, 256, 0, st>>>(LU, lda, b, kb, bs);
    WG_LAUNCH_CHECK();
  }
  const int nblk = wg_cdiv(n, BS);
  for (int q = nblk - 1; q >= 0; --q) {
    const int kb = q * BS;
    const int bs = BS < n - kb ? BS : n - kb;
    if (kb + bs < n) {
      lu_gemv_sub_kernel<<<wg_cdiv(bs, 8), 256, 0, st>>>(LU, lda, b, b, kb, bs, kb + bs, n);
      WG_LAUNCH_CHECK();
    }
    lu_trsv_upper_block_kernel<<<1, 256, 0, st>>>(LU, lda, b, kb, bs);
    WG_LAUNCH_CHECK();
  }
  return WG_OK;
}

}  // namespace wg
