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
//     warp redux + one st.async push of every CTA's candidate row into every CTA's shared memory
//     (completion through a transaction mbarrier), and never moves rows while factoring (implicit
//     pivoting); at the end it writes rows to their pivoted positions and moves the rest of those
//     rows inside the panel;
//   * the permutation of a whole panel is composed once on the device and applied to other
//     columns by two fully parallel kernels (gather / write) that take a two-segment column map;
//   * U12 = L11^-1 A12 by recursive TRSM (GEMM + register-resident 32-wide base);
//   * trailing update A22 -= L21 U12 on the FP64 tensor pipe.
// Schedule (factor_enqueue): a highest-priority CHAIN stream does everything the next panel needs
// (its columns' interchanges, TRSM, update, then its factorisation) while one or two BULK streams
// update all other columns; optionally a fourth stream carries the forward solve of a right-hand
// side along.  The whole thing is captured once per buffer set and replayed as a CUDA graph whose
// kernel nodes keep their priorities.  No host synchronisation anywhere: pivots, permutations and
// `info` stay on the device.
#include <limits.h>
#include <stdarg.h>

#include <vector>

#include "wg_common.cuh"
#include "wg_kernels.h"

namespace wg {

constexpr int PT = 512;       // threads per leaf-panel CTA
constexpr int PNW = PT / 32;  // warps per leaf-panel CTA
constexpr int MAXCL = 16;     // largest cluster
constexpr int MAXEXT = 254;   // widest "rest of the panel" a leaf has to move (NB - 2, the narrowest leaf)

static long long* g_dbg_buf = nullptr;  // developer probe: timestamps of the leaf starting at column g_dbg_c0
static int g_dbg_c0 = -1;
void lu_set_debug(long long* buf, int c0) {
  g_dbg_buf = buf;
  g_dbg_c0 = c0;
}
// Scheduling knobs.  Every plan carries its own copy (LuPlan::t); the process-wide setters below change the
// DEFAULTS, which a plan adopts the next time it is used unless it has been tuned individually
// (lu_plan_set_tuning).  A change bumps an epoch, which also invalidates the plan's captured CUDA graphs (the
// knobs are baked into a graph at capture time).
struct LuTuning {
  int nb = 256;
  int useGraph = 1;   // replay the factorisation as a CUDA graph (node priorities keep the chain ahead)
  int lookahead = 1;
  int split = 50;     // two bulk streams while the trailing matrix is large: far part starts at this % of n (0: off)
  int hiRest = 8192;  // bulk work moves to the high-priority bulk stream once the trailing matrix is this small
  int leaf32 = 0;
  int forceWb = 0;    // developer knob: force the leaf width class (4: the 8-rows-per-thread variant)
  int waveSolve = 1;  // triangular solves as two persistent wavefront launches (0: one launch per 128-row block)
};
static LuTuning g_tune;
static long long g_tune_epoch = 1;
static int g_profile = 0;
static int g_trace = 0;
// 32-column leaves (fully rotating register window) whenever one row per thread fits the cluster.  Off by
// default: measured 3 550-3 800 cycles per column against 2 800 for 16-column leaves (the candidate row is
// twice as long), which cancels the eight TRSM+GEMM pairs per panel it saves.
void lu_set_leaf32(int on) {
  g_tune.leaf32 = on;
  ++g_tune_epoch;
}
void lu_set_profile(int on) { g_profile = on; }
void lu_set_trace(int on) { g_trace = on; }
void lu_set_tuning(int nb, int use_graph) {
  if (nb >= 32 && nb <= 256 && nb % 16 == 0) g_tune.nb = nb;
  g_tune.useGraph = use_graph;
  ++g_tune_epoch;
}
void lu_set_lookahead(int on) {
  g_tune.lookahead = on;
  ++g_tune_epoch;
}
void lu_set_wave_solve(int on) {
  g_tune.waveSolve = on;
  ++g_tune_epoch;
}

// Strided batch of independent systems factored in lock step (one kernel launch serves all of them): the
// system index is a grid dimension of every kernel; one system = count 1.
struct Bt {
  int count = 1;
  long long sA = 0;  // doubles between consecutive operators
  long long sV = 0;  // doubles between consecutive right-hand sides
  int sI = 0;        // ints between consecutive systems' row-index tables (ipiv, permutations): n
  int sP = 0;        // ints between consecutive systems' per-panel counters
};

struct PanelArgs {
  double* A;
  long long sA;  // batch strides (system = blockIdx.y)
  int sI;
  long long lda;
  int n;
  int c0, w;      // leaf columns [c0, c0+w)
  int pc0, pc1;   // enclosing top-level panel columns
  int* ipiv;
  int* info;
  int* grpSrcTop;  // [n] : row (position at leaf start) that ends at position c0+i
  int* grpBDst;    // [n] : final position of the row that started at c0+i, or -1 if it stayed in the top w
  int vec;
  long long* dbg;  // optional phase timestamps (developer probe), may be nullptr
};

// bit pattern of |x| (non-negative as a signed integer for every x, NaN included)
__device__ __forceinline__ long long abs_bits(double x) {
  long long r;
  asm("and.b64 %0, %1, 0x7fffffffffffffff;\n" : "=l"(r) : "l"(__double_as_longlong(x)));
  return r;
}

// Lane holding the maximum of (key descending, pos ascending) over the lanes with valid != 0.
// Fast path: one REDUX on the high word of the key (sign, exponent, top mantissa bits) is almost
// always decisive; exact ties fall back to the full 64-bit key and then to the smallest position
// (LAPACK idamax picks the first maximal entry).
__device__ __forceinline__ int argmax_lane(long long key, int pos, bool valid) {
  const unsigned full = 0xffffffffu;
  const int hi = valid ? (int)(key >> 32) : INT_MIN;
  const int mhi = __reduce_max_sync(full, hi);
  unsigned cand = __ballot_sync(full, valid && hi == mhi);
  if (__popc(cand) <= 1) return cand ? __ffs(cand) - 1 : 0;
  const unsigned lo = (unsigned)key;
  const bool c1 = valid && hi == mhi;
  const unsigned mlo = __reduce_max_sync(full, c1 ? lo : 0u);
  cand = __ballot_sync(full, c1 && lo == mlo);
  if (__popc(cand) == 1) return __ffs(cand) - 1;
  const bool c2 = c1 && lo == mlo;
  const int mp = __reduce_min_sync(full, c2 ? pos : INT_MAX);
  cand = __ballot_sync(full, c2 && pos == mp);
  return __ffs(cand) - 1;
}

// remote 8-byte store that completes 8 transaction bytes on an mbarrier of the destination CTA
__device__ __forceinline__ void st_async_u64(uint32_t remote_addr, unsigned long long v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];\n" ::"r"(remote_addr),
               "l"(v), "r"(remote_bar)
               : "memory");
}

template <int W, int R>
struct LeafSmem {
  static constexpr int SL = W + 2;           // row values, key bits, position
  static constexpr int TC = W > 16 ? 16 : W; // columns per pass through the transposition tile
  static constexpr int TW = TC + 1;          // padded tile row
  static constexpr int MAXMOVE = 2 * W;      // most rows a leaf can move per CTA
  double warpSlot[PNW][SL];          // per-warp pivot candidates of the current column
  double cluSlot[2][MAXCL][SL];      // per-CTA candidates, double buffered by column parity
  double moveBuf[MAXMOVE][MAXEXT];   // rest-of-panel data of rows that change position
  double tile[PNW][R][32][TW];       // load/store transposition; W <= 16: also parks the finished columns
  unsigned long long bars[2];
  int moveDst[MAXMOVE];
  int moveCount;
};

// One leaf of the panel recursion: rows [c0, n) x columns [c0, c0+w), w <= W, factored by one cluster.
// Thread t of CTA r owns rows c0 + k*CL*PT + r*PT + t (k < R).  The not-yet-eliminated columns of a
// row live in a ROTATING register window (a[k][0] is always the current column), so the column loop is
// a real loop with static register indices and a small instruction footprint (a fully unrolled body is
// ~180 KB of SASS and stalls on instruction fetch).  W <= 16: finished columns are parked in the
// shared-memory tile that is also used for the coalesced load and store of the block.  W = 32 (one row per
// thread): the window rotates fully -- the finished entry re-enters at the tail, candidate rows are pushed
// with their finished tail zeroed so the update leaves it untouched -- and after W steps the registers hold
// the finished row in column order; the tile only transposes 16 columns at a time.
#ifdef WG_LEAF_CHECKS
__device__ int g_leaf_fail[8 + 48];
#endif
int lu_leaf_fail_read(int* out) {
#ifdef WG_LEAF_CHECKS
  return cudaMemcpyFromSymbol(out, g_leaf_fail, sizeof(int) * 56) == cudaSuccess ? 0 : -2;
#else
  for (int i = 0; i < 56; ++i) out[i] = 0;
  return 0;
#endif
}

template <int W, int R>
__global__ void __launch_bounds__(PT, 1) lu_panel_leaf_kernel(const PanelArgs p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using SM = LeafSmem<W, R>;
  SM& sm = *reinterpret_cast<SM*>(smraw);
  constexpr int SL = SM::SL;
  constexpr int TC = SM::TC;
  constexpr bool ROT = (W > 16);    // fully rotating window (finished entries stay in registers)
  constexpr int LPR = TC / 2;       // lanes per row when moving double2
  constexpr int RPI = 32 / LPR;     // rows per warp instruction

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned rank = cluster_ctarank(), CL = cluster_nctarank();
  const int c0 = p.c0, w = p.w;
  const int m = p.n - c0;
  const long long lda = p.lda;
  const int bz = blockIdx.y;   // system of the batch: one cluster per system
  double* __restrict__ A = p.A + (long long)bz * p.sA;
  int* __restrict__ ipiv = p.ipiv + (long long)bz * p.sI;
  int* __restrict__ info = p.info + bz;
  int* __restrict__ grpSrcTop = p.grpSrcTop + (long long)bz * p.sI;
  int* __restrict__ grpBDst = p.grpBDst + (long long)bz * p.sI;
  const int nLeft = c0 - p.pc0;
  const int ext = nLeft + (p.pc1 - (c0 + w));
  const bool dbg = p.dbg != nullptr && rank == 0 && tid == 0 && bz == 0;
  if (dbg) p.dbg[0] = clock64();

  if (tid == 0) {
    mbar_init(reinterpret_cast<uint64_t*>(&sm.bars[0]), 1);
    mbar_init(reinterpret_cast<uint64_t*>(&sm.bars[1]), 1);
    fence_mbar_init();
    sm.moveCount = 0;
  }
  // the top w rows always leave or change position: stage their rest-of-panel columns right away
  if (rank == 0 && ext > 0) {
    for (int rr = warp; rr < w && rr < m; rr += PNW) {
      const double* srow = A + (long long)(c0 + rr) * lda;
      for (int e = lane; e < ext; e += 32) {
        const int col = e < nLeft ? p.pc0 + e : c0 + w + (e - nLeft);
        cp_async8(&sm.moveBuf[rr][e], srow + col, true);
      }
    }
  }
  cp_async_commit();

  double a[R][W];
  int pos[R], spos[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int lr = k * (int)(CL * PT) + (int)rank * PT + warp * 32 + lane;
    pos[k] = lr < m ? c0 + lr : -1;
    spos[k] = pos[k];
  }
#pragma unroll
  for (int h = 0; h < W / TC; ++h) {   // TC columns per pass through the transposition tile
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int wbase = k * (int)(CL * PT) + (int)rank * PT + warp * 32;  // first row of this warp's 32
      const int lr = wbase + lane;
      const bool valid = lr < m;
      if (p.vec) {
#pragma unroll
        for (int i = 0; i < LPR; ++i) {
          const int r = i * RPI + lane / LPR, cp = lane % LPR;
          double2 v = make_double2(0.0, 0.0);
          if (wbase + r < m && h * TC + 2 * cp < w)
            v = *reinterpret_cast<const double2*>(A + (long long)(c0 + wbase + r) * lda + c0 + h * TC + 2 * cp);
          sm.tile[warp][k][r][2 * cp] = v.x;
          sm.tile[warp][k][r][2 * cp + 1] = v.y;
        }
      } else {
        const double* src = A + (long long)(c0 + lr) * lda + c0 + h * TC;
#pragma unroll
        for (int c = 0; c < TC; ++c) sm.tile[warp][k][lane][c] = (valid && h * TC + c < w) ? src[c] : 0.0;
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
      for (int c = 0; c < TC; ++c) a[k][h * TC + c] = sm.tile[warp][k][lane][c];
    __syncwarp();
  }
  cluster_sync_all();  // mbarriers of every CTA initialised before anybody signals them
  if (dbg) p.dbg[22] = clock64();

  const uint32_t barBase = smem_u32(&sm.bars[0]);

#pragma unroll 1
  for (int j = 0; j < w; ++j) {
    const int dj = c0 + j;
    const int par = j & 1;
    const bool dbgc = dbg && j == 8;
    if (dbgc) p.dbg[24] = clock64();
    if (tid == 0) mbar_expect_tx(reinterpret_cast<uint64_t*>(&sm.bars[par]), (uint32_t)(CL * SL * 8));
    // ---- thread-local, then warp-level candidate (current column = window slot 0)
    long long bk = -1;
    int bp = INT_MAX, kb = 0;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const bool act = pos[k] >= dj;
      // |a| as an integer key.  The sign bit is cleared with an opaque integer AND: fabs() -- and equally a C++
      // `bits & 0x7fff...`, which the compiler rewrites into fabs -- is a DADD on this machine and returns the
      // canonical NaN 0xFFF8... for a NaN operand; as a signed integer that loses against the -1 of an inactive
      // row and hands the column an out-of-range "pivot position" (an infinite entry is enough: inf - inf).
      const long long key = act ? abs_bits(a[k][0]) : -1LL;
      const int pp = act ? pos[k] : INT_MAX;
      if (key > bk || (key == bk && pp < bp)) {
        bk = key;
        bp = pp;
        kb = k;
      }
    }
    const int wl = argmax_lane(bk, bp, true);
    if (dbgc) p.dbg[25] = clock64();
    if (lane == wl) {
      double* slot = sm.warpSlot[warp];
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (k == kb) {
#pragma unroll
          for (int c = 0; c < W; ++c) slot[c] = a[k][c];
        }
      reinterpret_cast<long long*>(slot)[W] = bk;
      reinterpret_cast<long long*>(slot)[W + 1] = (long long)bp;
    }
    if (dbgc) p.dbg[26] = clock64();
    __syncthreads();
    if (dbgc) p.dbg[27] = clock64();
    // ---- CTA-level winner, pushed to every CTA of the cluster (st.async completes their mbarrier)
    if (warp * 32 < SL * (int)CL) {
      const double* s = sm.warpSlot[lane & (PNW - 1)];
      const long long k2 = reinterpret_cast<const long long*>(s)[W];
      const int p2 = (int)reinterpret_cast<const long long*>(s)[W + 1];
      const int ww = argmax_lane(k2, p2, lane < PNW);
      const unsigned long long* srcw = reinterpret_cast<const unsigned long long*>(sm.warpSlot[ww]);
      const uint32_t base = smem_u32(&sm.cluSlot[par][rank][0]);
      if constexpr (SL * MAXCL <= PT) {
        if (tid < SL * (int)CL) {
          const int dest = tid / SL, e = tid - dest * SL;
          st_async_u64(dsmem_map(base + 8u * e, dest), srcw[e], dsmem_map(barBase + 8u * par, dest));
        }
      } else {
        for (int idx = tid; idx < SL * (int)CL; idx += PT) {   // SL * CL exceeds the CTA (W = 32, 16 CTAs)
          const int dest = idx / SL, e = idx - dest * SL;
          unsigned long long v = srcw[e];
          if (ROT && e >= W - j && e < W) v = 0ull;   // finished tail of the window: must not take part in the update
          st_async_u64(dsmem_map(base + 8u * e, dest), v, dsmem_map(barBase + 8u * par, dest));
        }
      }
    }
    if (dbgc) p.dbg[28] = clock64();
    mbar_wait(reinterpret_cast<uint64_t*>(&sm.bars[par]), (uint32_t)((j >> 1) & 1));
    if (dbgc) p.dbg[29] = clock64();
    // ---- cluster-level winner = pivot row (window-aligned: urow[0] is the pivot element)
    const double* urow;
    {
      const double* s = sm.cluSlot[par][lane & ((int)CL - 1)];
      const long long k3 = reinterpret_cast<const long long*>(s)[W];
      const int p3 = (int)reinterpret_cast<const long long*>(s)[W + 1];
      urow = sm.cluSlot[par][argmax_lane(k3, p3, lane < (int)CL)];
    }
    const long long mkey = reinterpret_cast<const long long*>(urow)[W];
    // (a winner that is not an active row -- key < 0 -- cannot happen for any input now that the keys are
    // non-negative; the guard keeps row positions in range whatever the candidates hold)
    int mpos = (int)reinterpret_cast<const long long*>(urow)[W + 1];
    if (mkey < 0) mpos = dj;
    const bool sing = (mkey <= 0);
    const double rinv = sing ? 0.0 : 1.0 / urow[0];
#ifdef WG_LEAF_CHECKS
    if ((mpos < dj || mpos >= p.n) && atomicCAS(&g_leaf_fail[0], 0, 1) == 0) {
      g_leaf_fail[1] = c0; g_leaf_fail[2] = j; g_leaf_fail[3] = mpos; g_leaf_fail[4] = (int)(mkey >> 32);
      g_leaf_fail[5] = (int)rank; g_leaf_fail[6] = tid; g_leaf_fail[7] = (int)CL;
      for (int q2 = 0; q2 < PNW; ++q2) {
        g_leaf_fail[8 + q2] = (int)(reinterpret_cast<const long long*>(sm.warpSlot[q2])[W] >> 32);
        g_leaf_fail[24 + q2] = (int)reinterpret_cast<const long long*>(sm.warpSlot[q2])[W + 1];
      }
      g_leaf_fail[40] = (int)(reinterpret_cast<const long long*>(sm.cluSlot[par][0])[W] >> 32);
      g_leaf_fail[41] = (int)reinterpret_cast<const long long*>(sm.cluSlot[par][0])[W + 1];
      g_leaf_fail[42] = par; g_leaf_fail[43] = dj; g_leaf_fail[44] = pos[0]; g_leaf_fail[45] = (int)(bk >> 32); g_leaf_fail[46] = w; g_leaf_fail[47] = m;
    }
    if (mpos < dj || mpos >= p.n) mpos = dj;
#endif
    if (dbgc) p.dbg[30] = clock64() + (rinv == 12345.678 ? 1 : 0);
    // the pivot row comes from below the top block: stage its rest-of-panel columns now
    if (ext > 0) {
      int sp = -1;
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (pos[k] == mpos && spos[k] >= c0 + w) sp = spos[k];
      const unsigned bal = __ballot_sync(0xffffffffu, sp >= 0);
      if (bal) {
        const int L = __ffs(bal) - 1;
        sp = __shfl_sync(0xffffffffu, sp, L);
        int slot = 0;
        if (lane == L) {
          slot = W + atomicAdd(&sm.moveCount, 1);
#ifdef WG_LEAF_CHECKS
          if (slot >= SM::MAXMOVE && atomicCAS(&g_leaf_fail[0], 0, 2) == 0) {
            g_leaf_fail[1] = c0; g_leaf_fail[2] = j; g_leaf_fail[3] = slot; g_leaf_fail[4] = sp; g_leaf_fail[5] = mpos;
          }
          if (slot >= SM::MAXMOVE) slot = SM::MAXMOVE - 1;
#endif
          sm.moveDst[slot] = dj;
        }
        slot = __shfl_sync(0xffffffffu, slot, L);
        const double* srow = A + (long long)sp * lda;
        for (int e = lane; e < ext; e += 32) {
          const int col = e < nLeft ? p.pc0 + e : c0 + w + (e - nLeft);
          cp_async8(&sm.moveBuf[slot][e], srow + col, true);
        }
        cp_async_commit();
      }
    }
    if (dbgc) p.dbg[31] = clock64();
    if constexpr (!ROT) {
      double uv[W];
#pragma unroll
      for (int c = 1; c < W; ++c) uv[c] = urow[c];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (pos[k] == mpos)
          pos[k] = dj;
        else if (pos[k] == dj)
          pos[k] = mpos;
        const bool elim = pos[k] > dj && !sing;
        const double l = elim ? a[k][0] * rinv : a[k][0];   // multiplier, or the finished U / L entry
        sm.tile[warp][k][lane][j] = l;
        const double ml = elim ? -l : 0.0;
#pragma unroll
        for (int c = 1; c < W; ++c) a[k][c - 1] = fma(ml, uv[c], a[k][c]);  // eliminate + rotate window
      }
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        if (pos[k] == mpos)
          pos[k] = dj;
        else if (pos[k] == dj)
          pos[k] = mpos;
        const bool elim = pos[k] > dj && !sing;
        const double l = elim ? a[k][0] * rinv : a[k][0];
        const double ml = elim ? -l : 0.0;
        // the pushed row carries zeros in its finished tail, so the FMA is a plain move there
#pragma unroll
        for (int c = 1; c < W; ++c) a[k][c - 1] = fma(ml, urow[c], a[k][c]);
        a[k][W - 1] = l;
      }
    }
    if (dbg) p.dbg[2 + (j & 15)] = clock64();
    if (rank == 0 && tid == 0) {
      ipiv[dj] = mpos;
      if (sing && *info == 0) *info = dj + 1;
    }
  }
  if (dbg) p.dbg[1] = clock64();

  if constexpr (ROT) {
    // a leaf narrower than W: finish the rotation so that column i sits in register i
#pragma unroll 1
    for (int j = w; j < W; ++j) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const double t0 = a[k][0];
#pragma unroll
        for (int c = 1; c < W; ++c) a[k][c - 1] = a[k][c];
        a[k][W - 1] = t0;
      }
    }
  }

  // ---- rows go to their pivoted positions.  Unmoved rows of a warp are 32 consecutive matrix rows:
  //      the tile is stored so that every instruction writes whole 128-byte lines.
  __syncwarp();
#pragma unroll
  for (int h = 0; h < W / TC; ++h) {
    if constexpr (ROT) {
#pragma unroll
      for (int k = 0; k < R; ++k)
#pragma unroll
        for (int c = 0; c < TC; ++c) sm.tile[warp][k][lane][c] = a[k][h * TC + c];
      __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int wbase = k * (int)(CL * PT) + (int)rank * PT + warp * 32;
#ifdef WG_LEAF_CHECKS
      if (pos[k] >= p.n) {
        if (atomicCAS(&g_leaf_fail[0], 0, 3) == 0) { g_leaf_fail[1] = c0; g_leaf_fail[3] = pos[k]; g_leaf_fail[4] = spos[k]; g_leaf_fail[6] = tid; }
        pos[k] = spos[k];
      }
#endif
      const bool moved = pos[k] >= 0 && pos[k] != spos[k];
      if (p.vec) {
        const unsigned movedMask = __ballot_sync(0xffffffffu, moved);
#pragma unroll
        for (int i = 0; i < LPR; ++i) {
          const int r = i * RPI + lane / LPR, cp = lane % LPR;
          if (wbase + r < m && h * TC + 2 * cp < w && !((movedMask >> r) & 1u))
            *reinterpret_cast<double2*>(A + (long long)(c0 + wbase + r) * lda + c0 + h * TC + 2 * cp) =
                make_double2(sm.tile[warp][k][r][2 * cp], sm.tile[warp][k][r][2 * cp + 1]);
        }
        if (moved) {
          double* dst = A + (long long)pos[k] * lda + c0 + h * TC;
#pragma unroll
          for (int c = 0; c < TC; c += 2)
            if (h * TC + c < w)
              *reinterpret_cast<double2*>(dst + c) =
                  make_double2(sm.tile[warp][k][lane][c], sm.tile[warp][k][lane][c + 1]);
        }
      } else if (pos[k] >= 0) {
        double* dst = A + (long long)pos[k] * lda + c0 + h * TC;
#pragma unroll
        for (int c = 0; c < TC; ++c)
          if (h * TC + c < w) dst[c] = sm.tile[warp][k][lane][c];
      }
    }
    if constexpr (ROT) __syncwarp();
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const bool moved = pos[k] >= 0 && pos[k] != spos[k];
    if (pos[k] >= 0) {
      if (pos[k] < c0 + w) grpSrcTop[pos[k]] = spos[k];
      if (spos[k] < c0 + w) {
        grpBDst[spos[k]] = (pos[k] >= c0 + w) ? pos[k] : -1;
        sm.moveDst[spos[k] - c0] = moved ? pos[k] : -1;  // only threads of CTA 0 own top rows
      }
    }
  }
  if (dbg) p.dbg[20] = clock64();

  // ---- rest-of-panel columns of the moved rows: all reads (cp.async above) done, barrier, then write
  if (ext > 0) {  // uniform
    cp_async_wait<0>();
    cluster_sync_all();
    const int cnt = sm.moveCount;
    const int nslots = W + cnt;
    for (int s = warp; s < nslots; s += PNW) {
      if (s < W && (rank != 0 || s >= w || s >= m)) continue;
      const int d = sm.moveDst[s];
      if (d < 0) continue;
#ifdef WG_LEAF_CHECKS
      if (d >= p.n) {
        if (lane == 0 && atomicCAS(&g_leaf_fail[0], 0, 4) == 0) { g_leaf_fail[1] = c0; g_leaf_fail[2] = s; g_leaf_fail[3] = d; g_leaf_fail[4] = cnt; }
        continue;
      }
#endif
      double* drow = A + (long long)d * lda;
      for (int e = lane; e < ext; e += 32) {
        const int col = e < nLeft ? p.pc0 + e : c0 + w + (e - nLeft);
        drow[col] = sm.moveBuf[s][e];
      }
    }
  }
  if (dbg) p.dbg[21] = clock64();
}

template <int W, int R>
static size_t leaf_smem() {
  return sizeof(LeafSmem<W, R>) + 16;
}

template <int W, int R>
static int launch_leaf(const PanelArgs& pa, int CL, int batch, cudaStream_t st) {
  auto kern = lu_panel_leaf_kernel<W, R>;
  const size_t smem = leaf_smem<W, R>();
  if (wg_func_configured(reinterpret_cast<const void*>(kern), smem) < (long long)smem) {
    WG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, batch, 1);
  cfg.blockDim = dim3(PT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (g_wg_launch_prio != INT_MIN) {
    attr[1].id = cudaLaunchAttributePriority;
    attr[1].val.priority = g_wg_launch_prio;
    cfg.numAttrs = 2;
  }
  WG_CUDA(cudaLaunchKernelEx(&cfg, kern, pa));
  WG_LAUNCH_CHECK();
  return WG_OK;
}

// ------------------------------------------------------------------ panel permutation: compose
// Composes the leaf permutations of panel [k, k+nb) into
//   srcTop[k+i]  : row (position before the panel) whose content ends at position k+i
//   bDst/bSrc[k+t], t < nbot : position bDst receives the content that was at row bSrc (a top row)
__global__ void lu_perm_compose_kernel(int n, int k, int nb, int wb, const int* __restrict__ grpSrcTop,
                                       const int* __restrict__ grpBDst, int* __restrict__ posContent,
                                       int* __restrict__ srcTop, int* __restrict__ bDst,
                                       int* __restrict__ bSrc, int* __restrict__ nbotOut, int sI, int sP) {
  {
    const long long o = (long long)blockIdx.x * sI;   // system of the batch
    grpSrcTop += o;
    grpBDst += o;
    posContent += o;
    srcTop += o;
    bDst += o;
    bSrc += o;
    nbotOut += (long long)blockIdx.x * sP;
  }
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
// The row permutation of panel [k, k+nb) is applied to a set of columns given as two segments:
// c' in [0, ncol) -> col = c' < split ? c' + off0 : c' + off1.  The critical chain (next panel's columns)
// and the bulk (everything else outside the panel) use it with different maps and staging buffers.
struct ColMap {
  int ncol, split, off0, off1;
};

struct PermBatch {   // strides of the batch (system = blockIdx.z)
  long long sA, sS;
  int sI, sP;
};

__global__ void lu_perm_gather_kernel(const double* __restrict__ A, long long lda, ColMap cm, int k,
                                      const int* __restrict__ srcTop, const int* __restrict__ bSrc,
                                      const int* __restrict__ nbotPtr, double* __restrict__ S,
                                      double* __restrict__ S2, int vec, PermBatch pb) {
  A += (long long)blockIdx.z * pb.sA;
  S += (long long)blockIdx.z * pb.sS;
  S2 += (long long)blockIdx.z * pb.sS;
  srcTop += (long long)blockIdx.z * pb.sI;
  bSrc += (long long)blockIdx.z * pb.sI;
  nbotPtr += (long long)blockIdx.z * pb.sP;
  const int i = blockIdx.y;
  const int ncol = cm.ncol;
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
      const int col = c < cm.split ? c + cm.off0 : c + cm.off1;
      if (do1) *reinterpret_cast<double2*>(s1 + c) = *reinterpret_cast<const double2*>(r1 + col);
      if (do2) *reinterpret_cast<double2*>(s2 + c) = *reinterpret_cast<const double2*>(r2 + col);
    }
  } else {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncol; c += gridDim.x * blockDim.x) {
      const int col = c < cm.split ? c + cm.off0 : c + cm.off1;
      if (do1) s1[c] = r1[col];
      if (do2) s2[c] = r2[col];
    }
  }
}

__global__ void lu_perm_write_kernel(double* __restrict__ A, long long lda, ColMap cm, int k,
                                     const int* __restrict__ srcTop, const int* __restrict__ bDst,
                                     const int* __restrict__ nbotPtr, const double* __restrict__ S,
                                     const double* __restrict__ S2, int vec, PermBatch pb) {
  A += (long long)blockIdx.z * pb.sA;
  S += (long long)blockIdx.z * pb.sS;
  S2 += (long long)blockIdx.z * pb.sS;
  srcTop += (long long)blockIdx.z * pb.sI;
  bDst += (long long)blockIdx.z * pb.sI;
  nbotPtr += (long long)blockIdx.z * pb.sP;
  const int i = blockIdx.y;
  const int ncol = cm.ncol;
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
      const int col = c < cm.split ? c + cm.off0 : c + cm.off1;
      if (do1) *reinterpret_cast<double2*>(r1 + col) = *reinterpret_cast<const double2*>(s1 + c);
      if (do2) *reinterpret_cast<double2*>(r2 + col) = *reinterpret_cast<const double2*>(s2 + c);
    }
  } else {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncol; c += gridDim.x * blockDim.x) {
      const int col = c < cm.split ? c + cm.off0 : c + cm.off1;
      if (do1) r1[col] = s1[c];
      if (do2) r2[col] = s2[c];
    }
  }
}

// ------------------------------------------------------------------ TRSM base (unit lower, in place)
template <int W>
__global__ void __launch_bounds__(64) lu_trsm_base_kernel(const double* __restrict__ L, long long ldl,
                                                         double* __restrict__ B, long long ldb, int w,
                                                         int ncols, long long sA) {
  L += (long long)blockIdx.y * sA;   // system of the batch
  B += (long long)blockIdx.y * sA;
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
  // column-oriented substitution: all updates of one step are independent FMAs
#pragma unroll
  for (int k = 0; k < W - 1; ++k) {
#pragma unroll
    for (int i = k + 1; i < W; ++i) x[i] = fma(-Ls[i][k], x[k], x[i]);
  }
#pragma unroll
  for (int i = 0; i < W; ++i)
    if (i < w) B[(long long)i * ldb + col] = x[i];
}

// Unit-lower TRSM for up to 256 rows, one CTA per strip of 32 right-hand-side columns held in shared
// memory: B <- L^-1 B.  Blocked forward substitution with 32-row blocks: the diagonal block is solved
// by one warp (one column per lane, registers), the rows below are updated with DMMA 8x8x4 tiles whose
// A fragments (rows of L) come straight from global/L2 and whose B/C fragments live in the strip.
constexpr int TSC = 32;        // strip columns
constexpr int TLDB = TSC + 4;  // strip leading dimension (== 4 mod 16: conflict-free fragments)

__global__ void __launch_bounds__(256, 2) lu_trsm_strip_kernel(const double* __restrict__ L, long long ldl,
                                                           double* __restrict__ B, long long ldb, int w,
                                                           int ncols, long long sA) {
  L += (long long)blockIdx.y * sA;   // system of the batch
  B += (long long)blockIdx.y * sA;
  extern __shared__ __align__(16) unsigned char smraw[];
  double* Bs = reinterpret_cast<double*>(smraw);  // [wpad][TLDB]
  __shared__ double Ld[32][33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int col0 = blockIdx.x * TSC;
  const int wpad = (w + 31) & ~31;
  const bool cok = col0 + lane < ncols;
  for (int r = warp; r < wpad; r += 8)
    Bs[r * TLDB + lane] = (r < w && cok) ? B[(long long)r * ldb + col0 + lane] : 0.0;
  for (int r0 = 0; r0 < w; r0 += 32) {
    for (int e = tid; e < 32 * 32; e += 256) {
      const int r = e >> 5, c = e & 31;
      Ld[r][c] = (r0 + r < w && c < r) ? L[(long long)(r0 + r) * ldl + r0 + c] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      double x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = Bs[(r0 + i) * TLDB + lane];
#pragma unroll
      for (int k = 0; k < 31; ++k) {
#pragma unroll
        for (int i = k + 1; i < 32; ++i) x[i] = fma(-Ld[i][k], x[k], x[i]);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) Bs[(r0 + i) * TLDB + lane] = x[i];
    }
    __syncthreads();
    // rows below: C[8 x 32] -= Lrows[8 x 32] * X[32 x 32], one 8-row group per warp iteration
    for (int i0 = r0 + 32 + warp * 8; i0 < wpad; i0 += 64) {
      double acc[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const double2 cv = *reinterpret_cast<const double2*>(&Bs[(i0 + g) * TLDB + nt * 8 + 2 * t]);
        acc[nt][0] = cv.x;
        acc[nt][1] = cv.y;
      }
      const bool rok = i0 + g < w;
      const double* lrow = L + (long long)(rok ? i0 + g : 0) * ldl + r0 + t;
      double af[8];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) af[kk] = rok ? -__ldg(lrow + kk * 4) : 0.0;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const double bf = Bs[(r0 + kk * 4 + t) * TLDB + nt * 8 + g];
          dmma884(acc[nt][0], acc[nt][1], af[kk], bf);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        *reinterpret_cast<double2*>(&Bs[(i0 + g) * TLDB + nt * 8 + 2 * t]) = make_double2(acc[nt][0], acc[nt][1]);
    }
    __syncthreads();
  }
  for (int r = warp; r < w; r += 8)
    if (cok) B[(long long)r * ldb + col0 + lane] = Bs[r * TLDB + lane];
}

static inline GemmBatch gemm_batch(const Bt& bt) {
  GemmBatch gb;
  gb.count = bt.count;
  gb.sC = gb.sA = gb.sB = bt.sA;   // all three operands live inside the system's own matrix
  return gb;
}

static int trsm_rec(const double* L, long long ldl, double* B, long long ldb, int w, int ncols,
                    cudaStream_t st, const Bt& bt) {
  if (w <= 0 || ncols <= 0) return WG_OK;
  if (w <= 32) {
    const dim3 grid(wg_cdiv(ncols, 64), bt.count);
    if (w <= 8)
      WG_CUDA(launch_prio(lu_trsm_base_kernel<8>, grid, dim3(64), 0, st, L, ldl, B, ldb, w, ncols, bt.sA));
    else if (w <= 16)
      WG_CUDA(launch_prio(lu_trsm_base_kernel<16>, grid, dim3(64), 0, st, L, ldl, B, ldb, w, ncols, bt.sA));
    else
      WG_CUDA(launch_prio(lu_trsm_base_kernel<32>, grid, dim3(64), 0, st, L, ldl, B, ldb, w, ncols, bt.sA));
    WG_LAUNCH_CHECK();
    return WG_OK;
  }
  if (w <= 256) {
    const size_t smem = (size_t)((w + 31) & ~31) * TLDB * sizeof(double);
    const size_t smemMax = 256 * TLDB * sizeof(double);
    if (wg_func_configured(reinterpret_cast<const void*>(lu_trsm_strip_kernel), smemMax) < (long long)smemMax)
      WG_CUDA(cudaFuncSetAttribute(lu_trsm_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemMax));
    WG_CUDA(launch_prio(lu_trsm_strip_kernel, dim3(wg_cdiv(ncols, TSC), bt.count), dim3(256), smem, st, L, ldl, B, ldb, w,
                        ncols, bt.sA));
    WG_LAUNCH_CHECK();
    return WG_OK;
  }
  const int w1 = 32 * (((w + 31) / 32 + 1) / 2);
  int rc = trsm_rec(L, ldl, B, ldb, w1, ncols, st, bt);
  if (rc) return rc;
  rc = dgemm_sub(B + (long long)w1 * ldb, L + (long long)w1 * ldl, B, w - w1, ncols, w1, ldb, ldl, ldb, st, gemm_batch(bt));
  if (rc) return rc;
  return trsm_rec(L + (long long)w1 * ldl + w1, ldl, B + (long long)w1 * ldb, ldb, w - w1, ncols, st, bt);
}

// ------------------------------------------------------------------ solve kernels
__global__ void lu_rhs_perm_kernel(double* __restrict__ b, const int* __restrict__ srcTop,
                                   const int* __restrict__ bDst, const int* __restrict__ bSrc,
                                   const int* __restrict__ nbot, int n, int NB, long long sV, int sI, int sP) {
  b += (long long)blockIdx.x * sV;   // system of the batch
  srcTop += (long long)blockIdx.x * sI;
  bDst += (long long)blockIdx.x * sI;
  bSrc += (long long)blockIdx.x * sI;
  nbot += (long long)blockIdx.x * sP;
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

// the interchanges of ONE panel applied to the right-hand side (forward solve fused into the factorisation)
__global__ void lu_rhs_perm_panel_kernel(double* __restrict__ b, const int* __restrict__ srcTop,
                                         const int* __restrict__ bDst, const int* __restrict__ bSrc,
                                         const int* __restrict__ nbotPtr, int k, int nb, long long sV, int sI,
                                         int sP) {
  b += (long long)blockIdx.x * sV;   // system of the batch
  srcTop += (long long)blockIdx.x * sI;
  bDst += (long long)blockIdx.x * sI;
  bSrc += (long long)blockIdx.x * sI;
  nbotPtr += (long long)blockIdx.x * sP;
  const int tid = threadIdx.x;
  const int nbt = *nbotPtr;
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
}

// One step of the blocked triangular solve (BS = 128 rows per block).  Forward (unit lower) visits
// blocks top-down, backward (upper) bottom-up.  In the step for block [kb, kb+bs):
//   CTA 0     : applies the previous block's solution x_prev to its own rows, then solves the
//               diagonal block held in shared memory (32-wide sub-blocks, warp-shuffle substitution);
//   CTAs >= 1 : apply x_prev to all remaining rows, one warp per row (rows of L / U are contiguous).
// One launch per block; the kernel boundary is the only synchronisation.
constexpr int TS_BS = 128;
constexpr int TS_LD = TS_BS + 1;

// Solve of one 128-row diagonal block held in shared memory (T: [bs][TS_LD], ys: right-hand side -> solution),
// organised for LATENCY: the solve sits on the critical path of the whole triangular solve.
//  * Lower (unit) triangle: the four 32x32 diagonal sub-blocks are inverted beforehand (trisolve_prepare, off the
//    critical path: the factors are static) so each sub-block costs one 32x32 matrix-vector product instead of a
//    32-step dependent substitution.  L's sub-blocks are unit triangular with multipliers bounded by 1 (partial
//    pivoting), so their inverses are benign; the solution is refined afterwards anyway (wg_refine).
//  * Upper triangle: the diagonal sub-blocks carry the conditioning of the matrix, so they are solved by
//    substitution (warp shuffle), with every row pre-scaled by the reciprocal of its diagonal entry so that the
//    dependent chain per unknown is one shuffle + one FMA.
//  * Between sub-blocks the remaining rows are updated by pairs of threads (16 columns each, two FMA chains).
// Shared by the per-block step kernel and the wavefront kernel, so both produce the same bits.
constexpr int TS_LI = 33;   // leading dimension of an inverted 32x32 sub-block
constexpr size_t TS_SMEM = (size_t)(TS_BS * TS_LD + 2 * TS_BS + 4 * 32 * TS_LI) * sizeof(double);

// Linv[w] = inverse of the w-th diagonal 32x32 sub-block of the unit lower triangle of T (warps 0..3)
__device__ __forceinline__ void trisolve_prepare_lower(const double* __restrict__ T, double* __restrict__ Linv, int bs,
                                                       int lane, int warp) {
  if (warp >= 4) return;
  const int sb = warp * 32;
  if (sb >= bs) return;
  // lane j builds column j of X = L^-1:  x[r] = delta_rj - sum_{c<r} L[r][c] x[c]
  double x[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    double acc = (r == lane) ? 1.0 : 0.0;
    if (sb + r < bs) {
#pragma unroll
      for (int c = 0; c < r; ++c) acc = fma(-T[(sb + r) * TS_LD + sb + c], x[c], acc);
    }
    x[r] = acc;
    Linv[(warp * 32 + r) * TS_LI + lane] = acc;
  }
}

template <bool UPPER>
__device__ __forceinline__ void trisolve_diag_block(const double* __restrict__ T, const double* __restrict__ Linv,
                                                    double* __restrict__ ys, int bs, int tid, int lane, int warp) {
  const int nsb = (bs + 31) / 32;
  for (int q = 0; q < nsb; ++q) {
    const int sb = UPPER ? (nsb - 1 - q) * 32 : q * 32;
    const int sw = min(32, bs - sb);
    if (warp == 0) {
      if (!UPPER) {
        // y = Linv * r : row `lane`, four independent FMA chains over the 32 columns
        const double* li = Linv + (sb + lane) * TS_LI;
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          p0 = fma(li[c], (c < sw) ? ys[sb + c] : 0.0, p0);
          p1 = fma(li[c + 1], (c + 1 < sw) ? ys[sb + c + 1] : 0.0, p1);
          p2 = fma(li[c + 2], (c + 2 < sw) ? ys[sb + c + 2] : 0.0, p2);
          p3 = fma(li[c + 3], (c + 3 < sw) ? ys[sb + c + 3] : 0.0, p3);
        }
        const double yi = (p0 + p1) + (p2 + p3);
        __syncwarp();
        if (lane < sw) ys[sb + lane] = yi;
      } else {
        double t[32];
        double dg = 1.0;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          t[r] = (lane < sw && r < sw) ? T[(sb + lane) * TS_LD + sb + r] : 0.0;
          if (lane == r) dg = t[r];
        }
        const double dinv = (lane < sw) ? 1.0 / dg : 1.0;
#pragma unroll
        for (int r = 0; r < 32; ++r) t[r] *= dinv;     // row scaled by 1/diagonal: off the dependent chain
        double zi = (lane < sw ? ys[sb + lane] : 0.0) * dinv;
#pragma unroll
        for (int r = 31; r >= 0; --r) {
          if (r < sw) {
            const double yr = __shfl_sync(0xffffffffu, zi, r);
            if (lane < r) zi = fma(-t[r], yr, zi);
          }
        }
        if (lane < sw) ys[sb + lane] = zi;
      }
    }
    __syncthreads();
    // remaining rows of the block: two threads per row, 16 columns each (at most 96 rows: one pass of 256 threads)
    {
      const int nrem = UPPER ? sb : bs - sb - sw;
      const int rr = tid >> 1, half = tid & 1;
      const bool act = rr < nrem;
      const int i = UPPER ? rr : sb + sw + rr;
      double sum = 0.0;
      if (act) {
        const double* trow = T + i * TS_LD + sb + 16 * half;
        const double* yv = ys + sb + 16 * half;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          if (16 * half + c < sw) s0 = fma(trow[c], yv[c], s0);
          if (16 * half + c + 1 < sw) s1 = fma(trow[c + 1], yv[c + 1], s1);
        }
        sum = s0 + s1;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      if (act && half == 0) ys[i] -= sum;
    }
    __syncthreads();
  }
}

template <bool UPPER>
__global__ void __launch_bounds__(256) lu_trisolve_step_kernel(const double* __restrict__ LU, long long lda,
                                                               double* __restrict__ b, int n, int kprev,
                                                               int bsprev, int kb, int bs, long long sA,
                                                               long long sV) {
  LU += (long long)blockIdx.y * sA;   // system of the batch
  b += (long long)blockIdx.y * sV;
  extern __shared__ __align__(16) unsigned char smraw[];
  double* T = reinterpret_cast<double*>(smraw);  // [TS_BS][TS_LD]
  double* xs = T + TS_BS * TS_LD;                // [TS_BS] previous block's solution
  double* ys = xs + TS_BS;                       // [TS_BS] this block's right-hand side / solution
  double* Linv = ys + TS_BS;                     // [4][32][TS_LI] inverted diagonal sub-blocks (forward solve)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (kprev >= 0)
    for (int c = tid; c < bsprev; c += 256) xs[c] = b[kprev + c];
  if (blockIdx.x > 0) {
    __syncthreads();
    const int idx = (blockIdx.x - 1) * 8 + warp;
    const int r = UPPER ? idx : kb + bs + idx;
    if (r >= (UPPER ? kb : n)) return;
    const double* row = LU + (long long)r * lda + kprev;
    double s = 0.0;
    for (int c = lane; c < bsprev; c += 32) s = fma(row[c], xs[c], s);
    s = warp_sum(s);
    if (lane == 0) b[r] -= s;
    return;
  }
  // ---- CTA 0: the diagonal block streams into shared memory (cp.async) while the rows of this block
  //      receive the previous block's update straight from global memory (16 rows per warp, all loads
  //      of a warp in flight together).
  {
    const int c = tid & (TS_BS - 1);
    for (int i = tid >> 7; i < bs; i += 2)
      cp_async8(&T[i * TS_LD + c], LU + (long long)(kb + i) * lda + kb + c, c < bs);
    cp_async_commit();
  }
  if (tid < bs) ys[tid] = b[kb + tid];
  __syncthreads();
  if (kprev >= 0) {
    constexpr int RPW = TS_BS / 8;  // rows per warp
    constexpr int CPL = TS_BS / 32;  // columns per lane
    // branch-free, clamped addresses: all RPW*CPL loads of a lane are issued before the first FMA
    double v[RPW][CPL];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
      const int i = min(warp * RPW + rr, bs - 1);
      const double* row = LU + (long long)(kb + i) * lda + kprev;
#pragma unroll
      for (int u = 0; u < CPL; ++u) v[rr][u] = __ldg(row + min(lane + 32 * u, bsprev - 1));
    }
    double xv[CPL];
#pragma unroll
    for (int u = 0; u < CPL; ++u) xv[u] = (lane + 32 * u < bsprev) ? xs[lane + 32 * u] : 0.0;
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
      double sacc = 0.0;
#pragma unroll
      for (int u = 0; u < CPL; ++u) sacc = fma(v[rr][u], xv[u], sacc);
      const double t = warp_sum(sacc);
      const int i = warp * RPW + rr;
      if (lane == 0 && i < bs) ys[i] -= t;
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  if (!UPPER) {
    trisolve_prepare_lower(T, Linv, bs, lane, warp);
    __syncthreads();
  }
  trisolve_diag_block<UPPER>(T, Linv, ys, bs, tid, lane, warp);
  if (tid < bs) b[kb + tid] = ys[tid];
}

// Whole triangular solve in ONE persistent launch (wavefront over the 128-row blocks).  CTAs take row blocks
// in solve order from an atomic ticket (so a block is only ever waited on by CTAs that started after its owner:
// no co-residency requirement, no deadlock); the owner of block I streams its rows of L (or U) against the
// already published parts of the solution -- most of that work happens long before block I is due -- then,
// once `done` says block I-1 is published, applies it, solves the diagonal block from shared memory and
// publishes its 128 values (st + __threadfence + release store of the counter).  The per-row arithmetic is
// the step kernel's, in the same order (one 128-column dot product per earlier block, subtracted in solve
// order; same diagonal solve), so both kernels give the same bits.  The factors are read exactly once:
// 8*n^2/2 bytes per direction, HBM bound apart from the 128-step substitution chain of each diagonal block.
// sync[0] = ticket, sync[1] = blocks published (per system and direction; zeroed before the launch).
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// Inverses of the 128 x 128 diagonal blocks of L (unit lower) and U, computed once per factorisation (one CTA per
// block and triangle; thread j builds column j of the inverse by substitution, columns live in shared memory).
// With them the diagonal step of the wavefront solve is one 128 x 128 matrix-vector product instead of a 128-step
// dependent substitution -- the same device cuBLAS / MAGMA use for their triangular solves (trtri of the diagonal
// blocks).  The inverse-based solve is not componentwise backward stable the way substitution is; the error it
// adds is of the order cond(block) * eps and is removed, with the LU's own cond(A) * eps, by the refinement step
// that follows every solve (wg_refine).  dinv: [system][block][triangle (0 = L, 1 = U)][128][128], rows and columns
// beyond a short last block are those of the identity.
__global__ void __launch_bounds__(512) lu_diag_invert_kernel(const double* __restrict__ LU, long long lda, int n,
                                                             double* __restrict__ dinv, long long sA, long long sD) {
  // One CTA per diagonal block, both triangles at once and in place in shared memory (LAPACK dtrti2 order):
  // threads 0..255 own the rows of L^-1 (columns right to left), threads 256..511 the rows of U^-1 (columns left
  // to right), TWO threads per row (even / odd terms of the row's dot product, combined with one shuffle).  In
  // step s a thread pair reads only its own row and the original column, which is parked in `col`.
  extern __shared__ __align__(16) unsigned char smraw[];
  double* T = reinterpret_cast<double*>(smraw);   // [TS_BS][TS_LD]
  double* col = T + TS_BS * TS_LD;                // [2 parities][2 triangles][TS_BS]
  const int blk = blockIdx.x;
  LU += (long long)blockIdx.y * sA;
  double* out = dinv + (long long)blockIdx.y * sD + (long long)blk * 2 * TS_BS * TS_BS;
  const int kb = blk * TS_BS;
  const int bs = min(TS_BS, n - kb);
  const int tid = threadIdx.x;
  for (int e = tid; e < TS_BS * TS_BS; e += 512) {
    const int r = e >> 7, c = e & (TS_BS - 1);
    T[r * TS_LD + c] = (r < bs && c < bs) ? LU[(long long)(kb + r) * lda + kb + c] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  const bool up = tid >= 2 * TS_BS;
  const int i = (tid & (2 * TS_BS - 1)) >> 1, half = tid & 1;
  double* row = T + i * TS_LD;
  if (up && half == 0) row[i] = 1.0 / row[i];      // X_ii of the upper triangle (own row, read by nobody else before the barrier)
  __syncthreads();
  for (int s = 0; s < TS_BS - 1; ++s) {
    double* cl = col + (s & 1) * 2 * TS_BS + (up ? TS_BS : 0);
    const int j = up ? 1 + s : TS_BS - 2 - s;
    const bool act = up ? (i < j) : (i > j);
    if (act && half == 0) cl[i] = row[j];
    __syncthreads();
    // lower: x_i = -( l_i + sum_{j < c < i} X[i][c] l_c );  upper: x_i = -( sum_{i <= c < j} X[i][c] u_c ) * X_jj
    const int c0 = up ? i : j + 1, c1 = up ? j : i;      // terms c0 <= c < c1, this thread takes c0 + half, + 2, ...
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (act) {
      int c = c0 + half;
      for (; c + 6 < c1; c += 8) {      // four terms of this thread in flight
        const double r0 = row[c], r1 = row[c + 2], r2 = row[c + 4], r3 = row[c + 6];
        a0 = fma(r0, cl[c], a0);
        a1 = fma(r1, cl[c + 2], a1);
        a2 = fma(r2, cl[c + 4], a2);
        a3 = fma(r3, cl[c + 6], a3);
      }
      for (; c < c1; c += 2) a0 = fma(row[c], cl[c], a0);
    }
    double sum = (a0 + a1) + (a2 + a3);
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    if (act && half == 0) row[j] = up ? -sum * T[j * TS_LD + j] : -(cl[i] + sum);
    // (the other parity of `col` is written next; this one is not touched again before the barrier after that)
  }
  __syncthreads();
  for (int e = tid; e < TS_BS * TS_BS; e += 512) {
    const int r = e >> 7, c = e & (TS_BS - 1);
    const double v = T[r * TS_LD + c];
    out[e] = c < r ? v : (c == r ? 1.0 : 0.0);
    out[TS_BS * TS_BS + e] = c >= r ? v : 0.0;
  }
}

template <bool UPPER>
__global__ void __launch_bounds__(256, 1) lu_trisolve_wave_kernel(const double* __restrict__ LU, long long lda,
                                                                  double* __restrict__ b, int n, int* __restrict__ sync,
                                                                  const double* __restrict__ dinv, long long sA,
                                                                  long long sV, long long sD, int ablate) {
  // ablate: developer timing probe (wrong results when non-zero): 1 skip the diagonal step, 2 do not wait for
  // published blocks, 8 skip the streaming products
  LU += (long long)blockIdx.y * sA;   // system of the batch
  b += (long long)blockIdx.y * sV;
  dinv += (long long)blockIdx.y * sD;
  sync += 4 * blockIdx.y + (UPPER ? 2 : 0);
  extern __shared__ __align__(16) unsigned char smraw[];
  double* T = reinterpret_cast<double*>(smraw);  // [TS_BS][TS_LD] inverse of this block's diagonal block
  double* ys = T + TS_BS * TS_LD;                // [TS_BS] this block's right-hand side
  __shared__ int ticket;
  __shared__ int seenSh;      // published count as read by thread 0
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (n + TS_BS - 1) / TS_BS;
  constexpr int RPW = TS_BS / 8;   // rows per warp
  constexpr int CPL = TS_BS / 32;  // columns per lane
  int seen = 0;                    // blocks known to be published
  for (;;) {
    __syncthreads();               // previous block's use of `ticket`, T and ys is over
    if (tid == 0) ticket = atomicAdd(&sync[0], 1);
    __syncthreads();
    const int q = ticket;          // position in solve order
    if (q >= nblk) return;
    const int blk = UPPER ? nblk - 1 - q : q;
    const int kb = blk * TS_BS;
    const int bs = min(TS_BS, n - kb);
    {  // inverse diagonal block -> shared memory, asynchronously (static data: no dependency)
      const double* src = dinv + ((long long)blk * 2 + (UPPER ? 1 : 0)) * TS_BS * TS_BS;
      for (int e = tid; e < TS_BS * TS_BS; e += 256) cp_async8(&T[(e >> 7) * TS_LD + (e & (TS_BS - 1))], src + e, true);
      cp_async_commit();
    }
    double acc[RPW], part[RPW];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
      acc[rr] = b[kb + min(warp * RPW + rr, bs - 1)];
      part[rr] = 0.0;
    }
    if (ablate & 2) seen = INT_MAX;
    for (int k = 0; k < ((ablate & 8) ? 0 : q); ++k) {          // earlier blocks in solve order
      const int pblk = UPPER ? nblk - 1 - k : k;
      const int kp = pblk * TS_BS;
      const int bsp = min(TS_BS, n - kp);
      // the factor entries do not depend on the solution: all RPW*CPL loads of a lane are in flight before the wait
      double v[RPW][CPL];
#pragma unroll
      for (int rr = 0; rr < RPW; ++rr) {
        const double* row = LU + (long long)(kb + min(warp * RPW + rr, bs - 1)) * lda + kp;
#pragma unroll
        for (int u = 0; u < CPL; ++u) v[rr][u] = __ldcs(row + min(lane + 32 * u, bsp - 1));
      }
      if (k >= seen) {
        // ONE thread per CTA polls the published count and backs off by the number of blocks still ahead of the
        // one it needs; only the CTA that is next in line spins without a pause.  CTA-uniform branch: `seen` is
        // the same in every warp.
        if (tid == 0) {
          int d = ld_acquire_gpu(&sync[1]);
          while (d <= k) {
            const int gap = k - d;                       // whole block steps before block k is even started
            if (gap > 0)
              __nanosleep(gap >= 16 ? 8000 : 500 * gap);
            else if (k + 1 < q)
              __nanosleep(100);
            d = ld_acquire_gpu(&sync[1]);
          }
          seenSh = d;
        }
        __syncthreads();
        seen = seenSh;
        __syncthreads();   // everybody has read it before thread 0 can write the next one
      }
      double xv[CPL];
#pragma unroll
      for (int u = 0; u < CPL; ++u) xv[u] = (lane + 32 * u < bsp) ? __ldcg(b + kp + lane + 32 * u) : 0.0;
      // per-lane partial sums over ALL tiles of the row: the cross-lane reduction happens once per block, not per tile
#pragma unroll
      for (int rr = 0; rr < RPW; ++rr) {
#pragma unroll
        for (int u = 0; u < CPL; ++u)
          if (lane + 32 * u < bsp) part[rr] = fma(v[rr][u], xv[u], part[rr]);
      }
    }
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) acc[rr] -= warp_sum(part[rr]);
    if (lane == 0) {
#pragma unroll
      for (int rr = 0; rr < RPW; ++rr)
        if (warp * RPW + rr < bs) ys[warp * RPW + rr] = acc[rr];
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- diagonal step: y = Dinv * (right-hand side), rows as in the streaming products
    if (!(ablate & 1)) {
      double rv[CPL];
#pragma unroll
      for (int u = 0; u < CPL; ++u) rv[u] = (lane + 32 * u < bs) ? ys[lane + 32 * u] : 0.0;
      double yo = 0.0;
#pragma unroll
      for (int rr = 0; rr < RPW; ++rr) {
        const double* trow = T + (warp * RPW + rr) * TS_LD;
        double sacc = 0.0;
#pragma unroll
        for (int u = 0; u < CPL; ++u) sacc = fma(trow[lane + 32 * u], rv[u], sacc);
        sacc = warp_sum(sacc);
        if (lane == rr) yo = sacc;
      }
      if (lane < RPW && warp * RPW + lane < bs) b[kb + warp * RPW + lane] = yo;
    } else if (tid < bs) {
      b[kb + tid] = ys[tid];
    }
    __syncthreads();               // every thread's store precedes thread 0's release (cumulative at gpu scope)
    if (tid == 0) st_release_gpu(&sync[1], q + 1);
  }
}

// ------------------------------------------------------------------ host-side plan / scheduler
struct LuPlan {
  int n = 0;
  int batch = 1;          // independent systems factored in lock step by every launch (strided batch)
  int npanelSlots = 0;    // per-system stride of the per-panel counters
  long long sStage = 0;   // per-system stride of the bulk staging buffers
  Bt bt;                  // strides of the current call (operator / right-hand side strides come from the caller)
  LuTuning t;              // this plan's scheduling knobs
  long long tuneEpoch = 0; // epoch of the process-wide defaults this plan last adopted
  bool tunePinned = false; // tuned individually: ignores later changes of the defaults
  int maxCL = 16;
  int *grpSrcTop = nullptr, *grpBDst = nullptr;
  int *permSrcTop = nullptr, *permBDst = nullptr, *permBSrc = nullptr, *permNBot = nullptr;
  int* posContent = nullptr;
  int* solveSync = nullptr;                // [batch][4] ticket / published counters of the wavefront solves
  bool haveInv = false;                    // a factorisation of this plan has filled dinv
  double* dinv = nullptr;                  // [batch][nblk][2][128][128] inverses of the diagonal blocks of L and U
  long long sDinv = 0;                     // doubles per system in dinv
  int numSMs = 148;
  double *S = nullptr, *S2 = nullptr;      // row staging of the bulk permutation (256 x n each)
  double *Sc = nullptr, *Sc2 = nullptr;    // row staging of the chain permutation (256 x 256 each)
  cudaStream_t panelStream = nullptr;  // highest priority: the critical chain (next panel's columns + panel)
  cudaStream_t bulkHiStream = nullptr; // bulk work of the latency-bound end of the factorisation
  cudaStream_t farStream = nullptr;    // bulk work of the right-most columns while the trailing matrix is large
  double *Sf = nullptr, *Sf2 = nullptr;
  cudaEvent_t evFar = nullptr, evNear = nullptr;
  cudaStream_t rhsStream = nullptr;    // fused forward solve of the right-hand side (lowest priority)
  cudaEvent_t evRhsIn = nullptr, evRhsOut = nullptr;
  bool rhsPending = false;
  int prioChain = 0, prioBulkHi = 0;
  cudaEvent_t evPhase = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  // optional live timing of the trailing-update GEMM launches (bench roofline)
  std::vector<cudaEvent_t> profEv;
  std::vector<double> profFlops;
  int profCount = 0;
  // developer timeline: per top-level step {step start, fork, wide TRSM done, wide GEMM done} on the main
  // stream and {panel done} on the panel stream
  std::vector<cudaEvent_t> traceEv;
  int traceSteps = 0;
  // CUDA-graph cache of the factorisation for one set of buffers
  cudaGraphExec_t exec = nullptr;
  long long graphLaunches = 0;
  double* gA = nullptr;
  long long glda = 0;
  int* gipiv = nullptr;
  int* ginfo = nullptr;
  double* grhs = nullptr;
  long long gsA = 0, gsV = 0;   // batch strides baked into the factorisation graph
  // forward-solve state while the factorisation is being enqueued with a fused right-hand side
  double* rhs = nullptr;
  int rhsPrev = -1, rhsPrevBs = 0;
  // CUDA-graph cache of the triangular solves: two (factors, right-hand side) pairs
  struct SolveGraph {
    cudaGraphExec_t exec = nullptr;
    long long launches = 0, lda = 0, sA = 0, sV = 0, lastUse = 0;
    const double* LU = nullptr;
    double* b = nullptr;
    int backOnly = 0;
  };
  SolveGraph solveGraphs[2];
  long long solveUse = 0;
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

static int lu_plan_fill(LuPlan* p, int n, int batch);

int lu_plan_create(int n, LuPlan** out, int batch) {
  WG_REQUIRE(n > 0 && out != nullptr && batch >= 1 && batch <= 65535, "lu_plan_create: bad arguments");
  LuPlan* p = new LuPlan();
  const int rc = lu_plan_fill(p, n, batch);
  if (rc) {
    lu_plan_destroy(p);   // frees whatever was allocated before the failure
    return rc;
  }
  *out = p;
  return WG_OK;
}

static int lu_plan_fill(LuPlan* p, int n, int batch) {
  p->n = n;
  p->batch = batch;
  const size_t B = (size_t)batch;
  p->t = g_tune;
  p->tuneEpoch = g_tune_epoch;
  p->maxCL = probe_max_cluster<16, 2>();
  const int npanels = wg_cdiv(n, 32) + 1;
  p->npanelSlots = npanels;
  p->sStage = (long long)256 * n;
  p->bt.count = batch;
  p->bt.sI = n;
  p->bt.sP = npanels;
  WG_CUDA(cudaMalloc(&p->grpSrcTop, sizeof(int) * n * B));
  WG_CUDA(cudaMalloc(&p->grpBDst, sizeof(int) * n * B));
  WG_CUDA(cudaMalloc(&p->permSrcTop, sizeof(int) * n * B));
  WG_CUDA(cudaMalloc(&p->permBDst, sizeof(int) * n * B));
  WG_CUDA(cudaMalloc(&p->permBSrc, sizeof(int) * n * B));
  WG_CUDA(cudaMalloc(&p->permNBot, sizeof(int) * npanels * B));
  WG_CUDA(cudaMalloc(&p->posContent, sizeof(int) * n * B));
  WG_CUDA(cudaMemset(p->posContent, 0xff, sizeof(int) * n * B));
  WG_CUDA(cudaMalloc(&p->solveSync, sizeof(int) * 4 * B));
  p->sDinv = (long long)wg_cdiv(n, TS_BS) * 2 * TS_BS * TS_BS;
  WG_CUDA(cudaMalloc(&p->dinv, sizeof(double) * (size_t)p->sDinv * B));
  {
    int dev = 0;
    WG_CUDA(cudaGetDevice(&dev));
    WG_CUDA(cudaDeviceGetAttribute(&p->numSMs, cudaDevAttrMultiProcessorCount, dev));
  }
  WG_CUDA(cudaMemset(p->permNBot, 0, sizeof(int) * npanels * B));
  const size_t sbytes = sizeof(double) * (size_t)256 * (size_t)n * B;
  WG_CUDA(cudaMalloc(&p->S, sbytes));
  WG_CUDA(cudaMalloc(&p->S2, sbytes));
  WG_CUDA(cudaMalloc(&p->Sc, sizeof(double) * 256 * 256 * B));
  WG_CUDA(cudaMalloc(&p->Sc2, sizeof(double) * 256 * 256 * B));
  int prLow = 0, prHigh = 0;
  WG_CUDA(cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
  WG_CUDA(cudaStreamCreateWithPriority(&p->panelStream, cudaStreamNonBlocking, prHigh));
  p->prioChain = prHigh;
  p->prioBulkHi = prHigh < prLow ? prHigh + 1 : prHigh;
  WG_CUDA(cudaStreamCreateWithPriority(&p->bulkHiStream, cudaStreamNonBlocking, p->prioBulkHi));
  WG_CUDA(cudaEventCreateWithFlags(&p->evPhase, cudaEventDisableTiming));
  WG_CUDA(cudaStreamCreateWithPriority(&p->farStream, cudaStreamNonBlocking, prLow));
  WG_CUDA(cudaEventCreateWithFlags(&p->evFar, cudaEventDisableTiming));
  WG_CUDA(cudaEventCreateWithFlags(&p->evNear, cudaEventDisableTiming));
  WG_CUDA(cudaStreamCreateWithPriority(&p->rhsStream, cudaStreamNonBlocking, prLow));
  WG_CUDA(cudaEventCreateWithFlags(&p->evRhsIn, cudaEventDisableTiming));
  WG_CUDA(cudaEventCreateWithFlags(&p->evRhsOut, cudaEventDisableTiming));
  if (n >= 8192) {   // far-column staging of the split bulk (large systems only)
    WG_CUDA(cudaMalloc(&p->Sf, sbytes));
    WG_CUDA(cudaMalloc(&p->Sf2, sbytes));
  }
  WG_CUDA(cudaEventCreateWithFlags(&p->evFork, cudaEventDisableTiming));
  WG_CUDA(cudaEventCreateWithFlags(&p->evJoin, cudaEventDisableTiming));
  return WG_OK;
}

static void lu_drop_graphs(LuPlan* p) {
  if (p->exec) cudaGraphExecDestroy(p->exec);
  for (auto& c : p->solveGraphs) {
    if (c.exec) cudaGraphExecDestroy(c.exec);
    c.exec = nullptr;
  }
  p->exec = nullptr;
}

void lu_plan_destroy(LuPlan* p) {
  if (!p) return;
  lu_drop_graphs(p);
  if (p->panelStream) cudaStreamDestroy(p->panelStream);
  if (p->bulkHiStream) cudaStreamDestroy(p->bulkHiStream);
  if (p->farStream) cudaStreamDestroy(p->farStream);
  if (p->evFar) cudaEventDestroy(p->evFar);
  if (p->evNear) cudaEventDestroy(p->evNear);
  if (p->rhsStream) cudaStreamDestroy(p->rhsStream);
  if (p->evRhsIn) cudaEventDestroy(p->evRhsIn);
  if (p->evRhsOut) cudaEventDestroy(p->evRhsOut);
  cudaFree(p->Sf);
  cudaFree(p->Sf2);
  if (p->evPhase) cudaEventDestroy(p->evPhase);
  if (p->evFork) cudaEventDestroy(p->evFork);
  if (p->evJoin) cudaEventDestroy(p->evJoin);
  for (cudaEvent_t e : p->profEv) cudaEventDestroy(e);
  for (cudaEvent_t e : p->traceEv) cudaEventDestroy(e);
  cudaFree(p->grpSrcTop);
  cudaFree(p->grpBDst);
  cudaFree(p->permSrcTop);
  cudaFree(p->permBDst);
  cudaFree(p->permBSrc);
  cudaFree(p->permNBot);
  cudaFree(p->posContent);
  cudaFree(p->solveSync);
  cudaFree(p->dinv);
  cudaFree(p->S);
  cudaFree(p->S2);
  cudaFree(p->Sc);
  cudaFree(p->Sc2);
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
  pa.sA = c.p->bt.sA;
  pa.sI = c.p->bt.sI;
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
  pa.dbg = (g_dbg_buf && c0 == g_dbg_c0) ? g_dbg_buf : nullptr;
  auto pick_cl = [&](int rows_per_cta) {
    int cl = 1;
    while (cl * rows_per_cta < m) cl <<= 1;
    return cl;
  };
  // one row per thread whenever the cluster can hold the block: the column loop is issue bound
  // (4 warps per scheduler), so fewer rows per thread is ~20% faster per column
  const int nb_ = c.p->bt.count;
  if (c.wb == 32) return launch_leaf<32, 1>(pa, pick_cl(PT), nb_, c.st);
  if (c.wb == 16) {
    // a lock-step batch whose one-row-per-thread clusters would not be co-resident (one leaf CTA per SM) takes
    // two rows per thread: half the CTAs per system, one wave instead of two
    const bool oneWave = nb_ == 1 || nb_ * pick_cl(PT) <= c.p->numSMs || m <= PT;
    if (m <= c.p->maxCL * PT && oneWave) return launch_leaf<16, 1>(pa, pick_cl(PT), nb_, c.st);
    return launch_leaf<16, 2>(pa, pick_cl(PT * 2), nb_, c.st);
  }
  if (c.wb == 2) return launch_leaf<2, 16>(pa, pick_cl(PT * 16), nb_, c.st);  // up to 131072 rows: 16 rows x 2 columns per thread
  if (c.wb == 4) return launch_leaf<4, 8>(pa, pick_cl(PT * 8), nb_, c.st);   // up to 65536 rows: 8 rows x 4 columns per thread
  if (m <= c.p->maxCL * PT) return launch_leaf<8, 1>(pa, pick_cl(PT), nb_, c.st);
  if (m <= c.p->maxCL * PT * 2) return launch_leaf<8, 2>(pa, pick_cl(PT * 2), nb_, c.st);
  return launch_leaf<8, 4>(pa, pick_cl(PT * 4), nb_, c.st);
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
  rc = trsm_rec(A + (long long)c0 * lda + c0, lda, A + (long long)c0 * lda + c0 + w1, lda, w1, w - w1, c.st, c.p->bt);
  if (rc) return rc;
  rc = dgemm_sub(A + (long long)(c0 + w1) * lda + c0 + w1, A + (long long)(c0 + w1) * lda + c0,
                 A + (long long)c0 * lda + c0 + w1, n - c0 - w1, w - w1, w1, lda, lda, lda, c.st, gemm_batch(c.p->bt));
  if (rc) return rc;
  return panel_rec(c, c0 + w1, w - w1, pc0, pc1);
}

// brackets one trailing-update GEMM with CUDA events when profiling is on (never during graph capture)
static void prof_mark(LuPlan* p, cudaStream_t st, double flops, bool begin) {
  if (!g_profile || p->t.useGraph) return;
  const size_t idx = 2 * (size_t)p->profCount + (begin ? 0 : 1);
  while (p->profEv.size() <= idx) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    p->profEv.push_back(e);
  }
  cudaEventRecord(p->profEv[idx], st);
  if (begin) {
    if (p->profFlops.size() <= (size_t)p->profCount) p->profFlops.resize(p->profCount + 1);
    p->profFlops[p->profCount] = flops;
  } else {
    ++p->profCount;
  }
}

static void trace_mark(LuPlan* p, cudaStream_t st, int step, int slot) {
  if (!g_trace || p->t.useGraph) return;
  const size_t idx = 5 * (size_t)step + slot;
  while (p->traceEv.size() <= idx) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    p->traceEv.push_back(e);
  }
  cudaEventRecord(p->traceEv[idx], st);
  if (step + 1 > p->traceSteps) p->traceSteps = step + 1;
}

// out[5*step + slot] = ms since the first event (NaN where nothing was recorded); returns the step count
int lu_trace_read(LuPlan* p, double* out, int maxSteps) {
  const int ns = p->traceSteps < maxSteps ? p->traceSteps : maxSteps;
  for (int i = 0; i < 5 * ns; ++i) {
    float t = 0.f;
    out[i] = ((size_t)i < p->traceEv.size() && cudaEventElapsedTime(&t, p->traceEv[0], p->traceEv[i]) == cudaSuccess)
                 ? (double)t
                 : -1.0;
  }
  (void)cudaGetLastError();
  return ns;
}

// sums the bracketed launches since the last reset (caller has synchronised the stream)
int lu_profile_read(LuPlan* p, double* ms, double* flops, int* count) {
  double tms = 0.0, tfl = 0.0;
  for (int i = 0; i < p->profCount; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, p->profEv[2 * i], p->profEv[2 * i + 1]) == cudaSuccess) {
      tms += t;
      tfl += p->profFlops[i];
    }
  }
  if (ms) *ms = tms;
  if (flops) *flops = tfl;
  if (count) *count = p->profCount;
  p->profCount = 0;
  return WG_OK;
}

void lu_set_force_wb(int wb) {
  g_tune.forceWb = wb;
  ++g_tune_epoch;
}

static int panel_width_class(LuPlan* p, int m) {
  if (p->t.forceWb == 2 && m <= p->maxCL * PT * 16) return 2;
  if (p->t.forceWb == 4 && m <= p->maxCL * PT * 8) return 4;
  if (p->t.forceWb == 8 && m <= p->maxCL * PT * 4) return 8;
  if (p->t.leaf32 && m <= p->maxCL * PT) return 32;
  if (m <= p->maxCL * PT * 2) return 16;
  if (m <= p->maxCL * PT * 4) return 8;
  if (m <= p->maxCL * PT * 8) return 4;
  if (m <= p->maxCL * PT * 16) return 2;
  return 0;
}

void lu_set_split(int on) {
  g_tune.split = on == 1 ? 50 : on;
  ++g_tune_epoch;
}
void lu_set_bulk_hi(int rest) {
  g_tune.hiRest = rest;
  ++g_tune_epoch;
}

static void lu_drop_graphs(LuPlan* p);

// adopt changed process-wide defaults (unless this plan was tuned individually); captured graphs are dropped
static void lu_sync_tuning(LuPlan* p) {
  if (p->tunePinned || p->tuneEpoch == g_tune_epoch) return;
  p->t = g_tune;
  p->tuneEpoch = g_tune_epoch;
  lu_drop_graphs(p);
}

// per-plan tuning; a negative value keeps the current setting
void lu_plan_set_tuning(LuPlan* p, int nb, int use_graph, int lookahead, int split, int hi_rest) {
  if (nb >= 32 && nb <= 256 && nb % 16 == 0) p->t.nb = nb;
  if (use_graph >= 0) p->t.useGraph = use_graph;
  if (lookahead >= 0) p->t.lookahead = lookahead;
  if (split >= 0) p->t.split = split == 1 ? 50 : split;
  if (hi_rest >= 0) p->t.hiRest = hi_rest;
  p->tunePinned = true;
  lu_drop_graphs(p);
}

static int apply_perm(LuPlan* p, double* A, long long lda, ColMap cm, int k, int nb, int pnl, double* S, double* S2,
                      int vecA, cudaStream_t st, long long sStage) {
  if (cm.ncol <= 0) return WG_OK;
  const int vec = vecA && ((cm.ncol | cm.split | cm.off0 | cm.off1) % 2 == 0);
  dim3 grid(wg_cdiv(cm.ncol, 2048), nb, p->bt.count);
  const PermBatch pb{p->bt.sA, sStage, p->bt.sI, p->bt.sP};
  WG_CUDA(launch_prio(lu_perm_gather_kernel, grid, dim3(256), 0, st, A, lda, cm, k, p->permSrcTop, p->permBSrc,
                      p->permNBot + pnl, S, S2, vec, pb));
  WG_LAUNCH_CHECK();
  WG_CUDA(launch_prio(lu_perm_write_kernel, grid, dim3(256), 0, st, A, lda, cm, k, p->permSrcTop, p->permBDst,
                      p->permNBot + pnl, S, S2, vec, pb));
  WG_LAUNCH_CHECK();
  return WG_OK;
}

static int trisolve_configure() {
  const size_t smem = TS_SMEM;
  if (wg_func_configured(reinterpret_cast<const void*>(lu_trisolve_step_kernel<false>), smem) < (long long)smem) {
    WG_CUDA(cudaFuncSetAttribute(lu_trisolve_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WG_CUDA(cudaFuncSetAttribute(lu_trisolve_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WG_CUDA(cudaFuncSetAttribute(lu_trisolve_wave_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WG_CUDA(cudaFuncSetAttribute(lu_trisolve_wave_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  return WG_OK;
}

// Forward solve fused into the factorisation: once panel [k, k+nb) is factored and its interchanges have
// reached the finished columns left of it, the right-hand side takes the same interchanges and the forward
// substitution advances through the panel's 128-row blocks (the same step kernel as the stand-alone solve).
// It runs on its own lowest-priority stream, after the stream `after` has interchanged the finished columns
// for this panel; rhs_wait() orders the NEXT interchange of those columns behind it.  Nothing else depends on
// it, so it never sits on the critical path.
static int rhs_wait(LuPlan* p, cudaStream_t st) {
  if (p->rhs && p->rhsPending) {
    WG_CUDA(cudaStreamWaitEvent(st, p->evRhsOut, 0));
    p->rhsPending = false;
  }
  return WG_OK;
}

static int rhs_panel(LuPlan* p, const double* LU, long long lda, int k, int nb, int pnl, cudaStream_t after) {
  if (!p->rhs) return WG_OK;
  const int n = p->n;
  const int savedPrio = g_wg_launch_prio;
  g_wg_launch_prio = INT_MIN;
  cudaStream_t st = p->rhsStream;
  WG_CUDA(cudaEventRecord(p->evRhsIn, after));
  WG_CUDA(cudaStreamWaitEvent(st, p->evRhsIn, 0));
  WG_CUDA(launch_prio(lu_rhs_perm_panel_kernel, dim3(p->bt.count), dim3(1024), 0, st, p->rhs, p->permSrcTop, p->permBDst,
                      p->permBSrc, p->permNBot + pnl, k, nb, p->bt.sV, p->bt.sI, p->bt.sP));
  WG_LAUNCH_CHECK();
  const size_t smem = TS_SMEM;
  for (int kb = k; kb < k + nb; kb += TS_BS) {
    const int bs = TS_BS < k + nb - kb ? TS_BS : k + nb - kb;
    const int rest = n - kb - bs;
    const int grid = 1 + ((p->rhsPrev >= 0 && rest > 0) ? wg_cdiv(rest, 8) : 0);
    WG_CUDA(launch_prio(lu_trisolve_step_kernel<false>, dim3(grid, p->bt.count), dim3(256), smem, st, LU, lda, p->rhs, n,
                        p->rhsPrev, p->rhsPrevBs, kb, bs, p->bt.sA, p->bt.sV));
    WG_LAUNCH_CHECK();
    p->rhsPrev = kb;
    p->rhsPrevBs = bs;
  }
  WG_CUDA(cudaEventRecord(p->evRhsOut, st));
  p->rhsPending = true;
  g_wg_launch_prio = savedPrio;
  return WG_OK;
}

// Right-looking blocked LU with a look-ahead critical chain.  After panel k is factored, everything the
// NEXT panel needs -- the row interchanges, the triangular solve and the update of its own columns, then
// its factorisation -- runs on the highest-priority "chain" stream; the bulk of step k (interchanges of all
// other columns, U12, trailing update on the DMMA GEMM) runs beside it.  The chain is latency bound (16 SMs),
// the bulk is throughput bound, so one hides behind the other; once the trailing matrix is small the bulk
// also moves to a high-priority stream so that a second system sharing the GPU cannot starve this one's tail.
static int factor_enqueue(LuPlan* p, double* A, long long lda, int* ipiv, int* info, cudaStream_t st) {
  const int n = p->n, NB = p->t.nb;
  FactorCtx c{p, A, lda, ipiv, info, st, 16, 0};
  c.vecA = (lda % 2 == 0) && (p->bt.sA % 2 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
  const bool look = p->t.lookahead != 0;
  WG_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * p->bt.count, st));
  const Bt& bt = p->bt;
  const GemmBatch gb = gemm_batch(bt);
  const long long sS = p->sStage, sSc = 256 * 256;
  if (panel_width_class(p, n) == 0) {
    wg_set_error("lu_factor: n=%d exceeds the leaf-panel capacity (%d rows) of this build", n,
                 p->maxCL * PT * 16);
    return WG_ERR_UNSUPPORTED;
  }
  struct PrioGuard {
    ~PrioGuard() { g_wg_launch_prio = INT_MIN; }
  } prioGuard;
  p->rhsPrev = -1;
  p->rhsPrevBs = 0;
  p->rhsPending = false;
  if (p->rhs) {
    int rcc = trisolve_configure();
    if (rcc) return rcc;
  }
  cudaStream_t bs = st;    // stream of the bulk work
  int bulkPrio = INT_MIN;
  // While the trailing matrix is large the bulk columns are split at a fixed column X: the near part
  // [.., X) stays on the bulk stream (the chain depends on it), the far part [X, n) runs one step behind on
  // its own low-priority stream, so the latency-bound interchange + TRSM of one part overlaps the other's GEMM.
  const int npan = wg_cdiv(n, NB);
  const int X = (look && p->t.split > 0 && p->t.split < 100 && p->Sf && n >= 8192) ? NB * ((npan * p->t.split + 99) / 100) : n;
  cudaStream_t fs = p->farStream;
  bool farOn = false;
  bool panelDone = false;  // panel [k, k+nb) already factored by the chain of the previous step
  for (int k = 0, pnl = 0; k < n; k += NB, ++pnl) {
    const int nb = NB < n - k ? NB : n - k;
    int rc;
    trace_mark(p, bs, pnl, 0);
    if (!panelDone) {
      c.st = bs;
      c.wb = panel_width_class(p, n - k);
      rc = panel_rec(c, k, nb, k, k + nb);
      if (rc) return rc;
      WG_CUDA(launch_prio(lu_perm_compose_kernel, dim3(bt.count), dim3(256), 0, bs, n, k, nb, c.wb, p->grpSrcTop, p->grpBDst,
                          p->posContent, p->permSrcTop, p->permBDst, p->permBSrc, p->permNBot + pnl, bt.sI, bt.sP));
      WG_LAUNCH_CHECK();
    }
    panelDone = false;
    const int rest = n - k - nb;
    const int nb2 = NB < rest ? NB : rest;
    double* C22 = A + (long long)(k + nb) * lda + k + nb;
    const double* L21 = A + (long long)(k + nb) * lda + k;
    double* U12 = A + (long long)k * lda + k + nb;
    const double* L11 = A + (long long)k * lda + k;
    if (!(look && rest > nb2)) {
      rc = rhs_wait(p, bs);
      if (rc) return rc;
      rc = apply_perm(p, A, lda, ColMap{n - nb, k, 0, nb}, k, nb, pnl, p->S, p->S2, c.vecA, bs, sS);
      if (rc) return rc;
      rc = rhs_panel(p, A, lda, k, nb, pnl, bs);
      if (rc) return rc;
      if (rest <= 0) break;
      rc = trsm_rec(L11, lda, U12, lda, nb, rest, bs, bt);
      if (rc) return rc;
      prof_mark(p, bs, 2.0 * rest * (double)rest * nb * bt.count, true);
      rc = dgemm_sub(C22, L21, U12, rest, rest, nb, lda, lda, lda, bs, gb);
      if (rc) return rc;
      prof_mark(p, bs, 0.0, false);
      continue;
    }
    if (bs == st && rest <= p->t.hiRest) {
      WG_CUDA(cudaEventRecord(p->evPhase, st));
      WG_CUDA(cudaStreamWaitEvent(p->bulkHiStream, p->evPhase, 0));
      bs = p->bulkHiStream;
      bulkPrio = p->prioBulkHi;
    }
    const int R0 = k + nb + nb2;   // first bulk column right of the chain's block
    const bool useFar = X < n && R0 + 2 * NB <= X;
    if (farOn && !useFar) {        // the near part is used up: the far columns come back to the bulk stream
      WG_CUDA(cudaEventRecord(p->evFar, fs));
      WG_CUDA(cudaStreamWaitEvent(bs, p->evFar, 0));
      farOn = false;
    } else if (!farOn && useFar) {
      WG_CUDA(cudaEventRecord(p->evFar, bs));
      WG_CUDA(cudaStreamWaitEvent(fs, p->evFar, 0));
      farOn = true;
    }
    const int Xe = farOn ? X : n;  // end of the near part
    // ---- critical chain: columns of the next panel, then the next panel itself
    cudaStream_t cs = p->panelStream;
    WG_CUDA(cudaEventRecord(p->evFork, bs));
    WG_CUDA(cudaStreamWaitEvent(cs, p->evFork, 0));
    g_wg_launch_prio = p->prioChain;
    rc = apply_perm(p, A, lda, ColMap{nb2, nb2, k + nb, 0}, k, nb, pnl, p->Sc, p->Sc2, c.vecA, cs, sSc);
    if (rc) return rc;
    rc = trsm_rec(L11, lda, U12, lda, nb, nb2, cs, bt);
    if (rc) return rc;
    rc = dgemm_sub(C22, L21, U12, rest, nb2, nb, lda, lda, lda, cs, gb);
    if (rc) return rc;
    c.st = cs;
    c.wb = panel_width_class(p, rest);
    rc = panel_rec(c, k + nb, nb2, k + nb, k + nb + nb2);
    if (rc) return rc;
    WG_CUDA(launch_prio(lu_perm_compose_kernel, dim3(bt.count), dim3(256), 0, cs, n, k + nb, nb2, c.wb, p->grpSrcTop,
                        p->grpBDst, p->posContent, p->permSrcTop, p->permBDst, p->permBSrc, p->permNBot + pnl + 1, bt.sI,
                        bt.sP));
    WG_LAUNCH_CHECK();
    WG_CUDA(cudaEventRecord(p->evJoin, cs));
    trace_mark(p, cs, pnl, 4);
    g_wg_launch_prio = bulkPrio;
    // ---- bulk: right of the chain's block up to Xe; the finished columns left of the panel too unless the
    //      far stream is on (they are read by the far GEMM of the previous step, so it interchanges them itself)
    if (!farOn) {
      rc = rhs_wait(p, bs);
      if (rc) return rc;
    }
    rc = apply_perm(p, A, lda, farOn ? ColMap{Xe - R0, Xe - R0, R0, 0} : ColMap{k + (Xe - R0), k, 0, nb + nb2}, k, nb,
                    pnl, p->S, p->S2, c.vecA, bs, sS);
    if (rc) return rc;
    if (!farOn) {
      rc = rhs_panel(p, A, lda, k, nb, pnl, bs);
      if (rc) return rc;
    }
    trace_mark(p, bs, pnl, 1);
    rc = trsm_rec(L11, lda, U12 + nb2, lda, nb, Xe - R0, bs, bt);
    if (rc) return rc;
    trace_mark(p, bs, pnl, 2);
    prof_mark(p, bs, 2.0 * rest * (double)(Xe - R0) * nb * bt.count, true);
    rc = dgemm_sub(C22 + nb2, L21, U12 + nb2, rest, Xe - R0, nb, lda, lda, lda, bs, gb);
    if (rc) return rc;
    prof_mark(p, bs, 0.0, false);
    trace_mark(p, bs, pnl, 3);
    if (farOn) {   // ---- far columns [X, n) and the finished columns [0, k), running behind the near part
      WG_CUDA(cudaEventRecord(p->evNear, bs));
      g_wg_launch_prio = INT_MIN;
      const int off = X - (k + nb);
      rc = rhs_wait(p, fs);
      if (rc) return rc;
      rc = apply_perm(p, A, lda, ColMap{k + (n - X), k, 0, X - k}, k, nb, pnl, p->Sf, p->Sf2, c.vecA, fs, sS);
      if (rc) return rc;
      rc = rhs_panel(p, A, lda, k, nb, pnl, fs);
      if (rc) return rc;
      rc = trsm_rec(L11, lda, U12 + off, lda, nb, n - X, fs, bt);
      if (rc) return rc;
      rc = dgemm_sub(C22 + off, L21, U12 + off, rest, n - X, nb, lda, lda, lda, fs, gb);
      if (rc) return rc;
      WG_CUDA(cudaStreamWaitEvent(fs, p->evJoin, 0));   // its next step needs the next panel ...
      WG_CUDA(cudaStreamWaitEvent(fs, p->evNear, 0));   // ... and interchanges rows of L21 that this near GEMM reads
      g_wg_launch_prio = bulkPrio;
    }
    WG_CUDA(cudaStreamWaitEvent(bs, p->evJoin, 0));
    panelDone = true;
  }
  if (farOn) {
    WG_CUDA(cudaEventRecord(p->evFar, fs));
    WG_CUDA(cudaStreamWaitEvent(bs, p->evFar, 0));
  }
  {
    int rcw = rhs_wait(p, bs);
    if (rcw) return rcw;
  }
  if (bs != st) {
    WG_CUDA(cudaEventRecord(p->evPhase, bs));
    WG_CUDA(cudaStreamWaitEvent(st, p->evPhase, 0));
  }
  // inverses of the 128 x 128 diagonal blocks of L and U for the wavefront triangular solves
  {
    const size_t smemInv = (size_t)(TS_BS * TS_LD + 4 * TS_BS) * sizeof(double);
    if (wg_func_configured(reinterpret_cast<const void*>(lu_diag_invert_kernel), smemInv) < (long long)smemInv)
      WG_CUDA(cudaFuncSetAttribute(lu_diag_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemInv));
    lu_diag_invert_kernel<<<dim3(wg_cdiv(n, TS_BS), bt.count), 512, smemInv, st>>>(A, lda, n, p->dinv, bt.sA, p->sDinv);
    WG_LAUNCH_CHECK();
  }
  return WG_OK;
}

int lu_factor(LuPlan* p, double* A, long long lda, int* ipiv, int* info, cudaStream_t st, double* rhs,
              long long strideA, long long strideRhs) {
  WG_REQUIRE(p && A && ipiv && info && lda >= p->n, "lu_factor: bad arguments");
  WG_REQUIRE(p->batch == 1 || strideA >= (long long)p->n * lda, "lu_factor: batch stride smaller than one operator");
  p->rhs = rhs;   // non-null: the forward solve L y = P rhs rides along (finish with lu_solve(..., backwardOnly))
  p->haveInv = true;
  lu_sync_tuning(p);
  p->bt.sA = p->batch > 1 ? strideA : 0;
  p->bt.sV = (p->batch > 1 && rhs) ? strideRhs : 0;
  if (!p->t.useGraph) return factor_enqueue(p, A, lda, ipiv, info, st);
  // (the strides are baked into a captured graph, like the pointers)
  if (!p->exec || p->gA != A || p->glda != lda || p->gipiv != ipiv || p->ginfo != info || p->grhs != rhs ||
      p->gsA != p->bt.sA || p->gsV != p->bt.sV) {
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
    WG_CUDA(cudaGraphInstantiateWithFlags(&p->exec, graph, cudaGraphInstantiateFlagUseNodePriority));
    cudaGraphDestroy(graph);
    p->gA = A;
    p->glda = lda;
    p->gipiv = ipiv;
    p->ginfo = info;
    p->grhs = rhs;
    p->gsA = p->bt.sA;
    p->gsV = p->bt.sV;
  }
  WG_CUDA(cudaGraphLaunch(p->exec, st));
  g_wg_launches += p->graphLaunches;
  return WG_OK;
}

static int solve_enqueue(LuPlan* p, const double* LU, long long lda, double* b, cudaStream_t st, bool backwardOnly);

int lu_solve(LuPlan* p, const double* LU, long long lda, double* b, cudaStream_t st, bool backwardOnly,
             long long strideA, long long strideB) {
  WG_REQUIRE(p && LU && b && lda >= p->n, "lu_solve: bad arguments");
  if (!p->haveInv) {   // the wavefront solves use the diagonal-block inverses the factorisation of THIS plan leaves behind
    wg_set_error("lu_solve: no factorisation has been enqueued on this plan yet");
    return WG_ERR_STATE;
  }
  lu_sync_tuning(p);
  p->bt.sA = p->batch > 1 ? strideA : 0;
  p->bt.sV = p->batch > 1 ? strideB : 0;
  if (!p->t.useGraph) return solve_enqueue(p, LU, lda, b, st, backwardOnly);
  // two cached graphs (least recently used is replaced): the solution vector and the refinement residual
  // alternate as right-hand sides of the same factors
  LuPlan::SolveGraph* g = nullptr;
  for (auto& c : p->solveGraphs)
    if (c.exec && c.LU == LU && c.lda == lda && c.b == b && c.backOnly == (int)backwardOnly && c.sA == p->bt.sA &&
        c.sV == p->bt.sV)
      g = &c;
  if (!g) {
    for (auto& c : p->solveGraphs)
      if (!g && !c.exec) g = &c;
    if (!g) g = p->solveGraphs[1].lastUse < p->solveGraphs[0].lastUse ? &p->solveGraphs[1] : &p->solveGraphs[0];
    if (g->exec) {
      cudaGraphExecDestroy(g->exec);
      g->exec = nullptr;
    }
    cudaStream_t cs;
    WG_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    WG_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    const long long before = g_wg_launches;
    int rc = solve_enqueue(p, LU, lda, b, cs, backwardOnly);
    g->launches = g_wg_launches - before;
    g_wg_launches = before;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(cs, &graph);
    cudaStreamDestroy(cs);
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    WG_CUDA(e);
    WG_CUDA(cudaGraphInstantiate(&g->exec, graph, 0));
    cudaGraphDestroy(graph);
    g->LU = LU;
    g->lda = lda;
    g->b = b;
    g->backOnly = (int)backwardOnly;
    g->sA = p->bt.sA;
    g->sV = p->bt.sV;
  }
  g->lastUse = ++p->solveUse;
  WG_CUDA(cudaGraphLaunch(g->exec, st));
  g_wg_launches += g->launches;
  return WG_OK;
}

static int solve_enqueue(LuPlan* p, const double* LU, long long lda, double* b, cudaStream_t st, bool backwardOnly) {
  const int n = p->n, NB = p->t.nb;
  int rcc = trisolve_configure();
  if (rcc) return rcc;
  const size_t smem = TS_SMEM;
  const int nblk = wg_cdiv(n, TS_BS);
  int kprev = -1, bsprev = 0;
  if (!backwardOnly) {
    lu_rhs_perm_kernel<<<p->bt.count, 1024, 0, st>>>(b, p->permSrcTop, p->permBDst, p->permBSrc, p->permNBot, n, NB, p->bt.sV,
                                                     p->bt.sI, p->bt.sP);
    WG_LAUNCH_CHECK();
  }
  if (p->t.waveSolve) {
    // two persistent launches: every CTA takes row blocks from a ticket counter and waits on the published count
    WG_CUDA(cudaMemsetAsync(p->solveSync, 0, sizeof(int) * 4 * p->bt.count, st));
    const dim3 grid(nblk < p->numSMs ? nblk : p->numSMs, p->bt.count);
    if (!backwardOnly) {
      lu_trisolve_wave_kernel<false><<<grid, 256, smem, st>>>(LU, lda, b, n, p->solveSync, p->dinv, p->bt.sA, p->bt.sV,
                                                              p->sDinv, p->t.waveSolve >> 4);
      WG_LAUNCH_CHECK();
    }
    lu_trisolve_wave_kernel<true><<<grid, 256, smem, st>>>(LU, lda, b, n, p->solveSync, p->dinv, p->bt.sA, p->bt.sV,
                                                           p->sDinv, p->t.waveSolve >> 4);
    WG_LAUNCH_CHECK();
    return WG_OK;
  }
  for (int q = 0; q < nblk && !backwardOnly; ++q) {
    const int kb = q * TS_BS;
    const int bs = TS_BS < n - kb ? TS_BS : n - kb;
    const int rest = n - kb - bs;
    const int grid = 1 + ((kprev >= 0 && rest > 0) ? wg_cdiv(rest, 8) : 0);
    lu_trisolve_step_kernel<false><<<dim3(grid, p->bt.count), 256, smem, st>>>(LU, lda, b, n, kprev, bsprev, kb, bs, p->bt.sA,
                                                                              p->bt.sV);
    WG_LAUNCH_CHECK();
    kprev = kb;
    bsprev = bs;
  }
  kprev = -1;
  bsprev = 0;
  for (int q = nblk - 1; q >= 0; --q) {
    const int kb = q * TS_BS;
    const int bs = TS_BS < n - kb ? TS_BS : n - kb;
    const int grid = 1 + ((kprev >= 0 && kb > 0) ? wg_cdiv(kb, 8) : 0);
    lu_trisolve_step_kernel<true><<<dim3(grid, p->bt.count), 256, smem, st>>>(LU, lda, b, n, kprev, bsprev, kb, bs, p->bt.sA,
                                                                             p->bt.sV);
    WG_LAUNCH_CHECK();
    kprev = kb;
    bsprev = bs;
  }
  return WG_OK;
}

}  // namespace wg
