// Device side of the Boltzmann hot path (everything except the LU):
//   * basis / derivative tables on the Gauss-Lobatto grid        (polynomial.py:445-490, 691-862)
//   * per-row kinematics + source term                            (boltzmann.py:651-759)
//   * dense operator assembly, HBM-write bound                    (boltzmann.py:761-817)
//   * basis changes, spectra, truncation, moment quadrature       (boltzmann.py:141-238, 296-491;
//                                                                  polynomial.py:222-298, 539-567)
// Compiled with -fmad=false so that the element formulas round like the NumPy reference
// (separate multiply/add); the kernels here are bandwidth bound, not FLOP bound.
#include <math.h>

#include "wg_boltzmann.h"
#include "wg_common.cuh"

namespace wg {

// ============================================================================ basis tables
// Default WallGo bases: Cardinal in z, restricted Chebyshev in (p_z, p_par), spectral derivatives.
// Nodes: chi_a=-cos(a pi/M), rz_b=-cos(b pi/N), rp_g=-cos(g pi/(N-1)); on these nodes
// T_n(-cos t) = (-1)^n cos(n t) and U_{n-1}(-cos t) = (-1)^{n+1} sin(n t)/sin(t), evaluated with
// cospi/sinpi on an exactly reduced integer argument.
__device__ __forceinline__ double cospi_ratio(long long num, long long den) {
  num %= 2 * den;
  return cospi((double)num / (double)den);
}
__device__ __forceinline__ double sinpi_ratio(long long num, long long den) {
  num %= 2 * den;
  return sinpi((double)num / (double)den);
}

__global__ void wg_basis_kernel(int M, int N, const double* __restrict__ rz, double* __restrict__ Tchi,
                                double* __restrict__ Dchi, double* __restrict__ DchiFull,
                                double* __restrict__ Trz, double* __restrict__ Drz,
                                double* __restrict__ Trp) {
  const int M1 = M - 1, N1 = N - 1;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  // Lagrange-cardinal differentiation matrix on the M+1 Lobatto nodes: DchiFull[j][i] = C_i'(x_j)
  for (int e = tid; e < (M + 1) * (M + 1); e += nth) {
    const int j = e / (M + 1), i = e % (M + 1);
    double v;
    if (i != j) {
      const double dx = 2.0 * sinpi_ratio(i + j, 2 * M) * sinpi_ratio(j - i + 4LL * M, 2 * M);  // x_j - x_i
      const double cj = (j == 0 || j == M) ? 2.0 : 1.0, ci = (i == 0 || i == M) ? 2.0 : 1.0;
      v = (cj / ci) * (((i + j) & 1) ? -1.0 : 1.0) / dx;
    } else if (j == 0) {
      v = -(2.0 * M * M + 1.0) / 6.0;
    } else if (j == M) {
      v = (2.0 * M * M + 1.0) / 6.0;
    } else {
      const double x = -cospi_ratio(j, M), s = sinpi_ratio(j, M);
      v = -x / (2.0 * s * s);
    }
    DchiFull[e] = v;
    if (j >= 1 && j < M && i >= 1 && i < M) {
      Dchi[(j - 1) * M1 + (i - 1)] = v;
      Tchi[(j - 1) * M1 + (i - 1)] = (i == j) ? 1.0 : 0.0;
    }
  }
  for (int e = tid; e < N1 * N1; e += nth) {
    const int b = e / N1 + 1, n = e % N1 + 2;  // rz node b = 1..N-1, order n = 2..N
    const double sgn = (n & 1) ? -1.0 : 1.0;
    const double Tn = sgn * cospi_ratio((long long)n * b, N);
    const double x = rz[b - 1];
    Trz[e] = Tn - ((n & 1) ? x : 1.0);
    const double Un1 = -sgn * sinpi_ratio((long long)n * b, N) / sinpi_ratio(b, N);
    Drz[e] = (double)n * Un1 - ((n & 1) ? 1.0 : 0.0);
    const int g = e / N1, k = e % N1 + 1;  // rp node g = 0..N-2, order k = 1..N-1
    const double Tk = ((k & 1) ? -1.0 : 1.0) * cospi_ratio((long long)k * g, N - 1);
    Trp[e] = Tk - 1.0;
  }
}

// K = X (kron) Y for (N-1)x(N-1) factors: K[(b*N1+g)*q + (j*N1+k)] = X[b][j] * Y[g][k]
__global__ void wg_kron_kernel(int N1, const double* __restrict__ X, const double* __restrict__ Y,
                               double* __restrict__ K) {
  const int q = N1 * N1;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)q * q;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / q), c = (int)(e % q);
    K[e] = X[(r / N1) * N1 + c / N1] * Y[(r % N1) * N1 + c % N1];
  }
}

// ============================================================================ per-solve prologue
// profiles = [ T(M+1) | v(M+1) | msq(P x (M+1)) | dxidchi(M-1) ]; derivs = [ dT | dv | dmsq ] on M+1 nodes
// Batches: blockIdx.y (blockIdx.x for the one-CTA kernels, blockIdx.z for the assembly) is the system of a
// strided batch of independent systems processed in lock step; per-system arrays are contiguous ([B][...]).
__global__ void wg_profile_deriv_kernel(int M, int P, const double* __restrict__ DchiFull,
                                        const double* __restrict__ prof, double* __restrict__ derivs, int nprof) {
  const int L = M + 1;
  prof += (long long)blockIdx.x * nprof;
  derivs += (long long)blockIdx.x * (2 + P) * L;
  for (int e = threadIdx.x; e < (2 + P) * L; e += blockDim.x) {
    const int f = e / L, j = e % L;
    const double* fv = prof + f * L;
    const double* drow = DchiFull + j * L;
    double s = 0.0;
    for (int i = 0; i < L; ++i) s += drow[i] * fv[i];
    derivs[e] = s;
  }
}

__global__ void wg_rows_kernel(BzDims d, const double* __restrict__ prof, const double* __restrict__ derivs,
                               const double* __restrict__ pz, const double* __restrict__ pp,
                               const double* __restrict__ dpzdrz, const double* __restrict__ stat, double vw,
                               double collMult, double liouScale, double* __restrict__ w1, double* __restrict__ w2,
                               double* __restrict__ cT2, double* __restrict__ S, const double* __restrict__ vwArr,
                               int nprof, long long sV) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.n) return;
  const int L = d.M + 1;
  {
    const long long bz = blockIdx.y;
    prof += bz * nprof;
    derivs += bz * (2 + d.P) * L;
    w1 += bz * d.n;
    w2 += bz * d.n;
    cT2 += bz * d.M1;
    S += bz * sV;
    if (vwArr) vw = vwArr[bz];
  }
  const int gmm = r % d.N1, bt = (r / d.N1) % d.N1, al = (r / d.q) % d.M1, a = r / (d.q * d.M1);
  const double T = prof[al + 1], v = prof[L + al + 1], msq = prof[(2 + a) * L + al + 1];
  const double dxidchi = prof[(2 + d.P) * L + al];
  const double dT = derivs[al + 1], dv = derivs[L + al + 1], dmsq = derivs[(2 + a) * L + al + 1];
  const double pzv = pz[bt], ppv = pp[gmm];
  const double E = sqrt(msq + pzv * pzv + ppv * ppv);
  const double gw = 1.0 / sqrt(1.0 - vw * vw);
  const double Pw = gw * (pzv - vw * E);
  const double gpl = 1.0 / sqrt(1.0 - v * v);
  const double Epl = gpl * (E - v * pzv);
  const double Ppl = gpl * (pzv - v * E);
  const double uwupl = gw * gpl * (vw - v);
  const double dchidxi = 1.0 / dxidchi;
  const double drzdpz = 1.0 / dpzdrz[bt];
  const double x = Epl / T;
  const double s = stat[a];
  // boltzmann.py:876-886 ; MAX_EXPONENT = 1024 ln 2
  const double dfeq = (x > 709.782712893384) ? -0.0 : -1.0 / (exp(x) - 2.0 * s + exp(-x));
  S[r] = (dfeq / T) * dchidxi * (Pw * Ppl * (gpl * gpl) * dv + Pw * Epl * dT / T + 0.5 * dmsq * uwupl);
  w1[r] = liouScale * (dchidxi * Pw);
  w2[r] = liouScale * (dchidxi * drzdpz * (gw / 2.0) * dmsq);
  if (r < d.M1) {
    const double Ta = prof[r + 1];
    cT2[r] = collMult * (Ta * Ta);
  }
}

// ============================================================================ operator assembly
// One CTA = R consecutive (beta,gamma) rows at fixed (species a, position alpha); it writes R complete
// rows of the n x n operator.  The R-row tiles of K1 = Trz(x)Trp, K2 = Drz(x)Trp and of the collision
// tensor are staged in shared memory with TMA bulk copies; the K1 tile (scaled by the Liouville row
// weight) then lives in registers and is re-used for all (M-1) z-blocks of the row.
#ifndef WG_ASM_MINB
#define WG_ASM_MINB 2
#endif
// 32-byte global store (SASS STG.E.ENL2.256): half as many store instructions as double2 for the same bytes
__device__ __forceinline__ void st_global_v4(double* p, double a, double b, double c, double e) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(e) : "memory");
}

// V = doubles per thread and store: 2 (16-byte stores) or 4 (32-byte stores; needs lda % 4 == 0 and a
// 32-byte aligned operator)
template <int R, int V>
__global__ void __launch_bounds__(512, WG_ASM_MINB) wg_assemble_kernel(BzDims d, const double* __restrict__ K1,
                                                          const double* __restrict__ K2,
                                                          const double* __restrict__ Coll,
                                                          const double* __restrict__ Tchi,
                                                          const double* __restrict__ Dchi,
                                                          const double* __restrict__ w1,
                                                          const double* __restrict__ w2,
                                                          const double* __restrict__ cT2,
                                                          double* __restrict__ A, long long lda, long long sA,
                                                          int y0) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int q = d.q, M1 = d.M1, P = d.P;
  w1 += (long long)blockIdx.z * d.n;
  w2 += (long long)blockIdx.z * d.n;
  cT2 += (long long)blockIdx.z * M1;
  A += (long long)blockIdx.z * sA;
  // row-panel mode: the grid covers the (species, position) row groups [y0, y0 + gridDim.y) and A points at
  // the first row of the panel (the whole operator: y0 = 0)
  A -= (long long)y0 * q * lda;
  double* K1s = reinterpret_cast<double*>(smraw);  // [R][q]
  double* K2s = K1s + R * q;                       // [R][q]
  double* Cs = K2s + R * q;                        // [R][P*q]
  double* dch = Cs + (size_t)R * P * q;            // [M1]
  double* tch = dch + M1;                          // [M1]
  __shared__ __align__(8) uint64_t bar;

  const int by = blockIdx.y + y0;
  const int a = by / M1, al = by % M1;
  const int bg0 = blockIdx.x * R;
  const long long r0 = ((long long)a * M1 + al) * q + bg0;
  const int tid = threadIdx.x;

  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t kb = (uint32_t)(R * q * sizeof(double));
    const uint32_t cb = (uint32_t)((size_t)R * P * q * sizeof(double));
    mbar_expect_tx(&bar, 2 * kb + cb);
    tma_load_1d(K1s, K1 + (size_t)bg0 * q, kb, &bar);
    tma_load_1d(K2s, K2 + (size_t)bg0 * q, kb, &bar);
    tma_load_1d(Cs, Coll + ((size_t)a * q + bg0) * ((size_t)P * q), cb, &bar);
  }
  for (int i = tid; i < M1; i += blockDim.x) {
    dch[i] = Dchi[al * M1 + i];
    tch[i] = Tchi[al * M1 + i];
  }
  double rw1[R], rw2[R];
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    rw1[rr] = w1[r0 + rr];
    rw2[rr] = w2[r0 + rr];
  }
  const double ct2 = cT2[al];
  __syncthreads();
  mbar_wait(&bar, 0);

  for (int cp = tid; cp < q / V; cp += blockDim.x) {
    double wk[R][V];
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
#pragma unroll
      for (int v = 0; v < V; ++v) wk[rr][v] = rw1[rr] * K1s[rr * q + V * cp + v];
    for (int b = 0; b < P; ++b) {
      const bool same = (b == a);
      double* out = A + r0 * lda + (long long)b * M1 * q + V * cp;
      for (int i = 0; i < M1; ++i, out += q) {
        const double dv = same ? dch[i] : 0.0;
        const double tx = tch[i];
        if (tx == 0.0) {
#pragma unroll
          for (int rr = 0; rr < R; ++rr) {
            if constexpr (V == 4)
              st_global_v4(out + rr * lda, wk[rr][0] * dv, wk[rr][1] * dv, wk[rr][2] * dv, wk[rr][3] * dv);
            else
              *reinterpret_cast<double2*>(out + rr * lda) = make_double2(wk[rr][0] * dv, wk[rr][1] * dv);
          }
        } else {
          const double ctx = ct2 * tx;
#pragma unroll
          for (int rr = 0; rr < R; ++rr) {
            double val[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
              double x = wk[rr][v] * dv;
              if (same) x -= (rw2[rr] * tx) * K2s[rr * q + V * cp + v];
              x += ctx * Cs[((size_t)rr * P + b) * q + V * cp + v];
              val[v] = x;
            }
            if constexpr (V == 4)
              st_global_v4(out + rr * lda, val[0], val[1], val[2], val[3]);
            else
              *reinterpret_cast<double2*>(out + rr * lda) = make_double2(val[0], val[1]);
          }
        }
      }
    }
  }
}

// ============================================================================ residual for iterative refinement
// r = S - A x with A never stored: every operator entry is formed exactly as wg_assemble_kernel forms it (same
// expressions, same roundings: this file is compiled with -fmad=false) and multiplied into a DOUBLE-DOUBLE
// accumulator (error-free product through fma, error-free sum), so the residual of the FP64 system is known
// to ~1e-32 relative to |A||x| before it is rounded.  One solve with the existing LU factors on this residual
// then removes the cond(A)*eps forward error of the first solve (mixed-precision refinement, SURVEY.md
// section 7: "one step of FP64 iterative refinement").  Same CTA shape as the assembly: R rows per CTA.
__device__ __forceinline__ void dd_add_prod(double& hi, double& lo, double a, double b) {
  const double p = a * b;
  const double e = fma(a, b, -p);          // a*b = p + e exactly
  const double s = hi + p;
  const double bb = s - hi;
  const double err = (hi - (s - bb)) + (p - bb);   // hi + p = s + err exactly
  hi = s;
  lo += err + e;
}
__device__ __forceinline__ void dd_add_dd(double& hi, double& lo, double h2, double l2) {
  const double s = hi + h2;
  const double bb = s - hi;
  const double err = (hi - (s - bb)) + (h2 - bb);
  hi = s;
  lo += l2 + err;
}

template <int R, int V>
__global__ void __launch_bounds__(512, 1) wg_residual_kernel(BzDims d, const double* __restrict__ K1,
                                                             const double* __restrict__ K2,
                                                             const double* __restrict__ Coll,
                                                             const double* __restrict__ Tchi,
                                                             const double* __restrict__ Dchi,
                                                             const double* __restrict__ w1,
                                                             const double* __restrict__ w2,
                                                             const double* __restrict__ cT2,
                                                             const double* __restrict__ x, const double* __restrict__ S,
                                                             double* __restrict__ rout, long long sV) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int q = d.q, M1 = d.M1, P = d.P;
  w1 += (long long)blockIdx.z * d.n;
  w2 += (long long)blockIdx.z * d.n;
  cT2 += (long long)blockIdx.z * M1;
  x += (long long)blockIdx.z * sV;
  S += (long long)blockIdx.z * sV;
  rout += (long long)blockIdx.z * sV;
  double* K1s = reinterpret_cast<double*>(smraw);  // [R][q]
  double* K2s = K1s + R * q;                       // [R][q]
  double* Cs = K2s + R * q;                        // [R][P*q]
  double* dch = Cs + (size_t)R * P * q;            // [M1]
  double* tch = dch + M1;                          // [M1]
  __shared__ __align__(8) uint64_t bar;
  __shared__ double redh[16][R], redl[16][R];

  const int a = blockIdx.y / M1, al = blockIdx.y % M1;
  const int bg0 = blockIdx.x * R;
  const long long r0 = ((long long)a * M1 + al) * q + bg0;
  const int tid = threadIdx.x;

  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t kb = (uint32_t)(R * q * sizeof(double));
    const uint32_t cb = (uint32_t)((size_t)R * P * q * sizeof(double));
    mbar_expect_tx(&bar, 2 * kb + cb);
    tma_load_1d(K1s, K1 + (size_t)bg0 * q, kb, &bar);
    tma_load_1d(K2s, K2 + (size_t)bg0 * q, kb, &bar);
    tma_load_1d(Cs, Coll + ((size_t)a * q + bg0) * ((size_t)P * q), cb, &bar);
  }
  for (int i = tid; i < M1; i += blockDim.x) {
    dch[i] = Dchi[al * M1 + i];
    tch[i] = Tchi[al * M1 + i];
  }
  double rw1[R], rw2[R];
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    rw1[rr] = w1[r0 + rr];
    rw2[rr] = w2[r0 + rr];
  }
  const double ct2 = cT2[al];
  __syncthreads();
  mbar_wait(&bar, 0);

  double hi[R], lo[R];
#pragma unroll
  for (int rr = 0; rr < R; ++rr) hi[rr] = lo[rr] = 0.0;
  for (int cp = tid; cp < q / V; cp += blockDim.x) {
    double wk[R][V];
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
#pragma unroll
      for (int v = 0; v < V; ++v) wk[rr][v] = rw1[rr] * K1s[rr * q + V * cp + v];
    for (int b = 0; b < P; ++b) {
      const bool same = (b == a);
      const double* xin = x + (long long)b * M1 * q + V * cp;
      for (int i = 0; i < M1; ++i, xin += q) {
        const double dv = same ? dch[i] : 0.0;
        const double tx = tch[i];
        double xv[V];
#pragma unroll
        for (int v = 0; v < V; ++v) xv[v] = __ldg(xin + v);
        if (tx == 0.0) {
          if (dv == 0.0) continue;   // an exact zero block of the operator
#pragma unroll
          for (int rr = 0; rr < R; ++rr)
#pragma unroll
            for (int v = 0; v < V; ++v) dd_add_prod(hi[rr], lo[rr], wk[rr][v] * dv, xv[v]);
        } else {
          const double ctx = ct2 * tx;
#pragma unroll
          for (int rr = 0; rr < R; ++rr) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              double e = wk[rr][v] * dv;
              if (same) e -= (rw2[rr] * tx) * K2s[rr * q + V * cp + v];
              e += ctx * Cs[((size_t)rr * P + b) * q + V * cp + v];
              dd_add_prod(hi[rr], lo[rr], e, xv[v]);
            }
          }
        }
      }
    }
  }
  // ---- CTA reduction of the R double-double sums, then r = S - (hi + lo)
  const int lane = tid & 31, warp = tid >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double h2 = __shfl_xor_sync(0xffffffffu, hi[rr], o);
      const double l2 = __shfl_xor_sync(0xffffffffu, lo[rr], o);
      dd_add_dd(hi[rr], lo[rr], h2, l2);
    }
    if (lane == 0) {
      redh[warp][rr] = hi[rr];
      redl[warp][rr] = lo[rr];
    }
  }
  __syncthreads();
  if (tid < R) {
    double h = 0.0, l = 0.0;
    for (int w = 0; w < nw; ++w) dd_add_dd(h, l, redh[w][tid], redl[w][tid]);
    const double sv = S[r0 + tid];
    double rh = sv, rl = 0.0;
    dd_add_dd(rh, rl, -h, -l);
    rout[r0 + tid] = rh + rl;
  }
}

// x += dlt; stats[0] = max|dlt|, stats[1] = max|x| (after the update), per system
__global__ void wg_refine_update_kernel(int n, double* __restrict__ x, const double* __restrict__ dlt,
                                        double* __restrict__ stats, long long sV) {
  __shared__ double red[2][32];
  x += (long long)blockIdx.x * sV;
  dlt += (long long)blockIdx.x * sV;
  double md = 0.0, mx = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double dv = dlt[i];
    const double xn = x[i] + dv;
    x[i] = xn;
    md = fmax(md, fabs(dv));   // fmax drops NaN operands: a non-finite correction shows up in max|x| only
    mx = fmax(mx, fabs(xn));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    red[0][warp] = md;
    red[1][warp] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      md = fmax(md, red[0][w]);
      mx = fmax(mx, red[1][w]);
    }
    stats[2 * blockIdx.x] = md;
    stats[2 * blockIdx.x + 1] = mx;
  }
}

// ============================================================================ basis changes / moments
// in viewed as [outer][len][inner];  out[o][x][s] = sum_x' Mat[x][x'] * in[o][x'][s]
__global__ void wg_apply_axis_kernel(const double* __restrict__ in, double* __restrict__ out,
                                     const double* __restrict__ Mat, int outer, int len, int inner) {
  const long long total = (long long)outer * len * inner;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(e % inner), x = (int)((e / inner) % len);
    const long long o = e / ((long long)inner * len);
    const double* src = in + o * len * inner + s;
    const double* m = Mat + (long long)x * len;
    double acc = 0.0;
    for (int xp = 0; xp < len; ++xp) acc += m[xp] * src[(long long)xp * inner];
    out[e] = acc;
  }
}

__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < nw ? red[lane] : 0.0;
    t = warp_sum(t);
  }
  return t;  // valid in warp 0
}

// spectra[0..M1) = sum_{a,j,k}|c| ; [M1..M1+N1) = sum_{a,i,k}|c| ; [M1+N1..M1+2N1) = sum_{a,i,j}|c|
__global__ void wg_spectra_kernel(BzDims d, const double* __restrict__ c, double* __restrict__ spectra) {
  __shared__ double red[32];
  const int bin = blockIdx.x;
  c += (long long)blockIdx.y * d.n;
  spectra += (long long)blockIdx.y * (d.M1 + 2 * d.N1);
  const int M1 = d.M1, N1 = d.N1, q = d.q;
  double s = 0.0;
  if (bin < M1) {
    for (int e = threadIdx.x; e < d.P * q; e += blockDim.x)
      s += fabs(c[((long long)(e / q) * M1 + bin) * q + e % q]);
  } else if (bin < M1 + N1) {
    const int j = bin - M1;
    for (int e = threadIdx.x; e < d.P * M1 * N1; e += blockDim.x)
      s += fabs(c[((long long)(e / N1)) * q + j * N1 + e % N1]);
  } else {
    const int k = bin - M1 - N1;
    for (int e = threadIdx.x; e < d.P * M1 * N1; e += blockDim.x) s += fabs(c[(long long)e * N1 + k]);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) spectra[bin] = s;
}

// zero the truncated tails in place and reduce the truncation-error sums:
//   err[a*4+0] = sum |c| over the kept box, err[a*4+1..3] = sum |c| on the last kept chi / pz / pp slice
__global__ void wg_truncate_kernel(BzDims d, double* __restrict__ c, int keepM, int keepN1, int keepN2,
                                   double* __restrict__ err, const int* __restrict__ keepArr) {
  __shared__ double red[32];
  const int a = blockIdx.x;
  c += (long long)blockIdx.y * d.n;
  err += (long long)blockIdx.y * 4 * d.P;
  if (keepArr) {   // per-system keep counts of a batch
    keepM = keepArr[3 * blockIdx.y];
    keepN1 = keepArr[3 * blockIdx.y + 1];
    keepN2 = keepArr[3 * blockIdx.y + 2];
  }
  const int M1 = d.M1, N1 = d.N1, q = d.q;
  double tot = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
  for (int e = threadIdx.x; e < M1 * q; e += blockDim.x) {
    const int i = e / q, j = (e / N1) % N1, k = e % N1;
    double* p = c + (long long)a * M1 * q + e;
    if (i >= keepM || j >= keepN1 || k >= keepN2) {
      *p = 0.0;
    } else {
      const double v = fabs(*p);
      tot += v;
      if (i == keepM - 1) e1 += v;
      if (j == keepN1 - 1) e2 += v;
      if (k == keepN2 - 1) e3 += v;
    }
  }
  tot = block_sum(tot, red);
  e1 = block_sum(e1, red);
  e2 = block_sum(e2, red);
  e3 = block_sum(e3, red);
  if (threadIdx.x == 0) {
    err[a * 4 + 0] = tot;
    err[a * 4 + 1] = e1;
    err[a * 4 + 2] = e2;
    err[a * 4 + 3] = e3;
  }
}

// Delta_{00,02,20,11}[a][alpha] : Gauss-Chebyshev-Lobatto quadrature of the cardinal-basis deltaF
// (boltzmann.py:199-225, polynomial.py:546-561).  deltas layout [4][P][M1].
__global__ void wg_moments_kernel(BzDims d, const double* __restrict__ card, const double* __restrict__ prof,
                                  const double* __restrict__ rz, const double* __restrict__ rp,
                                  const double* __restrict__ pz, const double* __restrict__ pp,
                                  const double* __restrict__ dpzdrz, const double* __restrict__ dppdrp,
                                  double* __restrict__ deltas, int nprof) {
  __shared__ double red[32];
  const int a = blockIdx.x / d.M1, al = blockIdx.x % d.M1;
  card += (long long)blockIdx.y * d.n;
  prof += (long long)blockIdx.y * nprof;
  deltas += (long long)blockIdx.y * 4 * d.P * d.M1;
  const int N1 = d.N1, q = d.q, L = d.M + 1;
  const double msq = prof[(2 + a) * L + al + 1];
  const double* f = card + ((long long)a * d.M1 + al) * q;
  const double PI = 3.141592653589793;
  double s00 = 0.0, s02 = 0.0, s20 = 0.0, s11 = 0.0;
  for (int e = threadIdx.x; e < q; e += blockDim.x) {
    const int b = e / N1, g = e % N1;
    const double pzv = pz[b], ppv = pp[g];
    const double E = sqrt(msq + pzv * pzv + ppv * ppv);
    const double base = dpzdrz[b] * dppdrp[g] * ppv / (4.0 * (PI * PI) * E);
    double wz = PI / (double)d.N;
    double wp = PI / (double)(d.N - 1);
    if (g == 0) wp = wp / 2.0;
    wz = sqrt(1.0 - rz[b] * rz[b]) * wz;
    wp = sqrt(1.0 - rp[g] * rp[g]) * wp;
    const double fv = f[e];
    s00 += ((base * fv) * wz) * wp;
    s02 += (((pzv * pzv) * base * fv) * wz) * wp;
    s20 += (((E * E) * base * fv) * wz) * wp;
    s11 += (((E * pzv) * base * fv) * wz) * wp;
  }
  s00 = block_sum(s00, red);
  s02 = block_sum(s02, red);
  s20 = block_sum(s20, red);
  s11 = block_sum(s11, red);
  if (threadIdx.x == 0) {
    const int PM = d.P * d.M1, o = a * d.M1 + al;
    deltas[0 * PM + o] = s00;
    deltas[1 * PM + o] = s02;
    deltas[2 * PM + o] = s20;
    deltas[3 * PM + o] = s11;
  }
}

// y = Coll-operator * x : y[a,al,bg] = cT2[al] * sum_{b,i,jk} Tchi[al][i] C[a,bg,b,jk] x[b,i,jk]
// (the collision part of the operator applied to a vector; boltzmann.py:541-545)
__global__ void wg_collision_matvec_kernel(BzDims d, const double* __restrict__ Coll,
                                           const double* __restrict__ Tchi, const double* __restrict__ cT2,
                                           const double* __restrict__ x, double* __restrict__ y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= d.n) return;
  const int q = d.q, M1 = d.M1, P = d.P;
  const int bg = warp % q, al = (warp / q) % M1, a = warp / (q * M1);
  const double* crow = Coll + ((long long)a * q + bg) * ((long long)P * q);
  double s = 0.0;
  for (int i = 0; i < M1; ++i) {
    const double tx = Tchi[al * M1 + i];
    if (tx == 0.0) continue;
    for (int b = 0; b < P; ++b) {
      const double* xv = x + ((long long)b * M1 + i) * q;
      double t = 0.0;
      for (int jk = lane; jk < q; jk += 32) t += crow[b * q + jk] * xv[jk];
      s += tx * t;
    }
  }
  s = warp_sum(s);
  if (lane == 0) y[warp] = cT2[al] * s;
}

// Feedback of the moments into the wall equation of motion (consumers in equationOfMotion.py):
//   T30_out, T33_out per z  : EOM.deltaToTmunu (equationOfMotion.py:2064-2125)
//   potOut = 1/2 sum_a dof_a m2_a Delta00_a              : action term (:1545-1555)
//   dVout[f] = 1/2 sum_a dof_a dm2_a/dphi_f Delta00_a     : EOM source term (:1418-1429)
// out layout: [T30 (M1) | T33 (M1) | potOut (M1) | dVout (M1 x F)]
__global__ void wg_feedback_kernel(BzDims d, const double* __restrict__ deltas, const double* __restrict__ prof,
                                   const double* __restrict__ dofs, const double* __restrict__ dmsq, int F,
                                   double u0, double u3, double* __restrict__ out) {
  const int al = blockIdx.x * blockDim.x + threadIdx.x;
  if (al >= d.M1) return;
  const int PM = d.P * d.M1, L = d.M + 1;
  const double ub0 = u3, ub3 = u0;
  double t30 = 0.0, t33 = 0.0, pot = 0.0;
  for (int a = 0; a < d.P; ++a) {
    const int o = a * d.M1 + al;
    const double D00 = deltas[o], D02 = deltas[PM + o], D20 = deltas[2 * PM + o], D11 = deltas[3 * PM + o];
    const double msq = prof[(2 + a) * L + al + 1];
    const double dof = dofs[a];
    const double A1 = 3.0 * D20 - D02 - msq * D00;
    const double A2 = 3.0 * D02 - D20 + msq * D00;
    t30 += dof * (A1 * u3 * u0 + A2 * ub3 * ub0 + 2.0 * D11 * (u3 * ub0 + ub3 * u0)) / 2.0;
    t33 += dof * ((A1 * u3 * u3 + A2 * ub3 * ub3 + 4.0 * D11 * u3 * ub3) / 2.0 - (msq * D00 + D02 - D20) / 2.0);
    pot += dof * msq * D00;
  }
  out[al] = t30;
  out[d.M1 + al] = t33;
  out[2 * d.M1 + al] = pot / 2.0;
  for (int f = 0; f < F; ++f) {
    double sacc = 0.0;
    for (int a = 0; a < d.P; ++a) sacc += dofs[a] * dmsq[((long long)a * d.M1 + al) * F + f] * deltas[a * d.M1 + al];
    out[3 * d.M1 + al * F + f] = sacc / 2.0;
  }
}

// Collision tensor computed on a finer momentum grid brought to the solver's grid
// (CollisionArray.interpolateCollisionArray, collisionArray.py:391-478): the collocation axes (beta, gamma) are
// evaluated at the target nodes with the Lagrange cardinals of the source grid, the Chebyshev series in the
// polynomial axes (j, k) is truncated.  src [P][nF][nF][P][nF][nF]; out is written in the order
// (x, y, a, c, j, k) -- the reference reshapes the evaluated array to (P, nT, nT, P, nT, nT) WITHOUT a transpose
// (an identity re-labelling only for one species), reproduced here for drop-in fidelity.
__global__ void wg_collision_interp_kernel(int P, int nF, int nT, const double* __restrict__ Lz,
                                           const double* __restrict__ Lp, const double* __restrict__ src,
                                           double* __restrict__ out) {
  const long long total = (long long)nT * nT * P * P * nT * nT;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    long long r = e;
    const int k = (int)(r % nT); r /= nT;
    const int j = (int)(r % nT); r /= nT;
    const int c = (int)(r % P); r /= P;
    const int a = (int)(r % P); r /= P;
    const int y = (int)(r % nT); r /= nT;
    const int x = (int)r;
    double acc = 0.0;
    for (int b = 0; b < nF; ++b) {
      double inner = 0.0;
      for (int g = 0; g < nF; ++g)
        inner += Lp[y * nF + g] * src[((((long long)(a * nF + b) * nF + g) * P + c) * nF + j) * nF + k];
      acc += Lz[x * nF + b] * inner;
    }
    out[e] = acc;
  }
}

int bz_collision_interp(int P, int nF, int nT, const double* Lz, const double* Lp, const double* src, double* out,
                        cudaStream_t st) {
  const long long total = (long long)nT * nT * P * P * nT * nT;
  int blocks = wg_cdiv(total, 256);
  if (blocks > 2368) blocks = 2368;
  wg_collision_interp_kernel<<<blocks, 256, 0, st>>>(P, nF, nT, Lz, Lp, src, out);
  WG_LAUNCH_CHECK();
  return WG_OK;
}

// ============================================================================ host launchers
int bz_feedback(const BzDims& d, const double* deltas, const double* prof, const double* dofs, const double* dmsq,
                int F, double u0, double u3, double* out, cudaStream_t st) {
  wg_feedback_kernel<<<wg_cdiv(d.M1, 64), 64, 0, st>>>(d, deltas, prof, dofs, dmsq, F, u0, u3, out);
  WG_LAUNCH_CHECK();
  return WG_OK;
}

int bz_build_basis(const BzDims& d, const double* rz, double* Tchi, double* Dchi, double* DchiFull,
                   double* Trz, double* Drz, double* Trp, cudaStream_t st) {
  wg_basis_kernel<<<8, 256, 0, st>>>(d.M, d.N, rz, Tchi, Dchi, DchiFull, Trz, Drz, Trp);
  WG_LAUNCH_CHECK();
  return WG_OK;
}

int bz_build_kron(const BzDims& d, const double* Trz, const double* Drz, const double* Trp, double* K1,
                  double* K2, cudaStream_t st) {
  const int blocks = wg_cdiv((long long)d.q * d.q, 256);
  wg_kron_kernel<<<blocks < 1184 ? blocks : 1184, 256, 0, st>>>(d.N1, Trz, Trp, K1);
  WG_LAUNCH_CHECK();
  wg_kron_kernel<<<blocks < 1184 ? blocks : 1184, 256, 0, st>>>(d.N1, Drz, Trp, K2);
  WG_LAUNCH_CHECK();
  return WG_OK;
}

static int g_asm_wide = 1;   // 32-byte stores in the assembly kernel when alignment allows
void bz_set_asm_wide(int on) { g_asm_wide = on; }

int bz_assemble(const BzDims& d, const BzStatic& s, const double* prof, double vw, double collMult, int parts,
                double* derivs, double* w1, double* w2, double* cT2, double* A, long long lda, double* S,
                cudaStream_t st, const BzBatch& bb, int y0, int ny) {
  WG_REQUIRE(d.q % 4 == 0, "assemble: (N-1)^2 must be a multiple of 4 (N odd)");
  if (ny < 0) ny = d.P * d.M1 - y0;
  WG_REQUIRE(y0 >= 0 && ny >= 1 && y0 + ny <= d.P * d.M1, "assemble: row-group range out of bounds");
  WG_REQUIRE(lda % 2 == 0 && bb.sA % 2 == 0 && reinterpret_cast<uintptr_t>(A) % 16 == 0,
             "assemble: A must be 16B aligned, lda and the batch stride even");
  const int nprof = (2 + d.P) * (d.M + 1) + d.M1;
  wg_profile_deriv_kernel<<<bb.count, 256, 0, st>>>(d.M, d.P, s.DchiFull, prof, derivs, nprof);
  WG_LAUNCH_CHECK();
  wg_rows_kernel<<<dim3(wg_cdiv(d.n, 256), bb.count), 256, 0, st>>>(d, prof, derivs, s.pz, s.pp, s.dpzdrz, s.stat, vw,
                                                                   (parts & 2) ? collMult : 0.0, (parts & 1) ? 1.0 : 0.0,
                                                                   w1, w2, cT2, S, bb.vwArr, nprof, bb.sV);
  WG_LAUNCH_CHECK();
  constexpr int R = 4;
  const size_t smem = ((size_t)R * d.q * (2 + d.P) + 2 * d.M1) * sizeof(double);
  WG_REQUIRE(smem <= 200 * 1024, "assemble: basis tile does not fit in shared memory");
  if (wg_func_configured(reinterpret_cast<const void*>(wg_assemble_kernel<R, 4>), smem) < (long long)smem) {
    WG_CUDA(cudaFuncSetAttribute(wg_assemble_kernel<R, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WG_CUDA(cudaFuncSetAttribute(wg_assemble_kernel<R, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WG_CUDA(cudaFuncSetAttribute(wg_assemble_kernel<R, 2>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
    WG_CUDA(cudaFuncSetAttribute(wg_assemble_kernel<R, 4>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
  }
  dim3 grid(d.q / R, ny, bb.count);
  const bool wide = g_asm_wide && d.q % 4 == 0 && lda % 4 == 0 && bb.sA % 4 == 0 && reinterpret_cast<uintptr_t>(A) % 32 == 0;
  const int per = wide ? 4 : 2;
  int threads = ((d.q / per + 31) / 32) * 32;
  if (threads > 512) threads = 512;
  if (threads < 64) threads = 64;
  if (wide)
    wg_assemble_kernel<R, 4><<<grid, threads, smem, st>>>(d, s.K1, s.K2, s.Coll, s.Tchi, s.Dchi, w1, w2, cT2, A, lda, bb.sA, y0);
  else
    wg_assemble_kernel<R, 2><<<grid, threads, smem, st>>>(d, s.K1, s.K2, s.Coll, s.Tchi, s.Dchi, w1, w2, cT2, A, lda, bb.sA, y0);
  WG_LAUNCH_CHECK();
  return WG_OK;
}

// r = S - A x for the operator of (prof, vw, collMult): the per-row weights and the source are re-evaluated
// (O(n)), the operator entries are regenerated on the fly (never stored); Ssrc and r are n-vectors per system
int bz_residual(const BzDims& d, const BzStatic& s, const double* prof, double vw, double collMult, double* derivs,
                double* w1, double* w2, double* cT2, const double* x, double* Ssrc, double* r, cudaStream_t st,
                const BzBatch& bb) {
  WG_REQUIRE(d.q % 4 == 0, "residual: (N-1)^2 must be a multiple of 4 (N odd)");
  const int nprof = (2 + d.P) * (d.M + 1) + d.M1;
  const long long sV = bb.count > 1 ? bb.sV : 0;
  wg_profile_deriv_kernel<<<bb.count, 256, 0, st>>>(d.M, d.P, s.DchiFull, prof, derivs, nprof);
  WG_LAUNCH_CHECK();
  wg_rows_kernel<<<dim3(wg_cdiv(d.n, 256), bb.count), 256, 0, st>>>(d, prof, derivs, s.pz, s.pp, s.dpzdrz, s.stat, vw,
                                                                   collMult, 1.0, w1, w2, cT2, Ssrc, bb.vwArr, nprof, sV);
  WG_LAUNCH_CHECK();
  constexpr int R = 4;
  const size_t smem = ((size_t)R * d.q * (2 + d.P) + 2 * d.M1) * sizeof(double);
  WG_REQUIRE(smem <= 200 * 1024, "residual: basis tile does not fit in shared memory");
  if (wg_func_configured(reinterpret_cast<const void*>(wg_residual_kernel<R, 4>), smem) < (long long)smem) {
    WG_CUDA(cudaFuncSetAttribute(wg_residual_kernel<R, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WG_CUDA(cudaFuncSetAttribute(wg_residual_kernel<R, 4>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
  }
  dim3 grid(d.q / R, d.P * d.M1, bb.count);
  int threads = ((d.q / 4 + 31) / 32) * 32;
  if (threads > 512) threads = 512;
  if (threads < 64) threads = 64;
  wg_residual_kernel<R, 4><<<grid, threads, smem, st>>>(d, s.K1, s.K2, s.Coll, s.Tchi, s.Dchi, w1, w2, cT2, x, Ssrc, r, sV);
  WG_LAUNCH_CHECK();
  return WG_OK;
}

int bz_refine_update(int n, double* x, const double* dlt, double* stats, cudaStream_t st, int batch, long long sV) {
  wg_refine_update_kernel<<<batch, 1024, 0, st>>>(n, x, dlt, stats, batch > 1 ? sV : 0);
  WG_LAUNCH_CHECK();
  return WG_OK;
}

int bz_apply_axis(const double* in, double* out, const double* Mat, int outer, int len, int inner,
                  cudaStream_t st) {
  const long long total = (long long)outer * len * inner;
  int blocks = wg_cdiv(total, 256);
  if (blocks > 2368) blocks = 2368;
  wg_apply_axis_kernel<<<blocks, 256, 0, st>>>(in, out, Mat, outer, len, inner);
  WG_LAUNCH_CHECK();
  return WG_OK;
}

int bz_spectra(const BzDims& d, const double* c, double* spectra, cudaStream_t st, int batch) {
  wg_spectra_kernel<<<dim3(d.M1 + 2 * d.N1, batch), 256, 0, st>>>(d, c, spectra);
  WG_LAUNCH_CHECK();
  return WG_OK;
}

int bz_truncate(const BzDims& d, double* c, int keepM, int keepN1, int keepN2, double* err, cudaStream_t st, int batch,
                const int* keepArr) {
  wg_truncate_kernel<<<dim3(d.P, batch), 256, 0, st>>>(d, c, keepM, keepN1, keepN2, err, keepArr);
  WG_LAUNCH_CHECK();
  return WG_OK;
}

int bz_moments(const BzDims& d, const BzStatic& s, const double* card, const double* prof, double* deltas,
               cudaStream_t st, int batch) {
  const int nprof = (2 + d.P) * (d.M + 1) + d.M1;
  wg_moments_kernel<<<dim3(d.P * d.M1, batch), 128, 0, st>>>(d, card, prof, s.rz, s.rp, s.pz, s.pp, s.dpzdrz, s.dppdrp,
                                                            deltas, nprof);
  WG_LAUNCH_CHECK();
  return WG_OK;
}

int bz_collision_matvec(const BzDims& d, const BzStatic& s, const double* cT2, const double* x, double* y,
                        cudaStream_t st) {
  wg_collision_matvec_kernel<<<wg_cdiv((long long)d.n * 32, 256), 256, 0, st>>>(d, s.Coll, s.Tchi, cT2, x, y);
  WG_LAUNCH_CHECK();
  return WG_OK;
}

}  // namespace wg
