/*
 * wallgo_b200 -- C ABI of the B200-native Boltzmann hot path.
 *
 * The reference (Wall-Go/WallGo v1.1.2) is pure Python and has no FFI; the seam this
 * library plugs into is the class WallGo.BoltzmannSolver (src/WallGo/boltzmann.py:37-886,
 * constructed at src/WallGo/manager.py:605-619, consumed at
 * src/WallGo/equationOfMotion.py:1349-1355).  Each entry point below replaces one piece of
 * that class; the ctypes binding a maintainer adds is shown in INTEGRATION.md.
 *
 * Conventions
 *   - return value: 0 = ok, <0 = error (WG_ERR_*; text via wg_last_error()).  A singular pivot is
 *     reported LAPACK-style through wg_info() (>0 = 1-based column), mirroring numpy.linalg.LinAlgError
 *     raised from boltzmann.py:283.
 *   - "host" pointers are read synchronously during the call; "dev" pointers are CUDA device memory
 *     owned by the caller (e.g. torch tensors) and are used stream-ordered on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - all arrays are C-ordered float64.  n = P*(M-1)*(N-1)^2, q = (N-1)^2.
 *     row r = ((a*(M-1)+alpha)*(N-1)+beta)*(N-1)+gamma, col c = ((b*(M-1)+i)*(N-1)+j)*(N-1)+k
 *     (boltzmann.py:813-817, :286-292).
 *   - handles are not thread-safe (the reference is single-threaded); any number of solvers per process and
 *     per GPU, each used by one host thread at a time.
 *   - there is no CPU path: every call fails with WG_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef WALLGO_B200_H
#define WALLGO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define WG_OK 0
#define WG_ERR_ARG (-1)
#define WG_ERR_CUDA (-2)
#define WG_ERR_STATE (-3)
#define WG_ERR_UNSUPPORTED (-4)

typedef struct wg_solver wg_solver;
typedef struct wg_lu wg_lu;
typedef struct wg_batch wg_batch;

int wg_version(void);
const char* wg_last_error(void);
/* number of CUDA kernels this library has launched so far in this process (graph replays included) */
long long wg_launch_count(void);
/* Scheduling knobs of the LU.  The four setters below change the process-wide DEFAULTS: a solver adopts them the
 * next time it factors (its captured CUDA graphs are re-captured), unless it has been tuned individually with
 * wg_solver_set_tuning, which touches that solver only.
 * nb: top-level LU panel width (multiple of 16, 32..256); use_graph: replay the factorisation as a CUDA graph */
int wg_set_tuning(int nb, int use_graph);
/* on: factor panel k+1 on a high-priority stream while the trailing update of panel k runs (default 1) */
int wg_set_lookahead(int on);
/* rest: once the trailing matrix has at most this many rows the bulk of each LU step also runs on a
 * high-priority stream, so a second system sharing the GPU cannot starve this one's latency-bound tail
 * (default 8192; 0 disables) */
int wg_set_bulk_priority_rest(int rest);
/* percent: while the trailing matrix is large (n >= 8192) the bulk columns right of this percentage of n are
 * updated by a second, low-priority stream that runs one step behind the first, so the latency-bound
 * interchange + TRSM of one part overlaps the other's GEMM (default 50; 0 disables, 1 = default).  Pure
 * scheduling: results are bit-identical for every value. */
int wg_set_bulk_split(int percent);

/* The same knobs for ONE solver (negative value: keep); later changes of the defaults no longer affect it. */
int wg_solver_set_tuning(wg_solver* s, int nb, int use_graph, int lookahead, int split_percent, int bulk_hi_rest);

/* ---- lifecycle : BoltzmannSolver.__init__ + updateParticleList (boltzmann.py:52-139).  Every entry point
 * makes the solver's device current for the duration of the call and restores the caller's device, so solvers
 * on different GPUs can coexist in one process (one host thread per solver at a time). */
int wg_create(int P, int M, int N, int device, wg_solver** out);
int wg_destroy(wg_solver* s);
int wg_dims(const wg_solver* s, int* n, int* q);

/* ---- static data (host pointers)
 * grid: Grid.getCompactCoordinates / getCoordinates / getCompactificationDerivatives (grid.py:116-122, 219-274);
 *       chi[M-1], rz[N-1], rp[N-1], pz[N-1], pp[N-1], dpzdrz[N-1], dppdrp[N-1]                           */
int wg_set_grid(wg_solver* s, const double* chi, const double* rz, const double* rp, const double* pz,
                const double* pp, const double* dpzdrz, const double* dppdrp);
/* statistics[P]: +1 boson, -1 fermion (boltzmann.py:668-670) */
int wg_set_statistics(wg_solver* s, const double* statistics);
/* Evaluate the default-basis tables ON THE DEVICE (Cardinal in z, restricted Chebyshev in p_z,p_par,
 * spectral derivatives): Polynomial.matrix / derivMatrix (polynomial.py:635-862) and the two Kronecker
 * products T_rz(x)T_rp, D_rz(x)T_rp consumed by the assembly. */
int wg_build_basis(wg_solver* s);
/* Explicit tables for the other basis / derivative choices (boltzmann.py:673-726):
 * Tchi,Dchi [(M-1)^2], DchiFull [(M+1)^2], Trz,Drz,Trp [(N-1)^2] */
int wg_set_tables(wg_solver* s, const double* Tchi, const double* Dchi, const double* DchiFull,
                  const double* Trz, const double* Drz, const double* Trp);
int wg_get_tables(wg_solver* s, double* Tchi, double* Dchi, double* DchiFull, double* Trz, double* Drz,
                  double* Trp);
/* Basis-change matrices used by getDeltas: mats[3*stage+axis], stage 0 = solver basis -> Chebyshev,
 * 1 = Chebyshev -> solver basis, 2 = solver basis -> Cardinal; axis 0 = z [(M-1)^2], 1 = pz, 2 = pp
 * [(N-1)^2].  NULL = identity (Polynomial.changeBasis, polynomial.py:222-298). */
int wg_set_transforms(wg_solver* s, const double* const mats[9]);
/* Collision tensor C[a,beta,gamma,b,j,k] in the solver's momentum basis, P^2 q^2 doubles
 * (BoltzmannSolver.setCollisionArray, boltzmann.py:125-129).  on_device: pointer is device memory. */
int wg_set_collision(wg_solver* s, const double* C, int on_device);

/* Collision-file ingest, N_f -> N down-interpolation (CollisionArray.interpolateCollisionArray,
 * collisionArray.py:391-478) on the device.  src: tensor read from the files, [P][NF-1][NF-1][P][NF-1][NF-1],
 * Chebyshev basis in the polynomial axes; Lz, Lp [(NT-1) x (NF-1)]: Lagrange cardinals of the source grid's
 * rho_z / rho_par nodes evaluated on the target grid's nodes; out [P][NT-1][NT-1][P][NT-1][NT-1] (Chebyshev),
 * laid out as the reference lays it out.  All four are HOST pointers (once per run). */
int wg_interpolate_collision(int device, int P, int NF, int NT, const double* Lz, const double* Lp,
                             const double* src, double* out);

/* ---- per solve (device pointers, stream ordered)
 * profiles_dev = [ T(M+1) | v(M+1) plasma frame | msq(P x (M+1)) | dxi/dchi(M-1) ]
 * (BoltzmannSolver.setBackground + buildLinearEquations, boltzmann.py:116-123, 628-820).
 * Writes the dense operator A (n x n, leading dimension lda >= n, even) and the source S (n).
 * parts: bit 0 = Liouville term, bit 1 = collision term (3 = the operator; 1 and 2 give the
 * `liouville` and `collision` arrays that buildLinearEquations also returns, boltzmann.py:820). */
int wg_assemble(wg_solver* s, const double* profiles_dev, double vw, double collisionMultiplier, int parts,
                double* A_dev, long long lda, double* S_dev, void* stream);
/* Row-panel form of wg_assemble for operators that do not fit in memory (BASELINE configs[4]: P=4, M=64, N=31 is
 * 411 GB): writes the rows of the (species, position) groups [group0, group0+ngroups) -- q = (N-1)^2 rows each,
 * group g = a*(M-1)+alpha -- into panel_dev (ngroups*q x lda), exactly the bytes wg_assemble writes there; the
 * source S_dev (n) is written in full by every call.  A consumer streams the operator panel by panel. */
int wg_assemble_rows(wg_solver* s, const double* profiles_dev, double vw, double collisionMultiplier, int parts,
                     int group0, int ngroups, double* panel_dev, long long lda, double* S_dev, void* stream);
/* np.linalg.solve(operator, source) (boltzmann.py:283): LU in place with row partial pivoting ... */
int wg_factor(wg_solver* s, double* A_dev, long long lda, void* stream);
/* Both in one call, the way np.linalg.solve is used: the forward solve L y = P b rides along inside the
 * factorisation (on a stream that is off its critical path), only the backward solve remains afterwards.
 * b_dev (n) is overwritten by the solution; the factors stay valid for later wg_solve calls.  Lower latency
 * for one system alone (-0.6 ms at n = 15600); with several systems sharing the GPU the two separate calls
 * below give ~1.7 % more throughput, which is what the Python class uses. */
int wg_factor_solve(wg_solver* s, double* A_dev, long long lda, double* b_dev, void* stream);
/* ... and the triangular solves; b_dev (n) is overwritten by the solution.  May be called again with
 * another right-hand side (checkLinearization re-uses the factors, boltzmann.py:541-550). */
int wg_solve(wg_solver* s, const double* LU_dev, long long lda, double* b_dev, void* stream);
/* One step of mixed-precision iterative refinement of x_dev = deltaF: the residual
 * r = S - A x of the system of (profiles_dev, vw, collisionMultiplier) is formed with the operator entries
 * regenerated exactly as wg_assemble forms them (A is not needed in memory: LU_dev holds its factors) and
 * accumulated in double-double arithmetic; LU_dev solves for the correction, which is added to x_dev.  Removes the
 * cond(A)*eps forward error of the FP64 solve (np.linalg.solve, boltzmann.py:283, carries the same error).
 * stats_dev[2] = { max|correction|, max|x| }: their ratio is the forward error the first solve had; divided by
 * 2^-53 it estimates the condition number the solve experienced. */
int wg_refine(wg_solver* s, const double* profiles_dev, double vw, double collisionMultiplier, const double* LU_dev,
              long long lda, double* x_dev, double* stats_dev, void* stream);
/* Batched form for scans / many small systems: B independent systems (one solver handle, operator buffer,
 * source vector and stream each) are assembled, factored and solved; call b is enqueued on streams[b], so
 * the systems overlap on the GPU (a single LU cannot fill a B200 during its panel chain).  Equivalent to B
 * times wg_assemble(parts = 3) + wg_factor + wg_solve; S_dev[b] holds deltaF of system b afterwards.
 * Returns the first non-zero code (systems before it are enqueued). */
int wg_assemble_factor_solve_batched(int B, wg_solver* const* s, const double* const* profiles_dev,
                                     const double* vw, const double* collisionMultiplier, double* const* A_dev,
                                     const long long* lda, double* const* S_dev, void* const* streams);
/* synchronises `stream` and returns LAPACK-style info of the last wg_factor */
int wg_info(wg_solver* s, void* stream, int* info_host);
/* checkSpectralConvergence, first half (boltzmann.py:376-397): Chebyshev coefficients (kept inside the
 * solver) and spectra_dev[(M-1)+2(N-1)] = sum|c| along (chi | pz | pp). */
int wg_spectra(wg_solver* s, const double* deltaF_dev, double* spectra_dev, void* stream);
/* second half + estimateTruncationError + getDeltas quadrature (boltzmann.py:296-351, 431-491, 182-238).
 * Consumes the Chebyshev coefficients the preceding wg_spectra(deltaF_dev) left inside the solver: without that
 * call on the same deltaF_dev (or after another wg_moments) it returns WG_ERR_STATE.
 * keep* = number of coefficients kept per axis (0 zeroes the axis); roundtrip = 0 returns the input deltaF unchanged
 * (ETruncationOption.NONE).  Outputs: deltaF_out_dev[n], deltas_dev[4*P*(M-1)] (Delta00,02,20,11),
 * err_dev[4*P] (sum|c| kept, and on the last kept chi/pz/pp slice, per species). */
int wg_moments(wg_solver* s, const double* deltaF_dev, const double* profiles_dev, int keepM, int keepPz,
               int keepPp, int roundtrip, double* deltaF_out_dev, double* deltas_dev, double* err_dev,
               void* stream);
/* totalDOFs of each species (host, P doubles), used by wg_feedback */
int wg_set_dofs(wg_solver* s, const double* dofs);
/* Feedback of the moments into the wall equation of motion, on the device:
 *   out_dev = [ T30_out (M-1) | T33_out (M-1) | potentialOut (M-1) | dVout ((M-1) x nFields) ]
 * T30/T33: EOM.deltaToTmunu (equationOfMotion.py:2021-2127) for every interior z at once;
 * potentialOut = 1/2 sum_a dof_a m2_a Delta00_a (:1545-1555); dVout = 1/2 sum_a dof_a dm2_a/dphi Delta00_a
 * (:1418-1429).  deltas_dev as written by wg_moments; dmsq_dphi_dev[P][M-1][nFields] = Particle.msqDerivative
 * on the interior points (may be NULL when nFields == 0). */
int wg_feedback(wg_solver* s, const double* deltas_dev, const double* profiles_dev, const double* dmsq_dphi_dev,
                int nFields, double velocityMid, double* out_dev, void* stream);
/* y = (collision part of the last assembled operator) * x   (boltzmann.py:541-545) */
int wg_collision_apply(wg_solver* s, const double* x_dev, double* y_dev, void* stream);
/* basis change of a (P, M-1, N-1, N-1) array with the matrices of wg_set_transforms (stage 0..2);
 * out-of-place: in_dev and out_dev must differ */
int wg_change_basis(wg_solver* s, int stage, const double* in_dev, double* out_dev, void* stream);

/* Live timing of the dominant kernel (trailing-update GEMM of wg_factor) with CUDA events on the
 * launching stream.  enable != 0 switches bracketing on for later wg_factor calls; every call returns
 * and resets the sums of the launches bracketed so far (synchronise the stream first). */
int wg_profile_gemm(wg_solver* s, int enable, double* ms, double* flops, int* count);

/* ---- strided batches: many small systems per launch (north star (2): "a batched FP64 blocked LU";
 * BASELINE configs[0], n = 1900).  A wg_batch belongs to one solver (shares its grid, tables, statistics and
 * collision tensor, which must outlive it) and owns the scratch of `count` independent systems of that shape.
 * Every kernel of the path carries the system index as a grid dimension, so ONE launch sequence (and one
 * captured CUDA graph) assembles, factors, solves and reduces all `count` systems in lock step -- no per-system
 * stream or graph -- and each system's result is bit-identical to the single-system entry points above.
 * Layout: operators A_dev + b*strideA (strideA >= n*lda doubles, a multiple of 4), everything else contiguous
 * per system: profiles_dev[count][(2+P)(M+1)+(M-1)], vw_dev[count], S_dev / deltaF[count][n],
 * spectra_dev[count][(M-1)+2(N-1)], keep_dev[count][3] (int: keepM, keepPz, keepPp), deltas_dev[count][4 P (M-1)],
 * err_dev[count][4 P], info_host[count]. */
int wg_batch_create(wg_solver* s, int count, wg_batch** out);
int wg_batch_destroy(wg_batch* b);
/* wg_assemble for every system of the batch; the wall velocities come from device memory */
int wg_batch_assemble(wg_batch* b, const double* profiles_dev, const double* vw_dev, double collisionMultiplier,
                      int parts, double* A_dev, long long lda, long long strideA, double* S_dev, void* stream);
int wg_batch_factor(wg_batch* b, double* A_dev, long long lda, long long strideA, void* stream);
int wg_batch_solve(wg_batch* b, const double* LU_dev, long long lda, long long strideA, double* rhs_dev,
                   void* stream);
/* wg_refine for every system of the batch; stats_dev[count][2] */
int wg_batch_refine(wg_batch* b, const double* profiles_dev, const double* vw_dev, double collisionMultiplier,
                    const double* LU_dev, long long lda, long long strideA, double* x_dev, double* stats_dev,
                    void* stream);
int wg_batch_info(wg_batch* b, void* stream, int* info_host);
int wg_batch_spectra(wg_batch* b, const double* deltaF_dev, double* spectra_dev, void* stream);
int wg_batch_moments(wg_batch* b, const double* deltaF_dev, const double* profiles_dev, const int* keep_dev,
                     int roundtrip, double* deltaF_out_dev, double* deltas_dev, double* err_dev, void* stream);

/* ---- generic dense pieces (used by the solver; exported for tests and other callers) */
int wg_lu_create(int n, int device, wg_lu** out);
/* `batch` systems of order n factored / solved in lock step by every launch: operators at A_dev + b*strideA,
 * pivots ipiv_dev[batch][n], info_dev[batch], right-hand sides at b_dev + b*strideB */
int wg_lu_create_batched(int n, int batch, int device, wg_lu** out);
int wg_lu_factor_batched(wg_lu* p, double* A_dev, long long lda, long long strideA, int* ipiv_dev, int* info_dev,
                         void* stream);
int wg_lu_solve_batched(wg_lu* p, const double* LU_dev, long long lda, long long strideA, double* b_dev,
                        long long strideB, void* stream);
int wg_lu_destroy(wg_lu* p);
int wg_lu_factor(wg_lu* p, double* A_dev, long long lda, int* ipiv_dev, int* info_dev, void* stream);
int wg_lu_solve(wg_lu* p, const double* LU_dev, long long lda, double* b_dev, void* stream);
/* C[MxN] -= A[MxK] * B[KxN], row-major, FP64 tensor cores */
int wg_dgemm_sub(double* C_dev, const double* A_dev, const double* B_dev, int M, int N, int K,
                 long long ldc, long long lda, long long ldb, void* stream);

#ifdef __cplusplus
}
#endif
#endif
