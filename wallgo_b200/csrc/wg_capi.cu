// extern "C" boundary of libwallgo_b200.so (see include/wallgo_b200.h).
#include <limits.h>
#include <stdlib.h>
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/wallgo_b200.h"
#include "wg_boltzmann.h"
#include "wg_common.cuh"
#include "wg_kernels.h"

static thread_local char g_err[512] = "";
std::atomic<long long> g_wg_launches{0};
thread_local int g_wg_launch_prio = INT_MIN;
int g_wg_sync_every_launch = (getenv("WG_SYNC_EVERY_LAUNCH") != nullptr);
long long wg_func_configured(const void* func, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, long long> seen;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(dev, func);
  auto it = seen.find(key);
  const long long before = (it == seen.end()) ? -1 : it->second;
  if ((long long)smem > before) seen[key] = (long long)smem;
  return before;
}

void wg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Every entry point runs with the handle's device current and restores the caller's device on return, so
// engines on different GPUs can live in one process and the host's own torch code keeps its device.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) ok = (cudaSetDevice(device) == cudaSuccess);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
#define WG_ON_DEVICE(dev)                                          \
  DeviceGuard _guard(dev);                                         \
  if (!_guard.ok) {                                                \
    wg_set_error("cudaSetDevice(%d) failed", (int)(dev));          \
    return WG_ERR_CUDA;                                            \
  }

struct wg_lu {
  wg::LuPlan* plan = nullptr;
  int device = 0;
};

struct wg_solver {
  wg::BzDims d;
  int device = 0;
  wg::BzStatic st = {};
  std::vector<void*> allocs;
  double* mats[9] = {};  // transforms (device) or nullptr
  double* matsSpare[9] = {};  // buffers of transforms that are currently the identity
  // per-solve scratch
  double *derivs = nullptr, *w1 = nullptr, *w2 = nullptr, *cT2 = nullptr;
  double *cheb = nullptr, *bufA = nullptr, *bufB = nullptr;
  int *ipiv = nullptr, *info = nullptr;
  double* dofs = nullptr;
  wg::LuPlan* plan = nullptr;
  bool haveGrid = false, haveTables = false, haveColl = false, haveStat = false, assembled = false;
  const double* chebOf = nullptr;  // deltaF whose Chebyshev coefficients wg_spectra left in `cheb` (nullptr: none)
};

// Strided batch of `count` independent systems of one solver's shape, advanced in lock step by every kernel launch
// (the system index is a grid dimension): shares the solver's static tables, owns the per-system scratch.
struct wg_batch {
  wg_solver* s = nullptr;
  int device = 0;      // (kept here too: wg_batch_destroy must not look into a solver that may already be gone)
  int count = 0;
  std::vector<void*> allocs;
  double *derivs = nullptr, *w1 = nullptr, *w2 = nullptr, *cT2 = nullptr;
  double *cheb = nullptr, *bufA = nullptr, *bufB = nullptr;
  int *ipiv = nullptr, *info = nullptr;
  wg::LuPlan* plan = nullptr;
  const double* chebOf = nullptr;
};

template <typename T>
static int dev_alloc(wg_solver* s, T** p, size_t count) {
  WG_CUDA(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  s->allocs.push_back(*p);
  return WG_OK;
}

static int require_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    wg_set_error("no CUDA device available (%s); wallgo_b200 has no CPU path",
                 e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return WG_ERR_CUDA;
  }
  WG_REQUIRE(device >= 0 && device < count, "device index out of range");
  cudaDeviceProp prop;
  WG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    wg_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                 prop.minor);
    return WG_ERR_UNSUPPORTED;
  }
  return WG_OK;
}

extern "C" {

int wg_version(void) { return 100; }
long long wg_launch_count(void) { return g_wg_launches.load(); }
const char* wg_last_error(void) { return g_err; }
int wg_set_tuning(int nb, int use_graph) {
  wg::lu_set_tuning(nb, use_graph);
  return WG_OK;
}
int wg_debug_gemm_variant(int v) {
  wg::gemm_set_variant(v);
  return WG_OK;
}
int wg_profile_gemm(wg_solver* s, int enable, double* ms, double* flops, int* count) {
  WG_REQUIRE(s, "wg_profile_gemm: NULL solver");
  WG_ON_DEVICE(s->device);
  wg::lu_set_profile(enable);
  return wg::lu_profile_read(s->plan, ms, flops, count);
}
int wg_set_bulk_priority_rest(int rest) {
  wg::lu_set_bulk_hi(rest);
  return WG_OK;
}
/* developer knob (not part of the public header): 32-column leaf panels on/off */
int wg_debug_leaf32(int on) {
  wg::lu_set_leaf32(on);
  return WG_OK;
}
int wg_set_bulk_split(int percent) {
  wg::lu_set_split(percent);
  return WG_OK;
}
/* developer knob (not part of the public header): 32-byte stores in the assembly kernel on/off */
int wg_debug_asm_wide(int on) {
  wg::bz_set_asm_wide(on);
  return WG_OK;
}
/* developer probe (not part of the public header; built with -DWG_LEAF_CHECKS): first bounds violation in a leaf */
int wg_debug_leaf_fail(int* out8) { return wg::lu_leaf_fail_read(out8); }
/* developer knob (not part of the public header): force the leaf width class of the LU panels */
int wg_debug_force_wb(int wb) {
  wg::lu_set_force_wb(wb);
  return WG_OK;
}
int wg_solver_set_tuning(wg_solver* s, int nb, int use_graph, int lookahead, int split_percent, int bulk_hi_rest) {
  WG_REQUIRE(s && s->plan, "wg_solver_set_tuning: NULL solver");
  wg::lu_plan_set_tuning(s->plan, nb, use_graph, lookahead, split_percent, bulk_hi_rest);
  return WG_OK;
}
/* developer knob (not part of the public header): triangular solves as two persistent wavefront launches (1,
 * default) or one launch per 128-row block (0); same bits either way */
int wg_debug_wave_solve(int on) {
  wg::lu_set_wave_solve(on);
  return WG_OK;
}
int wg_set_lookahead(int on) {
  wg::lu_set_lookahead(on);
  return WG_OK;
}

/* developer probe (not part of the public header): clock64 phase stamps of one leaf panel */
int wg_debug_leaf(long long* dev_buf, int c0) {
  wg::lu_set_debug(dev_buf, c0);
  return WG_OK;
}

int wg_create(int P, int M, int N, int device, wg_solver** out) {
  WG_REQUIRE(out != nullptr, "wg_create: out is NULL");
  WG_REQUIRE(P >= 1 && M >= 3 && N >= 3 && (N % 2 == 1), "wg_create: need P>=1, M>=3, N>=3 and N odd");
  int rc = require_device(device);
  if (rc) return rc;
  WG_ON_DEVICE(device);
  wg_solver* s = new wg_solver();
  s->device = device;
  wg::BzDims& d = s->d;
  d.P = P;
  d.M = M;
  d.N = N;
  d.M1 = M - 1;
  d.N1 = N - 1;
  d.q = d.N1 * d.N1;
  d.n = P * d.M1 * d.q;
  const size_t M1 = d.M1, N1 = d.N1, q = d.q, n = d.n;
#define A_(ptr, cnt)                      \
  rc = dev_alloc(s, &(ptr), (cnt));       \
  if (rc) {                               \
    wg_destroy(s);                        \
    return rc;                            \
  }
  A_(s->st.chi, M1) A_(s->st.rz, N1) A_(s->st.rp, N1) A_(s->st.pz, N1) A_(s->st.pp, N1)
  A_(s->st.dpzdrz, N1) A_(s->st.dppdrp, N1) A_(s->st.stat, (size_t)P)
  A_(s->st.Tchi, M1 * M1) A_(s->st.Dchi, M1 * M1) A_(s->st.DchiFull, (size_t)(M + 1) * (M + 1))
  A_(s->st.Trz, N1 * N1) A_(s->st.Drz, N1 * N1) A_(s->st.Trp, N1 * N1)
  A_(s->st.K1, q * q) A_(s->st.K2, q * q) A_(s->st.Coll, (size_t)P * q * P * q)
  A_(s->derivs, (size_t)(2 + P) * (M + 1)) A_(s->w1, n) A_(s->w2, n) A_(s->cT2, M1)
  A_(s->cheb, n) A_(s->bufA, n) A_(s->bufB, n) A_(s->ipiv, n) A_(s->info, (size_t)1) A_(s->dofs, (size_t)P)
#undef A_
  WG_CUDA(cudaMemset(s->info, 0, sizeof(int)));
  rc = wg::lu_plan_create(d.n, &s->plan);
  if (rc) {
    wg_destroy(s);
    return rc;
  }
  *out = s;
  return WG_OK;
}

int wg_destroy(wg_solver* s) {
  if (!s) return WG_OK;
  WG_ON_DEVICE(s->device);
  for (void* p : s->allocs) cudaFree(p);
  for (int i = 0; i < 9; ++i) {
    if (s->mats[i]) cudaFree(s->mats[i]);
    if (s->matsSpare[i]) cudaFree(s->matsSpare[i]);
  }
  wg::lu_plan_destroy(s->plan);
  delete s;
  return WG_OK;
}

int wg_dims(const wg_solver* s, int* n, int* q) {
  WG_REQUIRE(s, "wg_dims: NULL solver");
  if (n) *n = s->d.n;
  if (q) *q = s->d.q;
  return WG_OK;
}

static int h2d(double* dst, const double* src, size_t count) {
  WG_REQUIRE(src != nullptr, "NULL host pointer");
  WG_CUDA(cudaMemcpy(dst, src, count * sizeof(double), cudaMemcpyHostToDevice));
  return WG_OK;
}

int wg_set_grid(wg_solver* s, const double* chi, const double* rz, const double* rp, const double* pz,
                const double* pp, const double* dpzdrz, const double* dppdrp) {
  WG_REQUIRE(s, "wg_set_grid: NULL solver");
  WG_ON_DEVICE(s->device);
  int rc = 0;
  rc |= h2d(s->st.chi, chi, s->d.M1);
  rc |= h2d(s->st.rz, rz, s->d.N1);
  rc |= h2d(s->st.rp, rp, s->d.N1);
  rc |= h2d(s->st.pz, pz, s->d.N1);
  rc |= h2d(s->st.pp, pp, s->d.N1);
  rc |= h2d(s->st.dpzdrz, dpzdrz, s->d.N1);
  rc |= h2d(s->st.dppdrp, dppdrp, s->d.N1);
  if (rc) return rc < 0 ? rc : WG_ERR_CUDA;
  s->haveGrid = true;
  return WG_OK;
}

int wg_set_statistics(wg_solver* s, const double* statistics) {
  WG_REQUIRE(s, "wg_set_statistics: NULL solver");
  WG_ON_DEVICE(s->device);
  int rc = h2d(s->st.stat, statistics, s->d.P);
  if (rc) return rc;
  s->haveStat = true;
  return WG_OK;
}

int wg_set_dofs(wg_solver* s, const double* dofs) {
  WG_REQUIRE(s, "wg_set_dofs: NULL solver");
  WG_ON_DEVICE(s->device);
  return h2d(s->dofs, dofs, s->d.P);
}

int wg_feedback(wg_solver* s, const double* deltas_dev, const double* profiles_dev, const double* dmsq_dphi_dev,
                int nFields, double velocityMid, double* out_dev, void* stream) {
  WG_REQUIRE(s && deltas_dev && profiles_dev && out_dev && nFields >= 0 && (nFields == 0 || dmsq_dphi_dev),
             "wg_feedback: bad argument");
  WG_ON_DEVICE(s->device);
  WG_REQUIRE(velocityMid > -1.0 && velocityMid < 1.0, "wg_feedback: |velocityMid| must be < 1");
  const double u0 = sqrt(1.0 / (1.0 - velocityMid * velocityMid));  // helpers.gammaSq
  const double u3 = u0 * velocityMid;
  return wg::bz_feedback(s->d, deltas_dev, profiles_dev, s->dofs, dmsq_dphi_dev, nFields, u0, u3, out_dev,
                         (cudaStream_t)stream);
}

int wg_build_basis(wg_solver* s) {
  WG_REQUIRE(s && s->haveGrid, "wg_build_basis: call wg_set_grid first");
  WG_ON_DEVICE(s->device);
  int rc = wg::bz_build_basis(s->d, s->st.rz, s->st.Tchi, s->st.Dchi, s->st.DchiFull, s->st.Trz, s->st.Drz,
                              s->st.Trp, 0);
  if (rc) return rc;
  rc = wg::bz_build_kron(s->d, s->st.Trz, s->st.Drz, s->st.Trp, s->st.K1, s->st.K2, 0);
  if (rc) return rc;
  WG_CUDA(cudaDeviceSynchronize());
  s->haveTables = true;
  return WG_OK;
}

int wg_set_tables(wg_solver* s, const double* Tchi, const double* Dchi, const double* DchiFull,
                  const double* Trz, const double* Drz, const double* Trp) {
  WG_REQUIRE(s, "wg_set_tables: NULL solver");
  WG_ON_DEVICE(s->device);
  const size_t M1 = s->d.M1, N1 = s->d.N1, L = s->d.M + 1;
  int rc = 0;
  rc |= h2d(s->st.Tchi, Tchi, M1 * M1);
  rc |= h2d(s->st.Dchi, Dchi, M1 * M1);
  rc |= h2d(s->st.DchiFull, DchiFull, L * L);
  rc |= h2d(s->st.Trz, Trz, N1 * N1);
  rc |= h2d(s->st.Drz, Drz, N1 * N1);
  rc |= h2d(s->st.Trp, Trp, N1 * N1);
  if (rc) return rc < 0 ? rc : WG_ERR_CUDA;
  rc = wg::bz_build_kron(s->d, s->st.Trz, s->st.Drz, s->st.Trp, s->st.K1, s->st.K2, 0);
  if (rc) return rc;
  WG_CUDA(cudaDeviceSynchronize());
  s->haveTables = true;
  return WG_OK;
}

int wg_get_tables(wg_solver* s, double* Tchi, double* Dchi, double* DchiFull, double* Trz, double* Drz,
                  double* Trp) {
  WG_REQUIRE(s && s->haveTables, "wg_get_tables: no tables yet");
  WG_ON_DEVICE(s->device);
  const size_t M1 = s->d.M1, N1 = s->d.N1, L = s->d.M + 1;
  if (Tchi) WG_CUDA(cudaMemcpy(Tchi, s->st.Tchi, M1 * M1 * 8, cudaMemcpyDeviceToHost));
  if (Dchi) WG_CUDA(cudaMemcpy(Dchi, s->st.Dchi, M1 * M1 * 8, cudaMemcpyDeviceToHost));
  if (DchiFull) WG_CUDA(cudaMemcpy(DchiFull, s->st.DchiFull, L * L * 8, cudaMemcpyDeviceToHost));
  if (Trz) WG_CUDA(cudaMemcpy(Trz, s->st.Trz, N1 * N1 * 8, cudaMemcpyDeviceToHost));
  if (Drz) WG_CUDA(cudaMemcpy(Drz, s->st.Drz, N1 * N1 * 8, cudaMemcpyDeviceToHost));
  if (Trp) WG_CUDA(cudaMemcpy(Trp, s->st.Trp, N1 * N1 * 8, cudaMemcpyDeviceToHost));
  return WG_OK;
}

int wg_set_transforms(wg_solver* s, const double* const mats[9]) {
  WG_REQUIRE(s && mats, "wg_set_transforms: NULL argument");
  WG_ON_DEVICE(s->device);
  for (int i = 0; i < 9; ++i) {
    const size_t len = (i % 3 == 0) ? s->d.M1 : s->d.N1;
    if (!mats[i]) {   // identity: the buffer (if any) is parked for a later call instead of being freed
      if (s->mats[i]) s->matsSpare[i] = s->mats[i];
      s->mats[i] = nullptr;
      continue;
    }
    if (!s->mats[i]) {
      s->mats[i] = s->matsSpare[i];
      s->matsSpare[i] = nullptr;
    }
    if (!s->mats[i]) WG_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->mats[i]), len * len * sizeof(double)));
    // ordered behind earlier work of the default stream only; callers synchronise their own stream before
    // replacing the matrices of a solver that still has basis changes in flight
    WG_CUDA(cudaMemcpy(s->mats[i], mats[i], len * len * sizeof(double), cudaMemcpyHostToDevice));
  }
  s->chebOf = nullptr;
  return WG_OK;
}

int wg_set_collision(wg_solver* s, const double* C, int on_device) {
  WG_REQUIRE(s && C, "wg_set_collision: NULL argument");
  WG_ON_DEVICE(s->device);
  const size_t cnt = (size_t)s->d.P * s->d.q * s->d.P * s->d.q;
  WG_CUDA(cudaMemcpy(s->st.Coll, C, cnt * sizeof(double),
                     on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  s->haveColl = true;
  return WG_OK;
}

int wg_assemble(wg_solver* s, const double* profiles_dev, double vw, double collisionMultiplier, int parts,
                double* A_dev, long long lda, double* S_dev, void* stream) {
  WG_REQUIRE(s && profiles_dev && A_dev && S_dev, "wg_assemble: NULL argument");
  WG_ON_DEVICE(s->device);
  if (!(s->haveGrid && s->haveTables && s->haveColl && s->haveStat)) {
    wg_set_error("wg_assemble: grid, statistics, tables and collision tensor must be set first");
    return WG_ERR_STATE;
  }
  WG_REQUIRE(lda >= s->d.n, "wg_assemble: lda < n");
  int rc = wg::bz_assemble(s->d, s->st, profiles_dev, vw, collisionMultiplier, parts, s->derivs, s->w1, s->w2,
                           s->cT2, A_dev, lda, S_dev, (cudaStream_t)stream);
  if (rc == WG_OK && (parts & 2)) s->assembled = true;
  return rc;
}

int wg_assemble_rows(wg_solver* s, const double* profiles_dev, double vw, double collisionMultiplier, int parts,
                     int group0, int ngroups, double* panel_dev, long long lda, double* S_dev, void* stream) {
  WG_REQUIRE(s && profiles_dev && panel_dev && S_dev, "wg_assemble_rows: NULL argument");
  WG_ON_DEVICE(s->device);
  if (!(s->haveGrid && s->haveTables && s->haveColl && s->haveStat)) {
    wg_set_error("wg_assemble_rows: grid, statistics, tables and collision tensor must be set first");
    return WG_ERR_STATE;
  }
  WG_REQUIRE(lda >= s->d.n, "wg_assemble_rows: lda < n");
  WG_REQUIRE(group0 >= 0 && ngroups >= 1 && group0 + ngroups <= s->d.P * s->d.M1, "wg_assemble_rows: group range out of bounds");
  return wg::bz_assemble(s->d, s->st, profiles_dev, vw, collisionMultiplier, parts, s->derivs, s->w1, s->w2, s->cT2,
                         panel_dev, lda, S_dev, (cudaStream_t)stream, wg::BzBatch(), group0, ngroups);
}

int wg_factor(wg_solver* s, double* A_dev, long long lda, void* stream) {
  WG_REQUIRE(s && A_dev, "wg_factor: NULL argument");
  WG_ON_DEVICE(s->device);
  return wg::lu_factor(s->plan, A_dev, lda, s->ipiv, s->info, (cudaStream_t)stream);
}

int wg_factor_solve(wg_solver* s, double* A_dev, long long lda, double* b_dev, void* stream) {
  WG_REQUIRE(s && A_dev && b_dev, "wg_factor_solve: NULL argument");
  WG_ON_DEVICE(s->device);
  int rc = wg::lu_factor(s->plan, A_dev, lda, s->ipiv, s->info, (cudaStream_t)stream, b_dev);
  if (rc) return rc;
  return wg::lu_solve(s->plan, A_dev, lda, b_dev, (cudaStream_t)stream, true);
}

int wg_solve(wg_solver* s, const double* LU_dev, long long lda, double* b_dev, void* stream) {
  WG_REQUIRE(s && LU_dev && b_dev, "wg_solve: NULL argument");
  WG_ON_DEVICE(s->device);
  return wg::lu_solve(s->plan, LU_dev, lda, b_dev, (cudaStream_t)stream);
}

int wg_refine(wg_solver* s, const double* profiles_dev, double vw, double collisionMultiplier, const double* LU_dev,
              long long lda, double* x_dev, double* stats_dev, void* stream) {
  WG_REQUIRE(s && profiles_dev && LU_dev && x_dev && stats_dev, "wg_refine: NULL argument");
  WG_ON_DEVICE(s->device);
  if (!(s->haveGrid && s->haveTables && s->haveColl && s->haveStat)) {
    wg_set_error("wg_refine: grid, statistics, tables and collision tensor must be set first");
    return WG_ERR_STATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = wg::bz_residual(s->d, s->st, profiles_dev, vw, collisionMultiplier, s->derivs, s->w1, s->w2, s->cT2, x_dev,
                           s->bufA, s->bufB, st);
  if (rc) return rc;
  rc = wg::lu_solve(s->plan, LU_dev, lda, s->bufB, st);
  if (rc) return rc;
  return wg::bz_refine_update(s->d.n, x_dev, s->bufB, stats_dev, st);
}

int wg_assemble_factor_solve_batched(int B, wg_solver* const* s, const double* const* profiles_dev,
                                     const double* vw, const double* collisionMultiplier, double* const* A_dev,
                                     const long long* lda, double* const* S_dev, void* const* streams) {
  WG_REQUIRE(B >= 0 && (B == 0 || (s && profiles_dev && vw && collisionMultiplier && A_dev && lda && S_dev && streams)),
             "wg_assemble_factor_solve_batched: NULL argument");
  for (int b = 0; b < B; ++b) {
    int rc = wg_assemble(s[b], profiles_dev[b], vw[b], collisionMultiplier[b], 3, A_dev[b], lda[b], S_dev[b], streams[b]);
    if (rc) return rc;
    rc = wg_factor(s[b], A_dev[b], lda[b], streams[b]);
    if (rc) return rc;
    rc = wg_solve(s[b], A_dev[b], lda[b], S_dev[b], streams[b]);
    if (rc) return rc;
  }
  return WG_OK;
}

int wg_info(wg_solver* s, void* stream, int* info_host) {
  WG_REQUIRE(s && info_host, "wg_info: NULL argument");
  WG_ON_DEVICE(s->device);
  WG_CUDA(cudaMemcpyAsync(info_host, s->info, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  WG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return WG_OK;
}

// apply the (up to three) axis matrices of one stage; result always lands in `out`
// (batch > 1: `in`/`out` hold `batch` contiguous arrays, tmpA/tmpB are scratch of the same size)
static int apply_stage(wg_solver* s, int stage, const double* in, double* out, cudaStream_t st, int batch = 1,
                       double* tmpA = nullptr, double* tmpB = nullptr) {
  const wg::BzDims& d = s->d;
  const double* const* m = s->mats + 3 * stage;
  int cnt = (m[0] != nullptr) + (m[1] != nullptr) + (m[2] != nullptr);
  if (cnt == 0) {
    if (in != out) WG_CUDA(cudaMemcpyAsync(out, in, sizeof(double) * d.n * batch, cudaMemcpyDeviceToDevice, st));
    return WG_OK;
  }
  const double* cur = in;
  double* tmp[2] = {tmpA ? tmpA : s->bufA, tmpB ? tmpB : s->bufB};
  int ti = 0;
  for (int ax = 0; ax < 3; ++ax) {
    if (!m[ax]) continue;
    --cnt;
    double* dst = (cnt == 0) ? out : tmp[ti];
    if (dst == cur) {  // never alias (only possible when out == in)
      wg_set_error("apply_stage: in-place basis change needs distinct buffers");
      return WG_ERR_ARG;
    }
    int rc;
    if (ax == 0)
      rc = wg::bz_apply_axis(cur, dst, m[ax], batch * d.P, d.M1, d.q, st);
    else if (ax == 1)
      rc = wg::bz_apply_axis(cur, dst, m[ax], batch * d.P * d.M1, d.N1, d.N1, st);
    else
      rc = wg::bz_apply_axis(cur, dst, m[ax], batch * d.P * d.M1 * d.N1, d.N1, 1, st);
    if (rc) return rc;
    cur = dst;
    ti ^= 1;
  }
  return WG_OK;
}

int wg_change_basis(wg_solver* s, int stage, const double* in_dev, double* out_dev, void* stream) {
  WG_REQUIRE(s && in_dev && out_dev && stage >= 0 && stage < 3 && in_dev != out_dev,
             "wg_change_basis: bad argument");
  WG_ON_DEVICE(s->device);
  return apply_stage(s, stage, in_dev, out_dev, (cudaStream_t)stream);
}

int wg_spectra(wg_solver* s, const double* deltaF_dev, double* spectra_dev, void* stream) {
  WG_REQUIRE(s && deltaF_dev && spectra_dev, "wg_spectra: NULL argument");
  WG_ON_DEVICE(s->device);
  cudaStream_t st = (cudaStream_t)stream;
  s->chebOf = nullptr;
  int rc = apply_stage(s, 0, deltaF_dev, s->cheb, st);
  if (rc) return rc;
  rc = wg::bz_spectra(s->d, s->cheb, spectra_dev, st);
  if (rc == WG_OK) s->chebOf = deltaF_dev;
  return rc;
}

int wg_moments(wg_solver* s, const double* deltaF_dev, const double* profiles_dev, int keepM, int keepPz,
               int keepPp, int roundtrip, double* deltaF_out_dev, double* deltas_dev, double* err_dev,
               void* stream) {
  WG_REQUIRE(s && deltaF_dev && profiles_dev && deltaF_out_dev && deltas_dev && err_dev,
             "wg_moments: NULL argument");
  WG_ON_DEVICE(s->device);
  // 0 is allowed: on grids with fewer than three coefficients along an axis the reference's truncation
  // zeroes the whole axis (boltzmann.py:400-446, the slice [-0:])
  WG_REQUIRE(keepM >= 0 && keepM <= s->d.M1 && keepPz >= 0 && keepPz <= s->d.N1 && keepPp >= 0 &&
                 keepPp <= s->d.N1,
             "wg_moments: keep counts out of range");
  WG_REQUIRE(s->haveGrid, "wg_moments: grid not set");
  cudaStream_t st = (cudaStream_t)stream;
  if (s->chebOf != deltaF_dev) {
    // the truncation works on the Chebyshev coefficients that wg_spectra leaves inside the solver; they are
    // consumed by this call (the buffer is re-used for the cardinal values)
    wg_set_error("wg_moments: call wg_spectra on the same deltaF_dev first (each wg_moments consumes one wg_spectra)");
    return WG_ERR_STATE;
  }
  s->chebOf = nullptr;
  int rc = wg::bz_truncate(s->d, s->cheb, keepM, keepPz, keepPp, err_dev, st);
  if (rc) return rc;
  if (roundtrip) {
    rc = apply_stage(s, 1, s->cheb, deltaF_out_dev, st);
    if (rc) return rc;
  } else if (deltaF_out_dev != deltaF_dev) {
    WG_CUDA(cudaMemcpyAsync(deltaF_out_dev, deltaF_dev, sizeof(double) * s->d.n, cudaMemcpyDeviceToDevice, st));
  }
  // cardinal-basis values; s->cheb is free again after the round trip, use it as the target
  double* card = s->cheb;
  rc = apply_stage(s, 2, deltaF_out_dev, card, st);
  if (rc) return rc;
  return wg::bz_moments(s->d, s->st, card, profiles_dev, deltas_dev, st);
}

int wg_collision_apply(wg_solver* s, const double* x_dev, double* y_dev, void* stream) {
  WG_REQUIRE(s && x_dev && y_dev, "wg_collision_apply: NULL argument");
  WG_ON_DEVICE(s->device);
  if (!s->assembled) {
    wg_set_error("wg_collision_apply: call wg_assemble first (needs T^2 of the current background)");
    return WG_ERR_STATE;
  }
  return wg::bz_collision_matvec(s->d, s->st, s->cT2, x_dev, y_dev, (cudaStream_t)stream);
}

int wg_interpolate_collision(int device, int P, int NF, int NT, const double* Lz, const double* Lp,
                             const double* src, double* out) {
  WG_REQUIRE(P >= 1 && NF >= 3 && NT >= 3 && NT <= NF && Lz && Lp && src && out, "wg_interpolate_collision: bad argument");
  int rc = require_device(device);
  if (rc) return rc;
  WG_ON_DEVICE(device);
  const size_t nF = NF - 1, nT = NT - 1;
  const size_t srcCount = (size_t)P * nF * nF * P * nF * nF, outCount = (size_t)P * nT * nT * P * nT * nT;
  double *dLz = nullptr, *dLp = nullptr, *dSrc = nullptr, *dOut = nullptr;
  auto release = [&]() {
    cudaFree(dLz);
    cudaFree(dLp);
    cudaFree(dSrc);
    cudaFree(dOut);
  };
  cudaError_t e = cudaMalloc(&dLz, nT * nF * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dLp, nT * nF * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dSrc, srcCount * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dOut, outCount * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(dLz, Lz, nT * nF * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dLp, Lp, nT * nF * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dSrc, src, srcCount * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    release();
    wg_set_error("wg_interpolate_collision: %s", cudaGetErrorString(e));
    return WG_ERR_CUDA;
  }
  rc = wg::bz_collision_interp(P, (int)nF, (int)nT, dLz, dLp, dSrc, dOut, 0);
  if (rc == WG_OK) {
    e = cudaMemcpy(out, dOut, outCount * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
      wg_set_error("wg_interpolate_collision: %s", cudaGetErrorString(e));
      rc = WG_ERR_CUDA;
    }
  }
  release();
  return rc;
}

/* ---- strided batches: B systems per launch (many small systems; north-star "batched LU") */
int wg_batch_create(wg_solver* s, int count, wg_batch** out) {
  WG_REQUIRE(s && out && count >= 1 && count <= 4096, "wg_batch_create: bad argument");
  WG_ON_DEVICE(s->device);
  wg_batch* b = new wg_batch();
  b->s = s;
  b->device = s->device;
  b->count = count;
  const wg::BzDims& d = s->d;
  const size_t B = (size_t)count, n = d.n;
  int rc = WG_OK;
  auto alloc = [&](void** p, size_t bytes) {
    if (rc) return;
    if (cudaMalloc(p, bytes) != cudaSuccess) {
      wg_set_error("wg_batch_create: out of device memory (%zu bytes)", bytes);
      (void)cudaGetLastError();
      rc = WG_ERR_CUDA;
      return;
    }
    b->allocs.push_back(*p);
  };
  alloc((void**)&b->derivs, 8 * B * (size_t)(2 + d.P) * (d.M + 1));
  alloc((void**)&b->w1, 8 * B * n);
  alloc((void**)&b->w2, 8 * B * n);
  alloc((void**)&b->cT2, 8 * B * d.M1);
  alloc((void**)&b->cheb, 8 * B * n);
  alloc((void**)&b->bufA, 8 * B * n);
  alloc((void**)&b->bufB, 8 * B * n);
  alloc((void**)&b->ipiv, 4 * B * n);
  alloc((void**)&b->info, 4 * B);
  if (!rc && cudaMemset(b->info, 0, 4 * B) != cudaSuccess) rc = WG_ERR_CUDA;
  if (!rc) rc = wg::lu_plan_create(d.n, &b->plan, count);
  if (rc) {
    wg_batch_destroy(b);
    return rc;
  }
  *out = b;
  return WG_OK;
}

int wg_batch_destroy(wg_batch* b) {
  if (!b) return WG_OK;
  WG_ON_DEVICE(b->device);
  for (void* p : b->allocs) cudaFree(p);
  wg::lu_plan_destroy(b->plan);
  delete b;
  return WG_OK;
}

int wg_batch_assemble(wg_batch* b, const double* profiles_dev, const double* vw_dev, double collisionMultiplier,
                      int parts, double* A_dev, long long lda, long long strideA, double* S_dev, void* stream) {
  WG_REQUIRE(b && profiles_dev && vw_dev && A_dev && S_dev, "wg_batch_assemble: NULL argument");
  wg_solver* s = b->s;
  WG_ON_DEVICE(s->device);
  if (!(s->haveGrid && s->haveTables && s->haveColl && s->haveStat)) {
    wg_set_error("wg_batch_assemble: grid, statistics, tables and collision tensor must be set first");
    return WG_ERR_STATE;
  }
  WG_REQUIRE(lda >= s->d.n && strideA >= lda * (long long)s->d.n, "wg_batch_assemble: lda < n or strideA < n*lda");
  wg::BzBatch bb;
  bb.count = b->count;
  bb.sA = strideA;
  bb.sV = s->d.n;
  bb.vwArr = vw_dev;
  return wg::bz_assemble(s->d, s->st, profiles_dev, 0.0, collisionMultiplier, parts, b->derivs, b->w1, b->w2, b->cT2,
                         A_dev, lda, S_dev, (cudaStream_t)stream, bb);
}

int wg_batch_factor(wg_batch* b, double* A_dev, long long lda, long long strideA, void* stream) {
  WG_REQUIRE(b && A_dev, "wg_batch_factor: NULL argument");
  WG_ON_DEVICE(b->s->device);
  return wg::lu_factor(b->plan, A_dev, lda, b->ipiv, b->info, (cudaStream_t)stream, nullptr, strideA, b->s->d.n);
}

int wg_batch_solve(wg_batch* b, const double* LU_dev, long long lda, long long strideA, double* rhs_dev, void* stream) {
  WG_REQUIRE(b && LU_dev && rhs_dev, "wg_batch_solve: NULL argument");
  WG_ON_DEVICE(b->s->device);
  return wg::lu_solve(b->plan, LU_dev, lda, rhs_dev, (cudaStream_t)stream, false, strideA, b->s->d.n);
}

int wg_batch_refine(wg_batch* b, const double* profiles_dev, const double* vw_dev, double collisionMultiplier,
                    const double* LU_dev, long long lda, long long strideA, double* x_dev, double* stats_dev, void* stream) {
  WG_REQUIRE(b && profiles_dev && vw_dev && LU_dev && x_dev && stats_dev, "wg_batch_refine: NULL argument");
  wg_solver* s = b->s;
  WG_ON_DEVICE(s->device);
  cudaStream_t st = (cudaStream_t)stream;
  wg::BzBatch bb;
  bb.count = b->count;
  bb.sA = strideA;
  bb.sV = s->d.n;
  bb.vwArr = vw_dev;
  int rc = wg::bz_residual(s->d, s->st, profiles_dev, 0.0, collisionMultiplier, b->derivs, b->w1, b->w2, b->cT2, x_dev,
                           b->bufA, b->bufB, st, bb);
  if (rc) return rc;
  rc = wg::lu_solve(b->plan, LU_dev, lda, b->bufB, st, false, strideA, s->d.n);
  if (rc) return rc;
  return wg::bz_refine_update(s->d.n, x_dev, b->bufB, stats_dev, st, b->count, s->d.n);
}

int wg_batch_info(wg_batch* b, void* stream, int* info_host) {
  WG_REQUIRE(b && info_host, "wg_batch_info: NULL argument");
  WG_ON_DEVICE(b->s->device);
  WG_CUDA(cudaMemcpyAsync(info_host, b->info, sizeof(int) * b->count, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  WG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return WG_OK;
}

int wg_batch_spectra(wg_batch* b, const double* deltaF_dev, double* spectra_dev, void* stream) {
  WG_REQUIRE(b && deltaF_dev && spectra_dev, "wg_batch_spectra: NULL argument");
  wg_solver* s = b->s;
  WG_ON_DEVICE(s->device);
  cudaStream_t st = (cudaStream_t)stream;
  b->chebOf = nullptr;
  int rc = apply_stage(s, 0, deltaF_dev, b->cheb, st, b->count, b->bufA, b->bufB);
  if (rc) return rc;
  rc = wg::bz_spectra(s->d, b->cheb, spectra_dev, st, b->count);
  if (rc == WG_OK) b->chebOf = deltaF_dev;
  return rc;
}

int wg_batch_moments(wg_batch* b, const double* deltaF_dev, const double* profiles_dev, const int* keep_dev,
                     int roundtrip, double* deltaF_out_dev, double* deltas_dev, double* err_dev, void* stream) {
  WG_REQUIRE(b && deltaF_dev && profiles_dev && keep_dev && deltaF_out_dev && deltas_dev && err_dev,
             "wg_batch_moments: NULL argument");
  wg_solver* s = b->s;
  WG_ON_DEVICE(s->device);
  WG_REQUIRE(s->haveGrid, "wg_batch_moments: grid not set");
  cudaStream_t st = (cudaStream_t)stream;
  if (b->chebOf != deltaF_dev) {
    wg_set_error("wg_batch_moments: call wg_batch_spectra on the same deltaF_dev first");
    return WG_ERR_STATE;
  }
  b->chebOf = nullptr;
  int rc = wg::bz_truncate(s->d, b->cheb, 0, 0, 0, err_dev, st, b->count, keep_dev);
  if (rc) return rc;
  if (roundtrip) {
    rc = apply_stage(s, 1, b->cheb, deltaF_out_dev, st, b->count, b->bufA, b->bufB);
    if (rc) return rc;
  } else if (deltaF_out_dev != deltaF_dev) {
    WG_CUDA(cudaMemcpyAsync(deltaF_out_dev, deltaF_dev, sizeof(double) * s->d.n * b->count, cudaMemcpyDeviceToDevice, st));
  }
  double* card = b->cheb;
  rc = apply_stage(s, 2, deltaF_out_dev, card, st, b->count, b->bufA, b->bufB);
  if (rc) return rc;
  return wg::bz_moments(s->d, s->st, card, profiles_dev, deltas_dev, st, b->count);
}

int wg_lu_create_batched(int n, int batch, int device, wg_lu** out) {
  WG_REQUIRE(out && n > 0 && batch >= 1, "wg_lu_create_batched: bad argument");
  int rc = require_device(device);
  if (rc) return rc;
  WG_ON_DEVICE(device);
  wg_lu* p = new wg_lu();
  p->device = device;
  rc = wg::lu_plan_create(n, &p->plan, batch);
  if (rc) {
    delete p;
    return rc;
  }
  *out = p;
  return WG_OK;
}

int wg_lu_factor_batched(wg_lu* p, double* A_dev, long long lda, long long strideA, int* ipiv_dev, int* info_dev,
                         void* stream) {
  WG_REQUIRE(p, "wg_lu_factor_batched: NULL plan");
  WG_ON_DEVICE(p->device);
  return wg::lu_factor(p->plan, A_dev, lda, ipiv_dev, info_dev, (cudaStream_t)stream, nullptr, strideA, 0);
}

int wg_lu_solve_batched(wg_lu* p, const double* LU_dev, long long lda, long long strideA, double* b_dev,
                        long long strideB, void* stream) {
  WG_REQUIRE(p, "wg_lu_solve_batched: NULL plan");
  WG_ON_DEVICE(p->device);
  return wg::lu_solve(p->plan, LU_dev, lda, b_dev, (cudaStream_t)stream, false, strideA, strideB);
}

int wg_lu_create(int n, int device, wg_lu** out) {
  WG_REQUIRE(out && n > 0, "wg_lu_create: bad argument");
  int rc = require_device(device);
  if (rc) return rc;
  WG_ON_DEVICE(device);
  wg_lu* p = new wg_lu();
  p->device = device;
  rc = wg::lu_plan_create(n, &p->plan);
  if (rc) {
    delete p;
    return rc;
  }
  *out = p;
  return WG_OK;
}

int wg_lu_destroy(wg_lu* p) {
  if (!p) return WG_OK;
  WG_ON_DEVICE(p->device);
  wg::lu_plan_destroy(p->plan);
  delete p;
  return WG_OK;
}

int wg_lu_factor(wg_lu* p, double* A_dev, long long lda, int* ipiv_dev, int* info_dev, void* stream) {
  WG_REQUIRE(p, "wg_lu_factor: NULL plan");
  WG_ON_DEVICE(p->device);
  return wg::lu_factor(p->plan, A_dev, lda, ipiv_dev, info_dev, (cudaStream_t)stream);
}

/* developer probe (not part of the public header): event timeline of the look-ahead LU */
int wg_debug_lu_trace(wg_lu* p, int on, double* out, int maxSteps) {
  wg::lu_set_trace(on);
  if (!p || !out) return 0;
  return wg::lu_trace_read(p->plan, out, maxSteps);
}

int wg_lu_solve(wg_lu* p, const double* LU_dev, long long lda, double* b_dev, void* stream) {
  WG_REQUIRE(p, "wg_lu_solve: NULL plan");
  WG_ON_DEVICE(p->device);
  return wg::lu_solve(p->plan, LU_dev, lda, b_dev, (cudaStream_t)stream);
}

int wg_dgemm_sub(double* C_dev, const double* A_dev, const double* B_dev, int M, int N, int K, long long ldc,
                 long long lda, long long ldb, void* stream) {
  WG_REQUIRE(C_dev && A_dev && B_dev && M >= 0 && N >= 0 && K >= 0, "wg_dgemm_sub: bad argument");
  return wg::dgemm_sub(C_dev, A_dev, B_dev, M, N, K, ldc, lda, ldb, (cudaStream_t)stream);
}

}  // extern "C"
