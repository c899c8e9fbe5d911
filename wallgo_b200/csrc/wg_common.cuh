// wallgo_b200 -- shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#define WG_OK 0
#define WG_ERR_ARG (-1)
#define WG_ERR_CUDA (-2)
#define WG_ERR_STATE (-3)
#define WG_ERR_UNSUPPORTED (-4)

void wg_set_error(const char* fmt, ...);
extern std::atomic<long long> g_wg_launches;  // kernels launched by this library (bench accounting)

#define WG_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) {                                                       \
      wg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return WG_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

#define WG_REQUIRE(cond, msg)                                    \
  do {                                                           \
    if (!(cond)) {                                               \
      wg_set_error("%s:%d %s", __FILE__, __LINE__, msg);         \
      return WG_ERR_ARG;                                         \
    }                                                            \
  } while (0)

extern int g_wg_sync_every_launch;  // developer aid (env WG_SYNC_EVERY_LAUNCH=1): attribute faults to the kernel
#define WG_LAUNCH_CHECK()                                     \
  do {                                                        \
    ++g_wg_launches;                                          \
    WG_CUDA(cudaPeekAtLastError());                           \
    if (g_wg_sync_every_launch) WG_CUDA(cudaDeviceSynchronize()); \
  } while (0)

static inline int wg_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Function attributes (opt-in shared-memory size, non-portable cluster size) are per DEVICE, not per process:
// returns the largest dynamic shared-memory size already configured for `func` on the CURRENT device, or -1 if
// the function has not been configured there yet, and records `smem` as configured (wg_capi.cu).
long long wg_func_configured(const void* func, size_t smem);

// Priority carried by kernel launches of the LU (INT_MIN: none).  Stream priorities are not kept by
// stream capture, so the critical-chain kernels carry theirs as a launch attribute: it becomes the
// kernel node's priority inside a captured CUDA graph and is harmless on a plain stream launch.
extern thread_local int g_wg_launch_prio;

#ifdef __CUDACC__
#include <limits.h>
namespace wg {

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_prio(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                      Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if (g_wg_launch_prio != INT_MIN) {
    attr[0].id = cudaLaunchAttributePriority;
    attr[0].val.priority = g_wg_launch_prio;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---------------------------------------------------------------- FP64 tensor-core MMA
// D(8x8) = A(8x4, row) * B(4x8, col) + C.  lane = 4*g + t :
//   a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].       (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------- cp.async (LDGSTS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  int sz = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool pred) {
  int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---------------------------------------------------------------- TMA 1-D bulk copy + mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok = 0;
  const uint32_t a = smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(phase)
        : "memory");
  }
}
// global -> shared bulk copy (SASS: UBLKCP); bytes must be a multiple of 16, both sides 16B aligned
__device__ __forceinline__ void tma_load_1d(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem)),
      "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------- thread-block clusters / DSMEM
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive();
  cluster_wait();
}
// map a local shared address to the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_map(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void dsmem_st_f64(uint32_t addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;\n" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void dsmem_st_u64(uint32_t addr, unsigned long long v) {
  asm volatile("st.shared::cluster.u64 [%0], %1;\n" ::"r"(addr), "l"(v) : "memory");
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace wg
#endif
