// Internal (C++) declarations shared by the wallgo_b200 translation units.
#pragma once
#include <cuda_runtime.h>

namespace wg {

// strided batch of independent products (blockIdx.z); one system: count 1, strides 0
struct GemmBatch {
  int count = 1;
  long long sC = 0, sA = 0, sB = 0;
};
// C[MxN] -= A[MxK] * B[KxN], row-major, FP64 tensor cores (wg_gemm.cu)
int dgemm_sub(double* C, const double* A, const double* B, int M, int N, int K, long long ldc,
              long long lda, long long ldb, cudaStream_t st, GemmBatch gb = GemmBatch());

void gemm_set_variant(int v);

// Blocked LU with partial pivoting (wg_lu.cu)
struct LuPlan;
// batch > 1: every launch serves `batch` independent systems in lock step (strided batch)
int lu_plan_create(int n, LuPlan** out, int batch = 1);
void lu_plan_destroy(LuPlan* p);
// rhs != nullptr: the forward solve of that right-hand side is fused into the factorisation
// strideA / strideRhs: doubles between consecutive systems of the batch (ipiv: n ints, info: 1 int per system)
int lu_factor(LuPlan* p, double* A, long long lda, int* ipiv, int* info, cudaStream_t st, double* rhs = nullptr,
              long long strideA = 0, long long strideRhs = 0);
int lu_solve(LuPlan* p, const double* LU, long long lda, double* b, cudaStream_t st, bool backwardOnly = false,
             long long strideA = 0, long long strideB = 0);
void lu_set_tuning(int nb, int use_graph);
void lu_set_wave_solve(int on);
void lu_plan_set_tuning(LuPlan* p, int nb, int use_graph, int lookahead, int split, int hi_rest);
void lu_set_debug(long long* dev_buf, int c0);
void lu_set_lookahead(int on);
void lu_set_profile(int on);
int lu_profile_read(LuPlan* p, double* ms, double* flops, int* count);
void lu_set_trace(int on);
void lu_set_leaf32(int on);
void lu_set_force_wb(int wb);
int lu_leaf_fail_read(int* out);
void lu_set_split(int on);
void lu_set_bulk_hi(int rest);
int lu_trace_read(LuPlan* p, double* out, int maxSteps);

}  // namespace wg
