// Internal declarations for the Boltzmann device path (wg_boltzmann.cu).
#pragma once
#include <cuda_runtime.h>

namespace wg {

struct BzDims {
  int P, M, N;   // species, spatial, momentum basis sizes (N odd)
  int M1, N1;    // M-1, N-1
  int q;         // (N-1)^2
  int n;         // P*(M-1)*q
};

// device pointers to everything that is constant for one grid / basis / collision tensor
struct BzStatic {
  double *chi, *rz, *rp, *pz, *pp, *dpzdrz, *dppdrp;  // grid (momentum maps fixed per run)
  double* stat;                                       // [P]  +1 boson, -1 fermion
  double *Tchi, *Dchi;                                // [M1*M1]
  double* DchiFull;                                   // [(M+1)^2]
  double *Trz, *Drz, *Trp;                            // [N1*N1]
  double *K1, *K2;                                    // [q*q]  Trz(x)Trp, Drz(x)Trp
  double* Coll;                                       // [(P q)^2]  C[a,bg,b,jk], solver basis
};

// strided batch of independent systems in lock step (count 1: one system)
struct BzBatch {
  int count = 1;
  long long sA = 0;               // doubles between consecutive operators
  long long sV = 0;               // doubles between consecutive source vectors
  const double* vwArr = nullptr;  // device array of wall velocities, one per system (nullptr: the scalar argument)
};

int bz_build_basis(const BzDims& d, const double* rz, double* Tchi, double* Dchi, double* DchiFull,
                   double* Trz, double* Drz, double* Trp, cudaStream_t st);
int bz_build_kron(const BzDims& d, const double* Trz, const double* Drz, const double* Trp, double* K1,
                  double* K2, cudaStream_t st);
void bz_set_asm_wide(int on);
int bz_assemble(const BzDims& d, const BzStatic& s, const double* prof, double vw, double collMult, int parts,
                double* derivs, double* w1, double* w2, double* cT2, double* A, long long lda, double* S,
                cudaStream_t st, const BzBatch& bb = BzBatch(), int y0 = 0, int ny = -1);
int bz_residual(const BzDims& d, const BzStatic& s, const double* prof, double vw, double collMult, double* derivs,
                double* w1, double* w2, double* cT2, const double* x, double* Ssrc, double* r, cudaStream_t st,
                const BzBatch& bb = BzBatch());
int bz_refine_update(int n, double* x, const double* dlt, double* stats, cudaStream_t st, int batch = 1, long long sV = 0);
int bz_apply_axis(const double* in, double* out, const double* Mat, int outer, int len, int inner,
                  cudaStream_t st);
int bz_spectra(const BzDims& d, const double* c, double* spectra, cudaStream_t st, int batch = 1);
int bz_truncate(const BzDims& d, double* c, int keepM, int keepN1, int keepN2, double* err, cudaStream_t st, int batch = 1,
                const int* keepArr = nullptr);
int bz_moments(const BzDims& d, const BzStatic& s, const double* card, const double* prof, double* deltas,
               cudaStream_t st, int batch = 1);
int bz_feedback(const BzDims& d, const double* deltas, const double* prof, const double* dofs, const double* dmsq,
                int F, double u0, double u3, double* out, cudaStream_t st);
int bz_collision_interp(int P, int nF, int nT, const double* Lz, const double* Lp, const double* src, double* out,
                        cudaStream_t st);
int bz_collision_matvec(const BzDims& d, const BzStatic& s, const double* cT2, const double* x, double* y,
                        cudaStream_t st);

}  // namespace wg
