/*
 * CPU ORACLE (test infrastructure only) -- dense LU with partial pivoting.
 *
 * The reference solves the Boltzmann system with numpy.linalg.solve
 * (/root/reference/src/WallGo/boltzmann.py:283 and :548), i.e. LAPACK dgesv from
 * whatever BLAS NumPy links (pyproject.toml:37-41 pins only numpy>=1.21.5; here
 * scipy-openblas 0.3.30).  LAPACK is not under /root/reference, so this file
 * restates its published algorithm: unblocked right-looking Gaussian elimination
 * with row partial pivoting (dgetf2: pivot = first row of maximal |a_ij| in the
 * column, row interchange, scale, rank-1 update) and the two triangular solves of
 * dgetrs.  Row-major storage, 0-based pivots.  Parity unpinned at the bit level
 * (the reference's tests only pin the residual, tests/test_Boltzmann.py:88-125);
 * tests compare solutions at 1e-10 and P*A = L*U identities.
 *
 * Only tests/ and bench.py's cpu_baseline may load this library.
 */
#include <math.h>
#include <stddef.h>

/* Factor A (n x n, row-major, leading dimension lda) in place: P*A = L*U.
 * ipiv[j] = row swapped with row j at step j.  Returns 0, or j+1 if U[j][j]==0. */
int wgo_lu_factor(double *A, int n, int lda, int *ipiv)
{
    int info = 0;
    for (int j = 0; j < n; ++j) {
        int p = j;
        double best = fabs(A[(size_t)j * lda + j]);
        for (int i = j + 1; i < n; ++i) {
            double v = fabs(A[(size_t)i * lda + j]);
            if (v > best) { best = v; p = i; }
        }
        ipiv[j] = p;
        if (A[(size_t)p * lda + j] == 0.0) { if (!info) info = j + 1; continue; }
        if (p != j) {
            double *rj = A + (size_t)j * lda, *rp = A + (size_t)p * lda;
            for (int c = 0; c < n; ++c) { double t = rj[c]; rj[c] = rp[c]; rp[c] = t; }
        }
        const double *uj = A + (size_t)j * lda;
        const double inv = 1.0 / uj[j];
        for (int i = j + 1; i < n; ++i) {
            double *ri = A + (size_t)i * lda;
            const double l = ri[j] * inv;
            ri[j] = l;
            for (int c = j + 1; c < n; ++c) ri[c] -= l * uj[c];
        }
    }
    return info;
}

/* Solve A x = b with the factors above; b (length n) is overwritten by x. */
void wgo_lu_solve(const double *LU, int n, int lda, const int *ipiv, double *b)
{
    for (int j = 0; j < n; ++j) {
        int p = ipiv[j];
        if (p != j) { double t = b[j]; b[j] = b[p]; b[p] = t; }
    }
    for (int i = 0; i < n; ++i) {
        const double *ri = LU + (size_t)i * lda;
        double s = b[i];
        for (int c = 0; c < i; ++c) s -= ri[c] * b[c];
        b[i] = s;
    }
    for (int i = n - 1; i >= 0; --i) {
        const double *ri = LU + (size_t)i * lda;
        double s = b[i];
        for (int c = i + 1; c < n; ++c) s -= ri[c] * b[c];
        b[i] = s / ri[i];
    }
}
