"""Host-side (NumPy) helpers for the small, per-run tables of the spectral grid.

These are O(M^2)/O(N^2) set-up computations (done once per grid or basis change, like the
reference does in polynomial.py); the per-solve work is on the GPU.  The default-basis
tables are evaluated on the device by wg_build_basis; this module provides the tables for
the other basis / derivative options and the basis-change matrices used by getDeltas.
"""
from __future__ import annotations

import numpy as np


def lobattoNodes(M: int, N: int):
    """Interior Gauss-Lobatto nodes (reference grid.py:116-122)."""
    chi = -np.cos(np.arange(1, M) * np.pi / M)
    rz = -np.cos(np.arange(1, N) * np.pi / N)
    rp = -np.cos(np.arange(0, N - 1) * np.pi / (N - 1))
    return chi, rz, rp


def _chebT(n, x):
    return np.cos(n[None, :] * np.arccos(np.clip(x, -1.0, 1.0))[:, None])


def _chebU(n, x):
    th = np.arccos(np.clip(x, -1.0, 1.0))[:, None]
    return np.sin((n[None, :] + 1) * th) / np.sin(th)


def chebyshevMatrix(x, size: int, restriction: str):
    """[point, n] matrix of the restricted basis (reference polynomial.py:484-488, 713-745)."""
    x = np.asarray(x, dtype=float)
    if restriction == "full":
        n = np.arange(2, size + 2)
        return _chebT(n, x) - np.where(n[None, :] % 2 == 0, 1.0, x[:, None])
    n = np.arange(1, size + 1)
    return _chebT(n, x) - 1.0


def chebyshevDerivMatrix(x, size: int, restriction: str):
    """d/dx of the restricted basis at interior points x (reference polynomial.py:857-860)."""
    x = np.asarray(x, dtype=float)
    if restriction == "full":
        n = np.arange(2, size + 2)
        return n[None, :] * _chebU(n - 1, x) - np.where(n[None, :] % 2 == 0, 0.0, 1.0)
    n = np.arange(1, size + 1)
    return n[None, :] * _chebU(n - 1, x)


def cardinalDerivFull(xfull):
    """D[j,i] = C_i'(x_j) for Lagrange cardinals on arbitrary nodes (reference polynomial.py:770-823),
    via barycentric weights."""
    x = np.asarray(xfull, dtype=float)
    L = x.size
    diff = x[:, None] - x[None, :]
    np.fill_diagonal(diff, 1.0)
    w = 1.0 / np.prod(diff, axis=1)              # barycentric weights
    D = (w[None, :] / w[:, None]) / diff          # D[j,i] = (w_i/w_j)/(x_j-x_i)
    np.fill_diagonal(D, 0.0)
    np.fill_diagonal(D, -np.sum(D, axis=1))
    return D


def finiteDiffMatrix(g):
    """Three-point non-uniform finite differences incl. one-sided ends (reference helpers.py:428-454)."""
    g = np.asarray(g, dtype=float)
    n = g.size
    D = np.zeros((n, n))
    h1 = g[1:-1] - g[:-2]
    h2 = g[2:] - g[1:-1]
    idx = np.arange(1, n - 1)
    D[idx, idx - 1] = -h2 / (h1 * (h1 + h2))
    D[idx, idx] = (h2 - h1) / (h1 * h2)
    D[idx, idx + 1] = h1 / (h2 * (h1 + h2))
    for s, (a, b, c) in ((0, (0, 1, 2)), (-1, (-1, -2, -3))):
        D[s, a] = (2 * g[a] - g[b] - g[c]) / ((g[a] - g[b]) * (g[a] - g[c]))
        D[s, b] = (g[c] - g[a]) / ((g[a] - g[b]) * (g[b] - g[c]))
        D[s, c] = (g[b] - g[a]) / ((g[a] - g[c]) * (g[c] - g[b]))
    return D


def buildTables(M, N, basisM, basisN, derivatives):
    """Tchi, Dchi, DchiFull, Trz, Drz, Trp for any basis/derivative option (boltzmann.py:673-726)."""
    chi, rz, rp = lobattoNodes(M, N)
    chiF = np.concatenate([[-1.0], chi, [1.0]])
    rzF = np.concatenate([[-1.0], rz, [1.0]])
    if derivatives == "Spectral":
        DchiFull = cardinalDerivFull(chiF)
        if basisM == "Cardinal":
            Tchi, Dchi = np.identity(M - 1), DchiFull[1:-1, 1:-1]
        else:
            Tchi, Dchi = chebyshevMatrix(chi, M - 1, "full"), chebyshevDerivMatrix(chi, M - 1, "full")
        if basisN == "Cardinal":
            Trz = np.identity(N - 1)
            Trp = np.identity(N - 1)
            Drz = cardinalDerivFull(rzF)[1:-1, 1:-1]
        else:
            Trz = chebyshevMatrix(rz, N - 1, "full")
            Trp = chebyshevMatrix(rp, N - 1, "partial")
            Drz = chebyshevDerivMatrix(rz, N - 1, "full")
    else:
        Tchi = np.identity(M - 1)
        Trz = np.identity(N - 1)
        Trp = np.identity(N - 1)
        DchiFull = finiteDiffMatrix(chiF)
        Dchi = DchiFull[1:-1, 1:-1]
        Drz = finiteDiffMatrix(rzF)[1:-1, 1:-1]
    return Tchi, Dchi, DchiFull, Trz, Drz, Trp


def basisChangeMatrices(M, N, basisM, basisN):
    """The nine optional matrices of wg_set_transforms (Polynomial.changeBasis, polynomial.py:222-298):
    stage 0 solver->Chebyshev, 1 Chebyshev->solver, 2 solver->Cardinal; axis z, pz, pp."""
    chi, rz, rp = lobattoNodes(M, N)
    cz = chebyshevMatrix(chi, M - 1, "full")
    cpz = chebyshevMatrix(rz, N - 1, "full")
    cpp = chebyshevMatrix(rp, N - 1, "partial")
    mats = [None] * 9
    if basisM == "Cardinal":
        mats[0], mats[3] = np.linalg.inv(cz), cz
    else:
        mats[6] = cz
    if basisN == "Cardinal":
        mats[1], mats[2] = np.linalg.inv(cpz), np.linalg.inv(cpp)
        mats[4], mats[5] = cpz, cpp
    else:
        mats[7], mats[8] = cpz, cpp
    return mats


def tailConverging(coeffs, offset: int) -> bool:
    """Sign of the least-squares slope of log|c_n| (reference polynomial.py:994-1033).  The slope of a
    straight-line fit has the sign of sum((x - mean x)(y - mean y)); that closed form decides whenever it
    is clear of rounding, anything borderline or non-finite goes through np.polyfit exactly as the reference
    does."""
    m = len(coeffs)
    if m < 2:
        return False
    with np.errstate(divide="ignore"):
        y = np.log(np.abs(coeffs))
    if np.all(np.isfinite(y)):
        xc = np.arange(m, dtype=np.float64)
        xc -= xc.mean()
        yc = y - y.mean()
        s = float(xc @ yc)
        if abs(s) > 1e-8 * float(np.abs(xc) @ np.abs(yc)):
            return s < 0.0
    x = np.arange(m) + offset
    fit = np.polyfit(x, y, 1, cov=(m >= 3))
    slope = fit[0][0] if m >= 3 else fit[0]
    return bool(slope < 0)


def spectralPeak(coeffs, power: int) -> int:
    """Histogram-mean index with weight (1+n)^power (reference polynomial.py:1000, 1036-1038)."""
    m = len(coeffs)
    if m < 2:
        return 0
    idx = np.arange(m)
    w = (1 + idx) ** power
    a = np.abs(coeffs)
    return int(np.sum(idx * w * a) / np.sum(a * w))
