"""
CPU ORACLE for the WallGo Boltzmann hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the algorithm implemented by the reference
(`/root/reference/src/WallGo`, v1.1.2).  It exists so that the CUDA product in
`wallgo_b200/` can be checked on a GPU box where the reference is not present.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import it.  The product package never imports this module and has no CPU path.

Parity pinning: the oracle is pinned against the *live reference* in this
build container by `oracle/make_golden.py` (which imports WallGo from
`/root/reference/src`), and against the reference's own known-answer tests
(`tests/test_Boltzmann.py:14-85` analytic Delta00, `:88-125` residual,
`:128-174` truncation decisions) re-stated in `tests/test_oracle_golden.py`.
The solve itself goes through LAPACK `dgesv` (`np.linalg.solve`), exactly as
`boltzmann.py:283` does; `oracle/lu_oracle.c` restates the published
partial-pivoting algorithm (LAPACK dgetf2/dgetrs) for an independent check.

All arrays are C-ordered float64.  Index conventions follow SURVEY.md section 8:
row r = ((a*(M-1)+alpha)*(N-1)+beta)*(N-1)+gamma   (collocation point),
col c = ((b*(M-1)+i    )*(N-1)+j   )*(N-1)+k      (basis function).
"""
from __future__ import annotations

import sys
from dataclasses import dataclass, field

import numpy as np

MAX_EXPONENT = sys.float_info.max_exp * np.log(2)  # boltzmann.py:43


# --------------------------------------------------------------------------
# Grid (grid.py:116-122, 339-378; grid3Scales.py:145-307)
# --------------------------------------------------------------------------
def lobatto_nodes(M: int, N: int):
    """Interior Gauss-Lobatto nodes chi[M-1], rz[N-1], rp[N-1] (grid.py:116-122)."""
    chi = -np.cos(np.arange(1, M) * np.pi / M)
    rz = -np.cos(np.arange(1, N) * np.pi / N)
    rp = -np.cos(np.arange(0, N - 1) * np.pi / (N - 1))
    return chi, rz, rp


def with_endpoints(x: np.ndarray, direction: str) -> np.ndarray:
    """grid.py:203-206: chi, rz get -1 and +1; rp only gets +1 appended."""
    if direction == "pp":
        return np.concatenate([x, [1.0]])
    return np.concatenate([[-1.0], x, [1.0]])


def momentum_maps(rz, rp, T0):
    """p_z, p_par and Jacobians (grid.py:339-345, 375-377)."""
    pz = 2 * T0 * np.arctanh(rz)
    pp = -T0 * np.log((1 - rp) / 2)
    dpzdrz = 2 * T0 / (1 - rz**2)
    dppdrp = T0 / (1 - rp)
    return pz, pp, dpzdrz, dppdrp


def position_map_simple(chi, L):
    """Grid.decompactify / compactificationDerivatives in z (grid.py:339, 373)."""
    return L * chi / np.sqrt(1 - chi**2), L / (1 - chi**2) ** 1.5


def position_map_3scales(chi, tailIn, tailOut, L, r=0.5, smoothing=0.1, center=0.0):
    """Grid3Scales z(chi), dz/dchi (grid3Scales.py:145-307)."""
    def a_of(tail):
        num = 4 * smoothing * L * r**2 * (2 * r * tail - L * (1 + smoothing))
        return np.sqrt(num) / abs(2 * r * tail - L * (1 + 2 * smoothing))

    aIn, aOut = a_of(tailIn), a_of(tailOut)

    def rat(tail, a, x, sgn_r, sgn_outer):
        # the four atanh terms of grid3Scales.py:214-264, real part of complex atanh
        rr = sgn_r * r
        root = np.sqrt(a**2 + (x - rr) ** 2)
        den = np.sqrt(a**2 + (1 - sgn_outer * rr) ** 2)
        arg = (1 - sgn_outer * x + sgn_outer * root) / den
        return np.arctanh(arg + 0j).real / den

    def total(x):
        x = np.asarray(x, dtype=float)
        cOut = 2 * r * tailOut - L
        cIn = 2 * r * tailIn - L
        t1 = (1 - r) * cOut * rat(tailOut, aOut, x, +1, +1) / r
        t2 = -(1 + r) * cOut * rat(tailOut, aOut, x, +1, -1) / r
        t3 = (1 - r) * cIn * rat(tailIn, aIn, x, -1, -1) / r
        t4 = -(1 + r) * cIn * rat(tailIn, aIn, x, -1, +1) / r
        t5 = (2 * tailIn + 2 * tailOut - 4 * smoothing * L / r) * np.arctanh(x)
        return (t1 + t2 + t3 + t4 + t5) / 2

    xi = total(chi) - total(0.0) + center
    d = (2 * tailIn - L / r) * (1 - (chi + r) / np.sqrt(aIn**2 + (chi + r) ** 2)) / 2
    d = d + (2 * tailOut - L / r) * (1 + (chi - r) / np.sqrt(aOut**2 + (chi - r) ** 2)) / 2
    d = d + (1 - 2 * smoothing) * L / r
    d = d / (1 - chi**2)
    return xi, d


# --------------------------------------------------------------------------
# Chebyshev / cardinal tables (polynomial.py:445-490, 691-862)
# --------------------------------------------------------------------------
def cheb_T(nmax: int, x: np.ndarray) -> np.ndarray:
    """T_n(x) for n=0..nmax by the three-term recurrence -> array [len(x), nmax+1]."""
    x = np.asarray(x, dtype=float)
    out = np.empty((x.size, nmax + 1))
    out[:, 0] = 1.0
    if nmax >= 1:
        out[:, 1] = x
    for n in range(1, nmax):
        out[:, n + 1] = 2 * x * out[:, n] - out[:, n - 1]
    return out


def cheb_U(nmax: int, x: np.ndarray) -> np.ndarray:
    """U_n(x), n=0..nmax."""
    x = np.asarray(x, dtype=float)
    out = np.empty((x.size, nmax + 1))
    out[:, 0] = 1.0
    if nmax >= 1:
        out[:, 1] = 2 * x
    for n in range(1, nmax):
        out[:, n + 1] = 2 * x * out[:, n] - out[:, n - 1]
    return out


def cheb_matrix(x: np.ndarray, size: int, restriction: str) -> np.ndarray:
    """Restricted Chebyshev matrix [point, n] (polynomial.py:484-488, 713-745).
    'full': n=2..size+1, T_n-1 (n even) / T_n-x (n odd); 'partial': n=1..size, T_n-1."""
    if restriction == "full":
        n = np.arange(2, size + 2)
        T = cheb_T(size + 1, x)[:, n]
        return T - np.where(n[None, :] % 2 == 0, 1.0, x[:, None])
    n = np.arange(1, size + 1)
    return cheb_T(size, x)[:, n] - 1.0


def cheb_deriv_matrix(x: np.ndarray, size: int, restriction: str) -> np.ndarray:
    """d/dx of the restricted basis at points x (polynomial.py:857-860)."""
    if restriction == "full":
        n = np.arange(2, size + 2)
        U = cheb_U(size + 1, x)
        return n[None, :] * U[:, n - 1] - np.where(n[None, :] % 2 == 0, 0.0, 1.0)
    n = np.arange(1, size + 1)
    return n[None, :] * cheb_U(size, x)[:, n - 1]


def cardinal_deriv_full(xfull: np.ndarray) -> np.ndarray:
    """D[j,i] = C_i'(x_j) for Lagrange cardinals on nodes xfull (polynomial.py:770-823)."""
    L = xfull.size
    D = np.empty((L, L))
    for i in range(L):
        for j in range(L):
            if i == j:
                D[j, i] = sum(1.0 / (xfull[i] - xfull[k]) for k in range(L) if k != i)
            else:
                p = 1.0
                for k in range(L):
                    if k != i and k != j:
                        p *= (xfull[j] - xfull[k]) / (xfull[i] - xfull[k])
                D[j, i] = p / (xfull[i] - xfull[j])
    return D


def finite_diff_matrix(g: np.ndarray) -> np.ndarray:
    """Three-point non-uniform finite differences (helpers.py:428-454)."""
    n = g.size
    D = np.zeros((n, n))
    for i in range(1, n - 1):
        dx1 = -(g[i] - g[i - 1])
        dx2 = g[i + 1] - g[i]
        D[i, i - 1] = dx2 / (dx1 * (dx2 - dx1))
        D[i, i] = -(dx1 + dx2) / (dx1 * dx2)
        D[i, i + 1] = dx1 / (dx2 * (dx1 - dx2))
    D[0, 0] = (2 * g[0] - g[1] - g[2]) / ((g[0] - g[1]) * (g[0] - g[2]))
    D[0, 1] = (-g[0] + g[2]) / ((g[0] - g[1]) * (g[1] - g[2]))
    D[0, 2] = (-g[0] + g[1]) / ((g[0] - g[2]) * (-g[1] + g[2]))
    D[-1, -1] = (2 * g[-1] - g[-2] - g[-3]) / ((g[-1] - g[-2]) * (g[-1] - g[-3]))
    D[-1, -2] = (-g[-1] + g[-3]) / ((g[-1] - g[-2]) * (g[-2] - g[-3]))
    D[-1, -3] = (-g[-1] + g[-2]) / ((g[-1] - g[-3]) * (-g[-2] + g[-3]))
    return D


@dataclass
class Tables:
    """Everything that is constant for one (M, N, bases, derivatives) choice."""
    M: int
    N: int
    basisM: str
    basisN: str
    derivatives: str
    chi: np.ndarray
    rz: np.ndarray
    rp: np.ndarray
    Tchi: np.ndarray   # (M-1)^2 intertwiner  [alpha, i]
    Dchi: np.ndarray   # (M-1)^2 d/dchi       [alpha, i]
    Trz: np.ndarray    # (N-1)^2              [beta, j]
    Drz: np.ndarray    # (N-1)^2 d/drz        [beta, j]
    Trp: np.ndarray    # (N-1)^2              [gamma, k]
    DchiFull: np.ndarray  # (M+1)^2 profile derivative matrix [point, node]
    # Chebyshev matrices used by basis changes (always the restricted Chebyshev ones)
    chebChi: np.ndarray
    chebRz: np.ndarray
    chebRp: np.ndarray


def build_tables(M, N, basisM="Cardinal", basisN="Chebyshev", derivatives="Spectral") -> Tables:
    """Basis and derivative tables used by buildLinearEquations (boltzmann.py:673-726)."""
    chi, rz, rp = lobatto_nodes(M, N)
    chiF = with_endpoints(chi, "z")
    rzF = with_endpoints(rz, "pz")
    chebChi = cheb_matrix(chi, M - 1, "full")
    chebRz = cheb_matrix(rz, N - 1, "full")
    chebRp = cheb_matrix(rp, N - 1, "partial")
    cardFull = cardinal_deriv_full(chiF)
    if derivatives == "Spectral":
        DchiFull = cardFull
        if basisM == "Cardinal":
            Tchi = np.identity(M - 1)
            Dchi = cardFull[1:-1, 1:-1]
        else:
            Tchi = chebChi
            Dchi = cheb_deriv_matrix(chi, M - 1, "full")
        if basisN == "Cardinal":
            Trz = np.identity(N - 1)
            Trp = np.identity(N - 1)
            Drz = cardinal_deriv_full(rzF)[1:-1, 1:-1]
        else:
            Trz, Trp = chebRz, chebRp
            Drz = cheb_deriv_matrix(rz, N - 1, "full")
    else:
        assert basisM == "Cardinal" and basisN == "Cardinal"  # boltzmann.py:97-100
        Tchi = np.identity(M - 1)
        Trz = np.identity(N - 1)
        Trp = np.identity(N - 1)
        DchiFull = finite_diff_matrix(chiF)
        Dchi = DchiFull[1:-1, 1:-1]
        Drz = finite_diff_matrix(rzF)[1:-1, 1:-1]
    return Tables(M, N, basisM, basisN, derivatives, chi, rz, rp, Tchi, Dchi, Trz, Drz, Trp,
                  DchiFull, chebChi, chebRz, chebRp)


def collision_to_chebyshev(Ccard: np.ndarray, tab: Tables) -> np.ndarray:
    """CollisionArray.changeBasis('Chebyshev') with inverseTranspose=True
    (collisionArray.py:356-388, polynomial.py:282-297): contract the two polynomial
    axes with transpose(inv(inv(T))) exactly as the reference forms it."""
    out = Ccard
    for axis, T in ((4, tab.chebRz), (5, tab.chebRp)):
        mat = np.transpose(np.linalg.inv(np.linalg.inv(T)))
        out = np.moveaxis(np.tensordot(mat, out, axes=([1], [axis])), 0, axis)
    return np.ascontiguousarray(out)


# --------------------------------------------------------------------------
# Background (containers.py:152-157, helpers.py:464-468)
# --------------------------------------------------------------------------
def boost_velocity(v, vmid):
    return (v - vmid) / (1.0 - v * vmid)


@dataclass
class Problem:
    """Inputs of one Boltzmann solve, already boosted to the plasma frame."""
    tab: Tables
    xi_dxidchi: np.ndarray        # dxi/dchi on interior nodes [M-1]
    pz: np.ndarray
    pp: np.ndarray
    dpzdrz: np.ndarray
    dppdrp: np.ndarray
    T: np.ndarray                 # [M+1]
    v: np.ndarray                 # [M+1] plasma-frame
    msq: np.ndarray               # [P, M+1]
    vw: float
    statistics: np.ndarray        # [P] +1 boson / -1 fermion
    collision: np.ndarray         # (P,N-1,N-1,P,N-1,N-1) in basisN
    collisionMultiplier: float = 1.0
    dofs: np.ndarray = field(default_factory=lambda: np.zeros(0))


def dfeq(x, s):
    """boltzmann.py:876-886."""
    with np.errstate(over="ignore"):
        return np.where(x > MAX_EXPONENT, -0.0, -1 / (np.exp(x) - 2 * s + np.exp(-x)))


def feq(x, s):
    """boltzmann.py:864-874."""
    with np.errstate(over="ignore"):
        return np.where(x > MAX_EXPONENT, 0.0, 1 / (np.exp(x) - s))


def row_quantities(pb: Problem):
    """Per-row kinematics (boltzmann.py:651-665, 694-698, 729-746)."""
    tab = pb.tab
    T = pb.T[1:-1][None, :, None, None]
    v = pb.v[1:-1][None, :, None, None]
    msq = pb.msq[:, 1:-1, None, None]
    pz = pb.pz[None, None, :, None]
    pp = pb.pp[None, None, None, :]
    E = np.sqrt(msq + pz**2 + pp**2)
    dT = (tab.DchiFull @ pb.T)[1:-1][None, :, None, None]
    dv = (tab.DchiFull @ pb.v)[1:-1][None, :, None, None]
    dmsq = (pb.msq @ tab.DchiFull.T)[:, 1:-1, None, None]
    gw = 1 / np.sqrt(1 - pb.vw**2)
    Pw = gw * (pz - pb.vw * E)
    gpl = 1 / np.sqrt(1 - v**2)
    Epl = gpl * (E - v * pz)
    Ppl = gpl * (pz - v * E)
    uwupl = gw * gpl * (pb.vw - v)
    dchidxi = 1 / pb.xi_dxidchi[None, :, None, None]
    drzdpz = 1 / pb.dpzdrz[None, None, :, None]
    s = pb.statistics[:, None, None, None]
    return dict(T=T, v=v, msq=msq, E=E, dT=dT, dv=dv, dmsq=dmsq, gw=gw, Pw=Pw, gpl=gpl,
                Epl=Epl, Ppl=Ppl, uwupl=uwupl, dchidxi=dchidxi, drzdpz=drzdpz, s=s)


def assemble(pb: Problem, want_collision: bool = False):
    """operator (n,n), source (n,) of boltzmann.py:628-820."""
    tab = pb.tab
    P = pb.msq.shape[0]
    M1, N1 = tab.M - 1, tab.N - 1
    q = N1 * N1
    n = P * M1 * q
    rq = row_quantities(pb)
    src = (dfeq(rq["Epl"] / rq["T"], rq["s"]) / rq["T"]) * rq["dchidxi"] * (
        rq["Pw"] * rq["Ppl"] * rq["gpl"] ** 2 * rq["dv"]
        + rq["Pw"] * rq["Epl"] * rq["dT"] / rq["T"]
        + 0.5 * rq["dmsq"] * rq["uwupl"]
    )
    w1 = np.broadcast_to(rq["dchidxi"] * rq["Pw"], (P, M1, N1, N1))
    w2 = np.broadcast_to(rq["dchidxi"] * rq["drzdpz"] * (rq["gw"] / 2) * rq["dmsq"], (P, M1, N1, N1))
    K1 = np.kron(tab.Trz, tab.Trp)   # [(beta,gamma),(j,k)]
    K2 = np.kron(tab.Drz, tab.Trp)
    A = np.zeros((P, M1, q, P, M1, q))
    Cq = pb.collision.reshape(P, q, P, q)
    cT2 = pb.collisionMultiplier * pb.T[1:-1] ** 2
    for a in range(P):
        w1a = w1[a].reshape(M1, q)
        w2a = w2[a].reshape(M1, q)
        # Liouville term 1: w1 * Dchi[alpha,i] * K1
        A[a, :, :, a, :, :] += (w1a[:, :, None, None] * tab.Dchi[:, None, :, None]) * K1[None, :, None, :]
        # Liouville term 2: -w2 * Tchi[alpha,i] * K2
        A[a, :, :, a, :, :] -= (w2a[:, :, None, None] * tab.Tchi[:, None, :, None]) * K2[None, :, None, :]
        for b in range(P):
            A[a, :, :, b, :, :] += (cT2[:, None, None, None] * tab.Tchi[:, None, :, None]) * Cq[a][None, :, b, None, :]
    op = A.reshape(n, n)
    if want_collision:
        coll = np.zeros((P, M1, q, P, M1, q))
        for a in range(P):
            for b in range(P):
                coll[a, :, :, b, :, :] = (cT2[:, None, None, None] * tab.Tchi[:, None, :, None]) * Cq[a][None, :, b, None, :]
        return op, src.reshape(n), coll.reshape(n, n)
    return op, src.reshape(n)


def solve(pb: Problem) -> np.ndarray:
    """solveBoltzmannEquations (boltzmann.py:240-294): LAPACK dgesv."""
    op, src = assemble(pb)
    P = pb.msq.shape[0]
    return np.linalg.solve(op, src).reshape(P, pb.tab.M - 1, pb.tab.N - 1, pb.tab.N - 1)


# --------------------------------------------------------------------------
# Basis changes, truncation, moments (boltzmann.py:141-238, 296-491)
# --------------------------------------------------------------------------
def _apply(mat, arr, axis):
    return np.moveaxis(np.tensordot(mat, arr, axes=([1], [axis])), 0, axis)


def to_chebyshev(dF, tab: Tables):
    out = dF
    if tab.basisM == "Cardinal":
        out = _apply(np.linalg.inv(tab.chebChi), out, 1)
    if tab.basisN == "Cardinal":
        out = _apply(np.linalg.inv(tab.chebRz), out, 2)
        out = _apply(np.linalg.inv(tab.chebRp), out, 3)
    return out


def from_chebyshev(c, tab: Tables):
    out = c
    if tab.basisM == "Cardinal":
        out = _apply(tab.chebChi, out, 1)
    if tab.basisN == "Cardinal":
        out = _apply(tab.chebRz, out, 2)
        out = _apply(tab.chebRp, out, 3)
    return out


def to_cardinal(dF, tab: Tables):
    out = dF
    if tab.basisM == "Chebyshev":
        out = _apply(tab.chebChi, out, 1)
    if tab.basisN == "Chebyshev":
        out = _apply(tab.chebRz, out, 2)
        out = _apply(tab.chebRp, out, 3)
    return out


def tail_converging(coeffs: np.ndarray, offset: int) -> bool:
    """SpectralConvergenceInfo.apparentConvergence (polynomial.py:994-1033)."""
    m = coeffs.size
    if m < 2:
        return False
    x = np.arange(m) + offset
    slope = np.polyfit(x, np.log(np.abs(coeffs)), 1, cov=(m >= 3))
    slope = slope[0][0] if m >= 3 else slope[0]
    return bool(slope < 0)


def spectral_peak(coeffs: np.ndarray, power: int) -> int:
    """polynomial.py:1000, 1036-1038."""
    if coeffs.size < 2:
        return 0
    idx = np.arange(coeffs.size)
    w = (1 + idx) ** power
    a = np.abs(coeffs)
    return int(np.sum(idx * w * a) / np.sum(a * w))


def check_spectral_convergence(dF, tab: Tables, option: str = "AUTO"):
    """boltzmann.py:353-491.  option in {'NONE','AUTO','THIRD'}."""
    c = to_chebyshev(dF, tab).copy()
    a = np.abs(c)
    sChi, sPz, sPp = a.sum(axis=(0, 2, 3)), a.sum(axis=(0, 1, 3)), a.sum(axis=(0, 1, 2))
    M1, N1 = tab.M - 1, tab.N - 1
    cutM, cutN = M1 // 3, N1 // 3
    conv = (
        tail_converging(sChi[M1 - cutM:], M1 - cutM) if cutM > 0 else tail_converging(sChi, 0),
        tail_converging(sPz[N1 - cutN:], N1 - cutN) if cutN > 0 else tail_converging(sPz, 0),
        tail_converging(sPp[N1 - cutN:], N1 - cutN) if cutN > 0 else tail_converging(sPp, 0),
    )
    shape = list(dF.shape)
    cut = [False, False, False]
    if option == "AUTO":
        cut = [not conv[0], not conv[1], not conv[2]]
    elif option == "THIRD":
        cut = [True, True, True]
    # note: python's a[-0:] is the whole axis (boltzmann.py:400-446 with cut == 0)
    if cut[0]:
        c[:, (M1 - cutM) if cutM else 0:, :, :] = 0
        shape[1] = M1 - cutM
    if cut[1]:
        c[:, :, (N1 - cutN) if cutN else 0:, :] = 0
        shape[2] = N1 - cutN
    if cut[2]:
        c[:, :, :, (N1 - cutN) if cutN else 0:] = 0
        shape[3] = N1 - cutN
    peaks = (spectral_peak(sChi[:shape[1]], 0), spectral_peak(sPz[:shape[2]], 1),
             spectral_peak(sPp[:shape[3]], 2))
    if option == "NONE":
        return dF, tuple(shape), peaks
    return from_chebyshev(c, tab), tuple(shape), peaks


def estimate_truncation_error(dF, shape, tab: Tables) -> float:
    """boltzmann.py:296-351."""
    c = to_chebyshev(dF, tab)[:shape[0], :shape[1], :shape[2], :shape[3]]
    a = np.abs(c)
    tot = a.sum(axis=(1, 2, 3))
    e1 = a[:, -1, :, :].sum(axis=(1, 2)) / tot
    e2 = a[:, :, -1, :].sum(axis=(1, 2)) / tot
    e3 = a[:, :, :, -1].sum(axis=(1, 2)) / tot
    return float(max(e1.max(), e2.max(), e3.max()))


def moments(dF, pb: Problem):
    """Delta00, Delta02, Delta20, Delta11 [P, M-1] (boltzmann.py:182-225, polynomial.py:539-567)."""
    tab = pb.tab
    card = to_cardinal(dF, tab)
    pz = pb.pz[None, None, :, None]
    pp = pb.pp[None, None, None, :]
    E = np.sqrt(pb.msq[:, 1:-1, None, None] + pz**2 + pp**2)
    base = pb.dpzdrz[None, None, :, None] * pb.dppdrp[None, None, None, :] * pp / (4 * np.pi**2 * E)
    wz = np.sqrt(1 - tab.rz**2) * (np.pi / tab.N)
    wp = np.full(tab.N - 1, np.pi / (tab.N - 1))
    wp[0] /= 2
    wp = np.sqrt(1 - tab.rp**2) * wp
    W = wz[None, None, :, None] * wp[None, None, None, :]
    out = []
    for extra in (1.0, pz**2, E**2, E * pz):
        out.append(np.sum(extra * base * card * W, axis=(2, 3)))
    return np.stack(out)  # [4, P, M-1]


def get_deltas(pb: Problem, dF=None, option: str = "AUTO"):
    """BoltzmannSolver.getDeltas (boltzmann.py:141-238)."""
    if dF is None:
        dF = solve(pb)
    dFt, shape, peaks = check_spectral_convergence(dF, pb.tab, option)
    err = estimate_truncation_error(dFt, shape, pb.tab)
    tail = (shape[1] != dF.shape[1], shape[2] != dF.shape[2], shape[3] != dF.shape[3])
    return dict(deltaF=dFt, Deltas=moments(dFt, pb), truncationError=err,
                truncatedTail=tail, spectralPeaks=peaks, shapeTruncated=shape)


def check_linearization(pb: Problem, dF=None):
    """boltzmann.py:500-626: the two linearisation criteria."""
    tab = pb.tab
    if dF is None:
        dF = solve(pb)
    P = pb.msq.shape[0]
    card = to_cardinal(dF, tab)
    sq = card * card
    # back to the solver basis
    sqb = sq
    if tab.basisM == "Chebyshev":
        sqb = _apply(np.linalg.inv(tab.chebChi), sqb, 1)
    if tab.basisN == "Chebyshev":
        sqb = _apply(np.linalg.inv(tab.chebRz), sqb, 2)
        sqb = _apply(np.linalg.inv(tab.chebRp), sqb, 3)
    op, _, coll = assemble(pb, want_collision=True)
    src2 = coll @ sqb.reshape(-1)
    dNl = np.linalg.solve(op, src2).reshape(dF.shape)
    dNlCard = to_cardinal(dNl, tab)
    rq = row_quantities(pb)
    fE = feq(rq["E"] / rq["T"], rq["s"])
    pz = pb.pz[None, None, :, None]
    pp = pb.pp[None, None, None, :]
    integrand = (pb.dofs[:, None, None, None] * rq["dmsq"] * pb.dpzdrz[None, None, :, None]
                 * pb.dppdrp[None, None, None, :] * pp / (4 * np.pi**2 * rq["E"]))
    wchi = np.sqrt(1 - tab.chi**2) * (np.pi / tab.M)
    wz = np.sqrt(1 - tab.rz**2) * (np.pi / tab.N)
    wp = np.full(tab.N - 1, np.pi / (tab.N - 1))
    wp[0] /= 2
    wp = np.sqrt(1 - tab.rp**2) * wp
    W = wchi[None, :, None, None] * wz[None, None, :, None] * wp[None, None, None, :]
    pEq = np.sum(fE * integrand * W)
    pDf = np.sum(card * integrand * W)
    pNl = np.sum(dNlCard * integrand * W)
    return abs(pDf / pEq), abs(pNl / (pEq + pDf))


def feedback(Deltas, msq_interior, dofs, velocityMid, dmsq=None):
    """Moment feedback into the wall EOM: T30_out, T33_out per z (EOM.deltaToTmunu,
    equationOfMotion.py:2064-2125), potentialOut (:1545-1555), dVout (:1418-1429).
    Deltas [4,P,M-1] = (D00,D02,D20,D11); msq_interior [P,M-1]; dmsq [P,M-1,F] or None."""
    D00, D02, D20, D11 = Deltas
    u0 = np.sqrt(1.0 / (1.0 - velocityMid * velocityMid))
    u3 = u0 * velocityMid
    ub0, ub3 = u3, u0
    dof = np.asarray(dofs, dtype=float)[:, None]
    A1 = 3 * D20 - D02 - msq_interior * D00
    A2 = 3 * D02 - D20 + msq_interior * D00
    T30 = np.sum(dof * (A1 * u3 * u0 + A2 * ub3 * ub0 + 2 * D11 * (u3 * ub0 + ub3 * u0)) / 2.0, axis=0)
    T33 = np.sum(dof * ((A1 * u3 * u3 + A2 * ub3 * ub3 + 4 * D11 * u3 * ub3) / 2.0
                        - (msq_interior * D00 + D02 - D20) / 2.0), axis=0)
    pot = np.sum(dof * msq_interior * D00, axis=0) / 2
    dV = None if dmsq is None else np.sum(dof[:, :, None] * dmsq * D00[:, :, None], axis=0) / 2
    return T30, T33, pot, dV


# --------------------------------------------------------------------------
# Deterministic synthetic problems
# --------------------------------------------------------------------------
def synthetic_problem(P, M, N, *, vw_mid=-0.5, r=0.5, noise=1e-3, seed=0, T0=100.0,
                      wall_L=0.05, collisionMultiplier=1.0, basisM="Cardinal",
                      basisN="Chebyshev", derivatives="Spectral", grid3=True) -> Problem:
    tab = build_tables(M, N, basisM, basisN, derivatives)
    if grid3:
        xi, dxidchi = position_map_3scales(tab.chi, 0.2, 0.2, wall_L, 0.5, 0.1)
        xiF = np.concatenate([[-np.inf], xi, [np.inf]])
    else:
        xi, dxidchi = position_map_simple(tab.chi, wall_L)
        xiF = np.concatenate([[-np.inf], xi, [np.inf]])
    pz, pp, dpzdrz, dppdrp = momentum_maps(tab.rz, tab.rp, T0)
    th = np.tanh(xiF / wall_L)
    phi = 50 * (1 - th)
    T = 100 + th
    vwall = -0.5 + 0.02 * th
    msq = np.stack([0.5 * (1 + 0.3 * a) * phi**2 for a in range(P)])
    stats = np.array([(-1.0 if a % 2 == 0 else 1.0) for a in range(P)])
    rng = np.random.default_rng(seed)
    q = (N - 1) ** 2
    Ccard = (r * np.identity(P * q) + noise * rng.standard_normal((P * q, P * q))).reshape(
        P, N - 1, N - 1, P, N - 1, N - 1)
    coll = collision_to_chebyshev(Ccard, tab) if basisN == "Chebyshev" else Ccard
    return Problem(tab=tab, xi_dxidchi=dxidchi, pz=pz, pp=pp, dpzdrz=dpzdrz, dppdrp=dppdrp,
                   T=T, v=boost_velocity(vwall, vw_mid), msq=msq,
                   vw=float(boost_velocity(0.0, vw_mid)), statistics=stats, collision=coll,
                   collisionMultiplier=collisionMultiplier, dofs=np.full(P, 12.0))


def solve_extended(A: np.ndarray, b: np.ndarray, iterations: int = 4):
    """Solution of the FP64 system A x = b to (nearly) extended precision: LAPACK solve, then iterative
    refinement with the residual formed in np.longdouble (x87 80-bit here: 64-bit mantissa) and the
    correction carried as a separate low part.  Test infrastructure: the yardstick for wg_refine, which no FP64
    solver -- LAPACK included -- matches better than cond(A) * 2^-53.  Returns (x_hi + x_lo rounded, x_lapack)."""
    from scipy.linalg import lu_factor, lu_solve

    lu = lu_factor(A)
    x0 = lu_solve(lu, b)
    Al = A.astype(np.longdouble)
    bl = b.astype(np.longdouble)
    x = x0.astype(np.longdouble)
    for _ in range(iterations):
        r = bl - Al @ x
        x = x + lu_solve(lu, np.asarray(r, dtype=np.float64)).astype(np.longdouble)
    return np.asarray(x, dtype=np.float64), x0
