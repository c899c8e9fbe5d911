"""Host-side widening of the hot path's consumer: the plasma-profile solve
of the wall equation of motion, vectorised over the grid points.

The reference solves Eq. (20) of arXiv:2204.13120 point by point (`EOM.findPlasmaProfile` ->
`findPlasmaProfilePoint`, equationOfMotion.py:1764-1941): for each of the M-1 interior z nodes one bounded
scalar minimisation (`scipy.optimize.minimize_scalar(method="Bounded")`), a geometric bracket search and one
`brentq` root find, every function value being a call into the user's effective potential with a scalar
temperature.  After the Boltzmann solve has moved to the GPU this loop is the largest remaining cost of a
`solveWall`.  The effective potential is user Python, so it cannot move to the
device; but it accepts arrays, and the M-1 problems are independent.  Here all points advance in lock step:

  * `minimizeBoundedVec` restates SciPy's bounded Brent minimiser (scipy/optimize/_optimize.py,
    `_minimize_scalar_bounded`) element-wise, iterate for iterate;
  * `brentqVec` restates SciPy's `brentq` (scipy/optimize/Zeros/brentq.c) element-wise, iterate for iterate;
  * `findPlasmaProfile` is the reference's control flow (minimum -> no-root fallback -> bracket -> root ->
    plasma velocity -> previous-point fallback) on vectors.

An element that has converged keeps its value while the others continue, so each point sees exactly the
sequence of temperatures the scalar code would feed it; results differ from the reference only through
last-bit differences between NumPy's scalar and array code paths inside the potential.

Opt-in: `wallgo_b200.dropin.install(vectoriseEOM=True)` rebinds `WallGo.EOM.findPlasmaProfile`.
"""
from __future__ import annotations

import sys

import numpy as np


def minimizeBoundedVec(func, lower, upper, xatol: float = 1e-5, maxiter: int = 500):
    """Element-wise bounded scalar minimisation; `func` maps an array of abscissae (one per problem) to the
    array of function values.  Returns (x, f(x), nfev per element).  Follows `_minimize_scalar_bounded`."""
    a = np.array(lower, dtype=np.float64, copy=True)
    b = np.array(upper, dtype=np.float64, copy=True)
    if a.shape != b.shape:
        a, b = np.broadcast_arrays(a, b)
        a, b = a.copy(), b.copy()
    sqrtEps = np.sqrt(2.2e-16)
    goldenMean = 0.5 * (3.0 - np.sqrt(5.0))
    fulc = a + goldenMean * (b - a)
    nfc, xf = fulc.copy(), fulc.copy()
    rat = np.zeros_like(a)
    e = np.zeros_like(a)
    x = xf.copy()
    fx = np.asarray(func(x), dtype=np.float64).copy()
    num = np.ones(a.shape, dtype=np.int64)
    ffulc, fnfc = fx.copy(), fx.copy()
    xm = 0.5 * (a + b)
    tol1 = sqrtEps * np.abs(xf) + xatol / 3.0
    tol2 = 2.0 * tol1
    active = np.abs(xf - xm) > (tol2 - 0.5 * (b - a))
    while np.any(active):
        golden = np.ones(a.shape, dtype=bool)
        tryPar = active & (np.abs(e) > tol1)
        if np.any(tryPar):
            with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
                r = (xf - nfc) * (fx - ffulc)
                q = (xf - fulc) * (fx - fnfc)
                p = (xf - fulc) * q - (xf - nfc) * r
                q = 2.0 * (q - r)
                p = np.where(q > 0.0, -p, p)
                q = np.abs(q)
                r = e.copy()
                e = np.where(tryPar, rat, e)
                ok = tryPar & (np.abs(p) < np.abs(0.5 * q * r)) & (p > q * (a - xf)) & (p < q * (b - xf))
                ratPar = (p + 0.0) / q
                xPar = xf + ratPar
                near = ((xPar - a) < tol2) | ((b - xPar) < tol2)
                si = np.sign(xm - xf) + ((xm - xf) == 0)
                ratPar = np.where(near, tol1 * si, ratPar)
            rat = np.where(ok, ratPar, rat)
            golden = ~ok
        g = active & golden
        if np.any(g):
            eG = np.where(xf >= xm, a - xf, b - xf)
            e = np.where(g, eG, e)
            rat = np.where(g, goldenMean * eG, rat)
        si = np.sign(rat) + (rat == 0)
        xNew = xf + si * np.maximum(np.abs(rat), tol1)
        x = np.where(active, xNew, x)
        fuAll = np.asarray(func(x), dtype=np.float64)
        fu = fuAll
        num = num + active
        better = active & (fu <= fx)
        worse = active & ~(fu <= fx)
        # fu <= fx: the new point becomes the best
        aB = np.where(better & (x >= xf), xf, a)
        bB = np.where(better & ~(x >= xf), xf, b)
        # fu > fx: the new point only shrinks the bracket
        aW = np.where(worse & (x < xf), x, aB)
        bW = np.where(worse & ~(x < xf), x, bB)
        c1 = worse & ((fu <= fnfc) | (nfc == xf))
        c2 = worse & ~c1 & ((fu <= ffulc) | (fulc == xf) | (fulc == nfc))
        newFulc = np.where(better | c1, nfc, np.where(c2, x, fulc))
        newFfulc = np.where(better | c1, fnfc, np.where(c2, fu, ffulc))
        newNfc = np.where(better, xf, np.where(c1, x, nfc))
        newFnfc = np.where(better, fx, np.where(c1, fu, fnfc))
        newXf = np.where(better, x, xf)
        newFx = np.where(better, fu, fx)
        a, b = aW, bW
        fulc, ffulc, nfc, fnfc, xf, fx = newFulc, newFfulc, newNfc, newFnfc, newXf, newFx
        xm = np.where(active, 0.5 * (a + b), xm)
        tol1 = np.where(active, sqrtEps * np.abs(xf) + xatol / 3.0, tol1)
        tol2 = 2.0 * tol1
        active = active & (num < maxiter) & (np.abs(xf - xm) > (tol2 - 0.5 * (b - a)))
    return xf, fx, num


def brentqVec(func, xa, xb, xtol: float = 2e-12, rtol: float = 8.881784197001252e-16, maxiter: int = 100,
              mask=None):
    """Element-wise Brent root finding on brackets [xa, xb]; `func` maps an array of abscissae to the array of
    function values.  Elements with mask == False are ignored (returned as NaN).  Returns (root, converged).
    Follows scipy/optimize/Zeros/brentq.c statement by statement."""
    xpre = np.array(xa, dtype=np.float64, copy=True)
    xcur = np.array(xb, dtype=np.float64, copy=True)
    live = np.ones(xpre.shape, dtype=bool) if mask is None else np.array(mask, dtype=bool, copy=True)
    xblk = np.zeros_like(xpre)
    fblk = np.zeros_like(xpre)
    spre = np.zeros_like(xpre)
    scur = np.zeros_like(xpre)
    # dead elements are evaluated at a harmless abscissa (their own xa) and never updated
    fpre = np.asarray(func(xpre), dtype=np.float64).copy()
    fcur = np.asarray(func(xcur), dtype=np.float64).copy()
    root = np.full(xpre.shape, np.nan)
    done = ~live
    converged = np.zeros(xpre.shape, dtype=bool)
    z = live & (fpre == 0)
    root = np.where(z, xpre, root)
    done |= z
    converged |= z
    z = live & ~done & (fcur == 0)
    root = np.where(z, xcur, root)
    done |= z
    converged |= z
    bad = live & ~done & (np.signbit(fpre) == np.signbit(fcur))
    if np.any(bad):
        raise ValueError("f(a) and f(b) must have different signs")
    for _ in range(maxiter):
        act = ~done
        if not np.any(act):
            break
        with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
            s = act & (fpre != 0) & (fcur != 0) & (np.signbit(fpre) != np.signbit(fcur))
            xblk = np.where(s, xpre, xblk)
            fblk = np.where(s, fpre, fblk)
            d = xcur - xpre
            spre = np.where(s, d, spre)
            scur = np.where(s, d, scur)
            sw = act & (np.abs(fblk) < np.abs(fcur))
            xpre, xcur, xblk = np.where(sw, xcur, xpre), np.where(sw, xblk, xcur), np.where(sw, xcur, xblk)
            fpre, fcur, fblk = np.where(sw, fcur, fpre), np.where(sw, fblk, fcur), np.where(sw, fcur, fblk)
            delta = (xtol + rtol * np.abs(xcur)) / 2
            sbis = (xblk - xcur) / 2
            fin = act & ((fcur == 0) | (np.abs(sbis) < delta))
            root = np.where(fin, xcur, root)
            done |= fin
            converged |= fin
            act = act & ~fin
            if not np.any(act):
                break
            tryI = act & (np.abs(spre) > delta) & (np.abs(fcur) < np.abs(fpre))
            interp = -fcur * (xcur - xpre) / (fcur - fpre)
            dpre = (fpre - fcur) / (xpre - xcur)
            dblk = (fblk - fcur) / (xblk - xcur)
            extrap = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre))
            stry = np.where(xpre == xblk, interp, extrap)
            good = tryI & (2 * np.abs(stry) < np.minimum(np.abs(spre), 3 * np.abs(sbis) - delta))
            newSpre = np.where(good, scur, sbis)
            newScur = np.where(good, stry, sbis)
            spre = np.where(act, newSpre, spre)
            scur = np.where(act, newScur, scur)
            xpre = np.where(act, xcur, xpre)
            fpre = np.where(act, fcur, fpre)
            step = np.where(np.abs(scur) > delta, scur, np.where(sbis > 0, delta, -delta))
            xcur = np.where(act, xcur + step, xcur)
        fnew = np.asarray(func(np.where(act, xcur, xpre)), dtype=np.float64)
        fcur = np.where(act, fnew, fcur)
    left = ~done
    root = np.where(left, xcur, root)     # brentq.c returns xcur with CONVERR after `iter` iterations
    return root, converged


class _ProfileProblem:
    """Left-hand side of Eq. (20) of arXiv:2204.13120 on all grid points at once (the vector form of
    EOM.temperatureProfileEqLHS, equationOfMotion.py:1968-2019)."""

    _flatPotentials: set = set()     # potential classes known to need the flattened temperature derivative

    def __init__(self, eom, fields, dPhidz, s1, s2):
        self.veff = eom.thermo.effectivePotential
        self._flatDerivT = type(self.veff) in self._flatPotentials
        self.helpers = sys.modules[type(eom).__module__.split(".")[0]].helpers     # WallGo.helpers of this EOM
        self.fields = fields
        self.kinetic = 0.5 * np.sum(np.asarray(dPhidz) ** 2, axis=1)
        self.s1 = np.asarray(s1, dtype=np.float64)
        self.s2 = np.asarray(s2, dtype=np.float64)

    def enthalpy(self, T):
        return -T * self._derivT(T)

    def _derivT(self, T):
        """dV/dT on all points.  EffectivePotential.derivT (effectivePotential.py:193-221) hands the potential a
        (stencil, points) temperature array; user potentials written for one field point at a time cannot always
        broadcast that against (points,) fields (InertDoublet example: inertDoubletModel.py:583-585).  For those
        the same finite-difference formula (helpers.derivative, helpers.py:74-185: same step, stencil, one-sided
        offsets and coefficients) is evaluated on FLATTENED (stencil*points,) arrays, which every potential that
        works element-wise accepts."""
        if not self._flatDerivT:
            try:
                return np.asarray(self.veff.derivT(self.fields, T), dtype=np.float64).reshape(T.shape)
            except (ValueError, TypeError, IndexError):
                self._flatDerivT = True
                type(self)._flatPotentials.add(type(self.veff))
        H = self.helpers
        eps = self.veff.effectivePotentialError
        scale = self.veff.derivativeSettings.temperatureVariationScale
        x = np.asarray(T, dtype=np.float64)
        dx = scale * eps ** (1.0 / 5.0)
        dx = (x + dx) - x                                   # exactly representable steps, per point
        offset = np.zeros(x.shape, dtype=int)
        offset += x - dx < 0.0
        offset += x - 2 * dx < 0.0                          # (the upper bound is +inf: no offsets on that side)
        pos = x[None, :] + H.FIRST_DERIV_POS["4"].T[:, offset.tolist()] * dx
        coeff = H.FIRST_DERIV_COEFF["4"].T[:, offset.tolist()] / dx
        S = pos.shape[0]
        fieldsRep = np.tile(np.asarray(self.fields), (S, 1)).view(type(self.fields))
        fx = np.asarray(self.veff.evaluate(fieldsRep, pos.reshape(-1)), dtype=np.float64).reshape(S, -1)
        return np.sum(coeff * fx, axis=0)

    def lhs(self, T):
        T = np.asarray(T, dtype=np.float64)
        enthalpy = self.enthalpy(T)
        veff = np.asarray(self.veff.evaluate(self.fields, T)).reshape(T.shape)
        return self.kinetic - veff - 0.5 * enthalpy + 0.5 * np.sqrt(4 * self.s1**2 + enthalpy**2) - self.s2

    def plasmaVelocity(self, T):
        enthalpy = self.enthalpy(np.asarray(T, dtype=np.float64))
        return (-enthalpy + np.sqrt(4 * self.s1**2 + enthalpy**2)) / (2 * self.s1)


def deltaToTmunuVec(particles, fields, velocityMid, offEquilDeltas):
    """EOM.deltaToTmunu (equationOfMotion.py:2021-2127) on all grid points at once, host NumPy: returns
    (T30_out, T33_out) with one entry per point of `fields`.  Same arithmetic order as the scalar routine."""
    u0 = np.sqrt(1.0 / (1.0 - velocityMid * velocityMid))
    u3 = u0 * velocityMid
    ub0, ub3 = u3, u0
    npts = np.asarray(offEquilDeltas.Delta00.coefficients).shape[1]
    t30 = np.zeros(npts)
    t33 = np.zeros(npts)
    for i, particle in enumerate(particles):
        d00 = np.asarray(offEquilDeltas.Delta00.coefficients)[i]
        d02 = np.asarray(offEquilDeltas.Delta02.coefficients)[i]
        d20 = np.asarray(offEquilDeltas.Delta20.coefficients)[i]
        d11 = np.asarray(offEquilDeltas.Delta11.coefficients)[i]
        msq = np.asarray(particle.msqVacuum(fields), dtype=np.float64).reshape(npts)
        a1 = 3 * d20 - d02 - msq * d00
        a2 = 3 * d02 - d20 + msq * d00
        t30 = t30 + particle.totalDOFs * (a1 * u3 * u0 + a2 * ub3 * ub0 + 2 * d11 * (u3 * ub0 + ub3 * u0)) / 2.0
        t33 = t33 + particle.totalDOFs * ((a1 * u3 * u3 + a2 * ub3 * ub3 + 4 * d11 * u3 * ub3) / 2.0
                                          - (msq * d00 + d02 - d20) / 2.0)
    return t30, t33


def _tmunuOut(eom, fields, velocityMid, offEquilDeltas):
    """Out-of-equilibrium T30, T33 for every interior z: the device kernel (wg_feedback, one launch) when the
    EOM's Boltzmann solver owns a GPU engine, the host vector form in the scan's host workers."""
    solver = getattr(eom, "boltzmannSolver", None)
    if solver is not None and getattr(solver, "_engine", None) is not None and hasattr(solver, "tmunuOut"):
        return solver.tmunuOut(offEquilDeltas, fields, velocityMid)
    return deltaToTmunuVec(eom.particles, fields, velocityMid, offEquilDeltas)


def findPlasmaProfile(eom, c1, c2, velocityMid, fields, dPhidz, offEquilDeltas, Tplus, Tminus):
    """Vector form of EOM.findPlasmaProfile (equationOfMotion.py:1764-1941); same arguments after `eom`,
    same return value (temperatureProfile, velocityProfile) and the same success flags on `eom`."""
    npts = len(eom.grid.xiValues)
    # out-of-equilibrium part of T^{30}, T^{33} on all points at once (the reference: M-1 calls of deltaToTmunu)
    t30, t33 = _tmunuOut(eom, fields, velocityMid, offEquilDeltas)
    s1 = c1 - t30
    s2 = c2 - t33
    pb = _ProfileProblem(eom, fields, dPhidz, s1, s2)
    upper = 2 * max(Tplus, Tminus)
    tMin, _, _ = minimizeBoundedVec(pb.lhs, np.zeros(npts), np.full(npts, upper))
    fMin = pb.lhs(tMin)
    noRoot = fMin >= 0                       # positive minimum: no root, return the minimum's position
    # ---- bracket the root on the requested side of the minimum
    with np.errstate(divide="ignore", invalid="ignore"):
        mult = np.maximum(Tplus / tMin, 1.2)
        if abs(eom.hydrodynamics.Tnucl - Tplus) < 1e-10:      # detonation: the solution below the minimum
            mult = np.minimum(Tminus / tMin, 0.8)
    tLow = tMin.copy()
    tTest = tMin * mult
    search = ~noRoot
    failed = np.zeros(npts, dtype=bool)
    count = np.zeros(npts, dtype=np.int64)
    while True:
        ft = pb.lhs(np.where(search, tTest, tMin))
        still = search & (ft < 0)
        if not np.any(still):
            break
        giveUp = still & (count > 100)
        failed |= giveUp
        search &= ~giveUp
        still &= ~giveUp
        tLow = np.where(still, tLow * mult, tLow)
        tTest = np.where(still, tTest * mult, tTest)
        count = count + still
        if not np.any(still):
            break
    solve = ~noRoot & ~failed
    T = np.where(noRoot, tMin, 0.0)
    if np.any(solve):
        safeLow = np.where(solve, tLow, tMin)
        safeHigh = np.where(solve, tTest, tMin)
        root, _ = brentqVec(pb.lhs, safeLow, safeHigh, xtol=1e-10, rtol=eom.errTol / 10, maxiter=100, mask=solve)
        T = np.where(solve, root, T)
    good = T > 0
    v = np.zeros(npts)
    if np.any(good):
        v = np.where(good, pb.plasmaVelocity(np.where(good, T, tMin)), 0.0)
    temperatureProfile = np.zeros(npts)
    velocityProfile = np.zeros(npts)
    eom.successTemperatureProfile = True
    for i in range(npts):                    # the reference's sequential fallback to the previous point
        if good[i]:
            temperatureProfile[i] = T[i]
            velocityProfile[i] = v[i]
            if noRoot[i]:
                eom.successTemperatureProfile = False
        else:
            temperatureProfile[i] = temperatureProfile[i - 1]
            velocityProfile[i] = velocityProfile[i - 1]
            eom.successTemperatureProfile = False
    eom.successPlasmaProfilePoint = bool(npts == 0 or (good[-1] and not noRoot[-1]))
    return temperatureProfile, velocityProfile


def violationOfEMConservation(eom, c1, c2, velocityMid, fields, dPhidz, offEquilDeltas, temperatureProfile,
                              velocityProfile, wallThickness):
    """Vector form of EOM.violationOfEMConservation (equationOfMotion.py:2145-2228) and of the violationEMPoint
    (:2230-2293) it calls once per grid point: the residuals of the energy-momentum equations on all points at
    once (same expressions), then the reference's own Polynomial quadrature."""
    WallGo = sys.modules[type(eom).__module__.split(".")[0]]
    t30, t33 = _tmunuOut(eom, fields, velocityMid, offEquilDeltas)
    T = np.asarray(temperatureProfile, dtype=np.float64)
    v = np.asarray(velocityProfile, dtype=np.float64)
    pb = _ProfileProblem(eom, fields, dPhidz, np.zeros(T.shape), np.zeros(T.shape))
    enthalpy = pb.enthalpy(T)
    veff = np.asarray(pb.veff.evaluate(fields, T), dtype=np.float64).reshape(T.shape)
    violationEM30 = (enthalpy * v / (1 - v**2) + t30 - c1) / c1
    violationEM33 = (pb.kinetic - veff + enthalpy * v**2 / (1 - v**2) + t33 - c2) / c2
    T30Poly = WallGo.Polynomial(violationEM30**2, eom.grid)
    T33Poly = WallGo.Polynomial(violationEM33**2, eom.grid)
    dzdchi, _, _ = eom.grid.getCompactificationDerivatives()
    return (np.sqrt(T30Poly.integrate(weight=dzdchi)) / wallThickness,
            np.sqrt(T33Poly.integrate(weight=dzdchi)) / wallThickness)


def install(eomClass):
    """Rebind `findPlasmaProfile` and `violationOfEMConservation` of the given EOM class (WallGo.EOM) to their
    vector forms."""
    def _violationVec(self, *args):
        if not getattr(self, "_profileVectorUnsupported", False):
            try:
                return violationOfEMConservation(self, *args)
            except (ValueError, TypeError, IndexError):
                pass
        return eomClass._violationOfEMConservationReference(self, *args)
    _violationVec.__doc__ = violationOfEMConservation.__doc__
    eomClass._violationOfEMConservationReference = eomClass.violationOfEMConservation
    eomClass.violationOfEMConservation = _violationVec

    def _vectorised(self, c1, c2, velocityMid, fields, dPhidz, offEquilDeltas, Tplus, Tminus):
        # The lock-step form needs an effective potential that accepts one temperature PER field point (element-wise
        # (points,) arrays; potentials that cannot take the (stencil, points) array of derivT are served by the
        # flattened derivative of _ProfileProblem._derivT).  If a user potential cannot even do that, the
        # reference's own per-point routine runs.
        if not getattr(self, "_profileVectorUnsupported", False):
            try:
                return findPlasmaProfile(self, c1, c2, velocityMid, fields, dPhidz, offEquilDeltas, Tplus, Tminus)
            except (ValueError, TypeError, IndexError):
                self._profileVectorUnsupported = True
        return eomClass._findPlasmaProfileReference(self, c1, c2, velocityMid, fields, dPhidz, offEquilDeltas,
                                                    Tplus, Tminus)
    _vectorised.__doc__ = findPlasmaProfile.__doc__
    eomClass._findPlasmaProfileReference = eomClass.findPlasmaProfile
    eomClass.findPlasmaProfile = _vectorised
    return eomClass


def uninstall(eomClass):
    """Undo install(): the reference's own routines are bound again."""
    if hasattr(eomClass, "_findPlasmaProfileReference"):
        eomClass.findPlasmaProfile = eomClass._findPlasmaProfileReference
        del eomClass._findPlasmaProfileReference
    if hasattr(eomClass, "_violationOfEMConservationReference"):
        eomClass.violationOfEMConservation = eomClass._violationOfEMConservationReference
        del eomClass._violationOfEMConservationReference
