"""Vectorised plasma-profile solve.  The two element-wise solvers are pinned to
SciPy's own scalar routines iterate for iterate (hermetic: SciPy only); the EOM routine is pinned to the
reference's `findPlasmaProfile` inside a real `solveWall` of the shipped Yukawa model (build container only)."""
import os
import sys
import types

import numpy as np
import pytest
import scipy.optimize

from conftest import ROOT
from wallgo_b200 import eomhost

REF = "/root/reference/src"


def _family(rng, n):
    """Parabola-like functions with two roots, the shape of Eq. (20): a (x-c)^2 + b (x-c)^4 - d."""
    a = rng.uniform(0.5, 3.0, n)
    b = rng.uniform(0.0, 0.5, n)
    c = rng.uniform(5.0, 90.0, n)
    d = rng.uniform(0.1, 50.0, n)
    vec = lambda x: a * (x - c) ** 2 + b * (x - c) ** 4 - d + np.log1p(x) * 1e-3   # noqa: E731
    # the scalar problems go through the same array arithmetic, so both sides see bit-identical function values
    scal = [(lambda x, i=i: float(vec(np.full(n, float(x)))[i])) for i in range(n)]
    return scal, vec, c


def test_bounded_minimiser_follows_scipy_iterate_for_iterate():
    rng = np.random.default_rng(0)
    n = 40
    scal, vec, _ = _family(rng, n)
    x, fx, nfev = eomhost.minimizeBoundedVec(vec, np.zeros(n), np.full(n, 200.0))
    for i in range(n):
        ref = scipy.optimize.minimize_scalar(scal[i], method="Bounded", bounds=[0, 200.0])
        assert ref.nfev == nfev[i]
        assert ref.x == x[i], (i, ref.x, x[i])
        assert ref.fun == fx[i]


def test_bounded_minimiser_edge_cases():
    # minimum at a bound, flat function, and different brackets per element
    vec = lambda x: np.array([x[0], -x[1], 0.0 * x[2] + 1.0, (x[3] - 3.0) ** 2])   # noqa: E731
    lo = np.array([0.0, 0.0, 1.0, 2.5])
    hi = np.array([1.0, 2.0, 4.0, 3.25])
    x, fx, nfev = eomhost.minimizeBoundedVec(vec, lo, hi)
    fs = [lambda t: t, lambda t: -t, lambda t: 0.0 * t + 1.0, lambda t: (t - 3.0) ** 2]
    for i in range(4):
        ref = scipy.optimize.minimize_scalar(fs[i], method="Bounded", bounds=[lo[i], hi[i]])
        assert ref.x == x[i] and ref.nfev == nfev[i]


@pytest.mark.parametrize("rtol", [1e-4, 4 * np.finfo(float).eps])
def test_brentq_follows_scipy_iterate_for_iterate(rtol):
    rng = np.random.default_rng(1)
    n = 40
    scal, vec, c = _family(rng, n)
    lo, hi = c.copy(), c + 60.0
    root, conv = eomhost.brentqVec(vec, lo, hi, xtol=1e-10, rtol=rtol, maxiter=100)
    assert np.all(conv)
    for i in range(n):
        ref, info = scipy.optimize.brentq(scal[i], lo[i], hi[i], xtol=1e-10, rtol=rtol, maxiter=100, full_output=True)
        assert ref == root[i], (i, ref, root[i])
    # a masked-out element is left alone, and a root exactly at an end point is returned as is
    root2, _ = eomhost.brentqVec(lambda x: x - 2.0, np.array([2.0, 0.0, 1.0]), np.array([5.0, 1.0, 4.0]),
                                 mask=np.array([True, False, True]))
    assert root2[0] == 2.0 and np.isnan(root2[1]) and abs(root2[2] - 2.0) < 1e-11
    with pytest.raises(ValueError):
        eomhost.brentqVec(lambda x: x + 1.0, np.array([0.0]), np.array([1.0]))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not present on this machine")
def test_findPlasmaProfile_matches_reference_inside_solveWall():
    """Every findPlasmaProfile call of a reference solveWall (Yukawa model, off-equilibrium on) is answered by
    both implementations: temperature and velocity profiles agree to 1e-8, flags are equal; then the same run
    with the vector form installed reproduces the wall velocity to 1e-6 (north-star tolerance)."""
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_trace
    import WallGo

    ref = WallGo.boltzmann.BoltzmannSolver
    origLoad = ref.loadCollisions
    origFind = WallGo.EOM.findPlasmaProfile
    worst = {"T": 0.0, "v": 0.0, "calls": 0}

    origViol = WallGo.EOM.violationOfEMConservation
    worst.update(viol=0.0, violCalls=0)

    def both(self, *args):
        Tr, vr = origFind(self, *args)
        flagR = self.successTemperatureProfile
        Tv, vv = eomhost.findPlasmaProfile(self, *args)
        assert self.successTemperatureProfile == flagR
        worst["T"] = max(worst["T"], float(np.max(np.abs(Tv - Tr) / np.abs(Tr))))
        worst["v"] = max(worst["v"], float(np.max(np.abs(vv - vr) / np.maximum(np.abs(vr), 1e-300))))
        worst["calls"] += 1
        return Tr, vr

    def bothViol(self, *args):
        r = origViol(self, *args)
        g = eomhost.violationOfEMConservation(self, *args)
        worst["viol"] = max(worst["viol"], abs(g[0] - r[0]) / abs(r[0]), abs(g[1] - r[1]) / abs(r[1]))
        worst["violCalls"] += 1
        return r

    ref.loadCollisions = lambda self, path: self.setCollisionArray(make_trace.rta_collisions(self))
    try:
        WallGo.EOM.findPlasmaProfile = both
        WallGo.EOM.violationOfEMConservation = bothViol
        manager, settings = make_trace.build_manager(8, 5)
        resRef = manager.solveWall(settings)
        assert worst["calls"] > 20 and worst["violCalls"] > 5
        assert worst["T"] < 1e-8 and worst["v"] < 1e-8, worst
        assert worst["viol"] < 1e-10, worst            # the energy-momentum residuals on all points at once
        WallGo.EOM.findPlasmaProfile = origFind
        WallGo.EOM.violationOfEMConservation = origViol
        eomhost.install(WallGo.EOM)
        manager, settings = make_trace.build_manager(8, 5)
        resVec = manager.solveWall(settings)
    finally:
        eomhost.uninstall(WallGo.EOM)
        WallGo.EOM.findPlasmaProfile = origFind
        WallGo.EOM.violationOfEMConservation = origViol
        ref.loadCollisions = origLoad
    assert resVec.success == resRef.success
    assert abs(resVec.wallVelocity - resRef.wallVelocity) <= 1e-6 * abs(resRef.wallVelocity)
    assert np.allclose(resVec.wallWidths, resRef.wallWidths, rtol=1e-6, atol=0)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not present on this machine")
def test_flattened_temperature_derivative_for_point_wise_potentials():
    """A user potential written for ONE field point at a time (the InertDoublet example: its resummed masses only
    broadcast a scalar temperature, inertDoubletModel.py:583-585) cannot take the (stencil, points) temperature
    array of EffectivePotential.derivT.  The vector plasma-profile solve then evaluates the same finite-difference
    formula on flattened arrays: it must equal the reference's derivT called point by point, bit for bit, including
    the one-sided stencils next to T = 0."""
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import WallGo

    class PointWisePotential(WallGo.EffectivePotential):
        fieldCount = 1
        effectivePotentialError = 1e-12

        def evaluate(self, fields, temperature):
            fields = WallGo.Fields(fields)
            v = fields.getField(0)
            T = np.asarray(temperature, dtype=float)
            thermal = 0.1 * T ** 2 * v ** 2 - 0.3 * T ** 4 + np.log1p(T) * v
            tree = -50.0 * v ** 2 + 0.02 * v ** 4
            if tree.shape != thermal.shape:                      # the same "HACK" as the InertDoublet example
                tree = tree * np.ones(thermal.shape[0])
            return np.array(tree + thermal)

    veff = PointWisePotential()
    veff.configureDerivatives(WallGo.VeffDerivativeSettings(temperatureVariationScale=1.0, fieldValueVariationScale=[10.0]))
    npts = 9
    fields = WallGo.Fields(np.linspace(0.0, 60.0, npts)[:, None])
    T = np.array([80.0, 75.0, 1e-3, 3e-3, 60.0, 55.0, 50.0, 45.0, 2e-4])      # some next to the lower bound
    with pytest.raises(ValueError):
        veff.derivT(fields, T)                                   # the (4, 9) temperature array does not broadcast

    class FakeEOM:
        pass
    eom = FakeEOM()
    eom.__class__.__module__ = "WallGo.equationOfMotion"
    eom.thermo = types.SimpleNamespace(effectivePotential=veff)
    eomhost._ProfileProblem._flatPotentials.discard(PointWisePotential)
    pb = eomhost._ProfileProblem(eom, fields, np.zeros((npts, 1)), np.ones(npts), np.ones(npts))
    got = pb._derivT(T)
    want = np.array([float(veff.derivT(fields.getFieldPoint(i), T[i])) for i in range(npts)])
    assert np.array_equal(got, want)
    assert PointWisePotential in eomhost._ProfileProblem._flatPotentials     # remembered: no second failed attempt
    lhs = pb.lhs(T)
    assert lhs.shape == (npts,) and np.all(np.isfinite(lhs))
