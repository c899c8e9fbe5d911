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

    def both(self, *args):
        Tr, vr = origFind(self, *args)
        flagR = self.successTemperatureProfile
        Tv, vv = eomhost.findPlasmaProfile(self, *args)
        assert self.successTemperatureProfile == flagR
        worst["T"] = max(worst["T"], float(np.max(np.abs(Tv - Tr) / np.abs(Tr))))
        worst["v"] = max(worst["v"], float(np.max(np.abs(vv - vr) / np.maximum(np.abs(vr), 1e-300))))
        worst["calls"] += 1
        return Tr, vr

    ref.loadCollisions = lambda self, path: self.setCollisionArray(make_trace.rta_collisions(self))
    try:
        WallGo.EOM.findPlasmaProfile = both
        manager, settings = make_trace.build_manager(8, 5)
        resRef = manager.solveWall(settings)
        assert worst["calls"] > 20
        assert worst["T"] < 1e-8 and worst["v"] < 1e-8, worst
        WallGo.EOM.findPlasmaProfile = origFind
        eomhost.install(WallGo.EOM)
        manager, settings = make_trace.build_manager(8, 5)
        resVec = manager.solveWall(settings)
    finally:
        WallGo.EOM.findPlasmaProfile = origFind
        ref.loadCollisions = origLoad
    assert resVec.success == resRef.success
    assert abs(resVec.wallVelocity - resRef.wallVelocity) <= 1e-6 * abs(resRef.wallVelocity)
    assert np.allclose(resVec.wallWidths, resRef.wallWidths, rtol=1e-6, atol=0)
