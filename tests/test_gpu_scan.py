"""Wall-velocity scan as the BASELINE metric defines a point: the unmodified reference's
`EOM.wallPressure(vw, wallParams0, None)` (equationOfMotion.py:786-1131) with the Boltzmann solves on the GPU.

  * solve server + forked host workers give the SAME rows, bit for bit, as the sequential in-process loop;
  * both agree with the reference's own BoltzmannSolver (host NumPy/LAPACK) on the same points;
  * the out-of-equilibrium T30/T33 that the vectorised plasma-profile solve takes from wg_feedback equal the
    reference's per-point EOM.deltaToTmunu.
Needs the reference copy under oracle/_ref."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rd():
    from oracle import ref_drivers
    if ref_drivers.reference_root() is None:
        pytest.skip("reference copy not present (oracle/fetch_ref.sh fills oracle/_ref in the build container)")
    return ref_drivers


@pytest.fixture(scope="module")
def yukawa(rd):
    WallGo = rd.load_reference()
    manager, settings, _ = rd.build_manager("yukawa", 16, 9)
    return WallGo, manager, settings


def test_scan_workers_equal_sequential_and_reference(cuda_device, rd, yukawa):
    import WallGo.manager as mgr
    from wallgo_b200 import dropin, vwscan

    WallGo, manager, settings = yukawa
    refCls = WallGo.boltzmann.BoltzmannSolver
    saved = mgr.BoltzmannSolver
    vws = np.linspace(0.15, 0.55, 7)
    cls = dropin.install(vectoriseEOM=False)
    origNew = rd.patch_load_collisions(cls)
    origRef = rd.patch_load_collisions(refCls)
    try:
        st0, st1 = {}, {}
        seq = vwscan.scanLocal(manager, settings, vws, range(len(vws)), workers=0, stats=st0)
        par = vwscan.scanLocal(manager, settings, vws, range(len(vws)), workers=3, concurrency=2, stats=st1)
        assert st0["solves"] == st1["solves"] > 2 * len(vws)
        assert np.array_equal(seq, par), "solve server + host workers must reproduce the sequential loop bit for bit"
        # the reference itself on three of the points
        mgr.BoltzmannSolver = refCls
        ws = manager.setupWallSolver(settings)
        ws.eom.pressAbsErrTol = 1e-8
        wp0 = WallGo.WallParams(widths=ws.initialWallThickness * np.ones(ws.eom.nbrFields),
                                offsets=np.zeros(ws.eom.nbrFields))
        for i in (0, 3, 6):
            out = ws.eom.wallPressure(float(vws[i]), wp0, boltzmannResultsInput=None)
            ref = vwscan._row(out, seq[i, 1 + 2 * ws.eom.nbrFields])
            scale = np.maximum(np.abs(ref), 1e-300)
            assert abs(seq[i, 0] - ref[0]) <= 1e-6 * abs(ref[0]), (i, seq[i, 0], ref[0])
            nF = ws.eom.nbrFields
            assert np.allclose(seq[i, 1:1 + 2 * nF], ref[1:1 + 2 * nF], rtol=1e-6, atol=1e-12)
            D, Dr = seq[i, 2 + 2 * nF:], ref[2 + 2 * nF:]
            assert np.max(np.abs(D - Dr)) <= 1e-6 * np.max(np.abs(Dr))
            del scale
    finally:
        mgr.BoltzmannSolver = saved
        refCls.loadCollisions = origRef
        if origNew is None:
            del cls.loadCollisions
        else:
            cls.loadCollisions = origNew


def test_vectorised_profile_uses_device_feedback(cuda_device, rd, yukawa):
    """eomhost.findPlasmaProfile with T30/T33 from wg_feedback == the reference's findPlasmaProfile (which
    calls deltaToTmunu point by point) inside a real wallPressure; and the kernel == deltaToTmunu directly."""
    import WallGo.manager as mgr
    from wallgo_b200 import dropin, eomhost

    WallGo, manager, settings = yukawa
    saved = mgr.BoltzmannSolver
    cls = dropin.install(vectoriseEOM=False)
    origNew = rd.patch_load_collisions(cls)
    try:
        ws = manager.setupWallSolver(settings)
        eom = ws.eom
        eom.pressAbsErrTol = 1e-8
        wp0 = WallGo.WallParams(widths=ws.initialWallThickness * np.ones(eom.nbrFields),
                                offsets=np.zeros(eom.nbrFields))
        plain = eom.wallPressure(0.3, wp0, boltzmannResultsInput=None)
        res, bg = plain[2], plain[3]
        # kernel vs the reference routine on the converged moments
        fields = bg.fieldProfiles[1:-1] if hasattr(bg.fieldProfiles, "__getitem__") else bg.fieldProfiles
        fieldsInt = WallGo.Fields(np.asarray(bg.fieldProfiles)[1:-1])
        t30, t33 = ws.boltzmannSolver.tmunuOut(res.Deltas, fieldsInt, bg.velocityMid)
        refT = np.array([eom.deltaToTmunu(i, fieldsInt.getFieldPoint(i), bg.velocityMid, res.Deltas)
                         for i in range(eom.grid.M - 1)])
        assert np.max(np.abs(t30 - refT[:, 0])) <= 1e-12 * np.max(np.abs(refT[:, 0]))
        assert np.max(np.abs(t33 - refT[:, 1])) <= 1e-12 * np.max(np.abs(refT[:, 1]))
        h30, h33 = eomhost.deltaToTmunuVec(eom.particles, fieldsInt, bg.velocityMid, res.Deltas)
        assert np.max(np.abs(h30 - refT[:, 0])) <= 1e-13 * np.max(np.abs(refT[:, 0]))
        assert np.max(np.abs(h33 - refT[:, 1])) <= 1e-13 * np.max(np.abs(refT[:, 1]))
        del fields
        # the vectorised profile solve (device feedback) inside the same wallPressure
        eomhost.install(WallGo.EOM)
        try:
            vec = eom.wallPressure(0.3, wp0, boltzmannResultsInput=None)
        finally:
            eomhost.uninstall(WallGo.EOM)
        assert abs(vec[0] - plain[0]) <= 1e-6 * abs(plain[0])
        assert np.allclose(vec[1].widths, plain[1].widths, rtol=1e-6)
        assert np.allclose(vec[3].temperatureProfile, plain[3].temperatureProfile, rtol=1e-7)
    finally:
        mgr.BoltzmannSolver = saved
        if origNew is None:
            del cls.loadCollisions
        else:
            cls.loadCollisions = origNew
