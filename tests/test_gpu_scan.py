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
        st0, st1, st2, st3 = {}, {}, {}, {}
        seq = vwscan.scanLocal(manager, settings, vws, range(len(vws)), workers=0, stats=st0)
        par = vwscan.scanLocal(manager, settings, vws, range(len(vws)), workers=3, concurrency=2, stats=st1,
                               reuseFactors=False)
        assert st0["solves"] == st1["solves"] > 2 * len(vws)
        assert np.array_equal(seq, par), "solve server + host workers must reproduce the sequential loop bit for bit"
        # factor re-use (the scan's default): from the third solve of a point on, refinement steps on the point's
        # last factors.  Same table for any number of workers; the plain table to solver accuracy.
        re3 = vwscan.scanLocal(manager, settings, vws, range(len(vws)), workers=3, stats=st2, reuseFactors=True)
        re1 = vwscan.scanLocal(manager, settings, vws, range(len(vws)), workers=1, stats=st3, reuseFactors=True)
        assert np.array_equal(re3, re1)
        r = st2["reuse"]
        assert r["enabled"] and r["slots"] == 3 and r["stale"] > 0, r
        assert r["fresh"] + r["stale"] == st2["solves"] == st0["solves"]
        assert np.allclose(re3[:, 0], seq[:, 0], rtol=1e-8, atol=0)
        assert np.max(np.abs(re3[:, 4:] - seq[:, 4:])) <= 1e-8 * np.max(np.abs(seq[:, 4:]))
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


def _raw_of(solver, bg):
    solver.setBackground(bg)
    b = solver.background
    dxidchi, _, _ = solver.grid.getCompactificationDerivatives()
    return (np.array(b.temperatureProfile, float), np.array(b.velocityProfile, float), solver._msq(b.fieldProfiles),
            np.array(dxidchi, float), float(b.velocityWall))


def _run_sequence(pool, raws, slot=None, seqs=None):
    out = {}
    for k, raw in enumerate(raws):
        assert pool.submit(k, raw=raw, slot=slot, seq=None if seqs is None else seqs[k])
        while pool.busy():
            for token, res in pool.poll():
                out[token] = res
    return [out[k] for k in range(len(raws))]


@pytest.mark.parametrize("P,M,N", [(2, 16, 9), (1, 30, 11)])
def test_slot_pool_factor_reuse_on_the_device(cuda_device, P, M, N):
    """slotPool(reuseFactors=True) on the CUDA engine (twin of tests/test_reuse_cpu.py): a sequence of nearby
    systems converges on the stale factors to the fresh LU + refinement answer (1e-12), positions 0 and 1 are the
    fresh path bit for bit, a far system falls back to a factorisation, and the run is reproducible."""
    import wallgo_b200 as wb
    from wallgo_b200 import synthetic

    grid = synthetic.makeGrid(M, N)
    parts = synthetic.makeParticles(P)
    s = wb.BoltzmannSolver(grid, device=cuda_device)
    s.updateParticleList(parts)
    s.setCollisionArray(synthetic.makeCollision(grid, parts, r=0.05))
    raws = [_raw_of(s, synthetic.makeBackground(grid, velocityMid=-0.5, wallL=w))
            for w in (0.05, 0.0505, 0.0508, 0.0509, 0.05095)]
    fresh = _run_sequence(s.slotPool(1), raws)
    pool = s.slotPool(2, reuseFactors=True)
    got = _run_sequence(pool, raws, slot=1, seqs=range(5))
    st = dict(pool.reuseStats)
    assert st["fresh"] == 2 and st["stale"] == 3 and st["fallbacks"] == 0, st
    for g, f in zip(got, fresh):
        assert np.max(np.abs(g.deltaF - f.deltaF)) <= 1e-12 * np.max(np.abs(f.deltaF))
        for name in ("Delta00", "Delta02", "Delta20", "Delta11"):
            a, b = getattr(g.Deltas, name).coefficients, getattr(f.Deltas, name).coefficients
            assert np.max(np.abs(a - b)) <= 1e-11 * np.max(np.abs(b))
        assert tuple(g.truncatedTail) == tuple(f.truncatedTail) and tuple(g.spectralPeaks) == tuple(f.spectralPeaks)
    assert np.array_equal(got[0].deltaF, fresh[0].deltaF) and np.array_equal(got[1].deltaF, fresh[1].deltaF)
    again = _run_sequence(s.slotPool(2, reuseFactors=True), raws, slot=0, seqs=range(5))
    assert all(np.array_equal(a.deltaF, g.deltaF) for a, g in zip(again, got)), "re-use must be reproducible"
    # a far member (temperature doubled: the collision term grows fourfold) falls back to the fresh path
    far = list(raws[:4])
    for k in (2, 3):
        far[k] = (2.0 * far[k][0],) + far[k][1:]
    fresh = _run_sequence(s.slotPool(1), far)
    pool = s.slotPool(1, reuseFactors=True)
    got = _run_sequence(pool, far, slot=0, seqs=range(4))
    st = dict(pool.reuseStats)
    assert st["fallbacks"] == 1 and st["fresh"] == 3 and st["stale"] == 1, st
    assert np.array_equal(got[2].deltaF, fresh[2].deltaF)
    assert np.max(np.abs(got[3].deltaF - fresh[3].deltaF)) <= 1e-12 * np.max(np.abs(fresh[3].deltaF))
