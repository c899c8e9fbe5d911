"""The wall-velocity scan server (wallgo_b200.vwscan) end to end on CPU: forked host workers run the unmodified
reference `EOM.wallPressure`, the server answers their Boltzmann solves from a slot pool -- here backed by the
oracle TEST DOUBLE of the engine (tests/fake_engine.py), so what is checked is the plumbing: message format,
per-point sequence numbers, one slot per worker, the factorisation limit, and that FACTOR RE-USE leaves the scan
table unchanged to solver accuracy and independent of the number of workers.  Needs the reference copy
(oracle/_ref, made by build())."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def rd():
    from oracle import ref_drivers
    if ref_drivers.reference_root() is None:
        pytest.skip("reference copy not present (oracle/fetch_ref.sh fills oracle/_ref in the build container)")
    return ref_drivers


def test_scan_server_with_and_without_factor_reuse(rd, monkeypatch):
    from fake_engine import FakeEngine
    import wallgo_b200.engine as engine_mod
    from wallgo_b200 import dropin, vwscan

    WallGo = rd.load_reference()
    import WallGo.manager as mgr

    monkeypatch.setattr(engine_mod, "BoltzmannEngine", FakeEngine)
    monkeypatch.setattr(vwscan, "_physicalGpus", lambda: 1)
    manager, settings, _ = rd.build_manager("yukawa", 12, 7)
    saved = mgr.BoltzmannSolver
    cls = dropin.install(vectoriseEOM=False)
    origNew = rd.patch_load_collisions(cls)
    vws = np.linspace(0.2, 0.5, 4)
    try:
        st = [dict() for _ in range(4)]
        seq = vwscan.scanLocal(manager, settings, vws, range(4), workers=0, stats=st[0])
        plain = vwscan.scanLocal(manager, settings, vws, range(4), workers=2, stats=st[1], reuseFactors=False)
        re2 = vwscan.scanLocal(manager, settings, vws, range(4), workers=2, stats=st[2], reuseFactors=True)
        re1 = vwscan.scanLocal(manager, settings, vws, range(4), workers=1, stats=st[3], reuseFactors=True)
    finally:
        mgr.BoltzmannSolver = saved
        if origNew is None:
            del cls.loadCollisions
        else:
            cls.loadCollisions = origNew
    assert np.array_equal(seq, plain), "server + workers without re-use must reproduce the in-process loop bit for bit"
    assert not st[1]["reuse"]["enabled"] and st[1]["reuse"]["stale"] == 0
    assert st[1]["reuse"]["fresh"] == st[1]["solves"] == st[0]["solves"]
    # with re-use: the same table whatever the number of workers, and the plain table to solver accuracy
    assert np.array_equal(re2, re1)
    r = st[2]["reuse"]
    assert r["enabled"] and r["slots"] == 2 and r["stale"] > 0
    assert r["fresh"] + r["stale"] == st[2]["solves"] == st[0]["solves"]        # same iteration counts per point
    assert r["fresh"] - r["fallbacks"] >= 2 * len(vws)                           # two factorisations open every point
    nF = 1
    assert np.allclose(re2[:, 0], seq[:, 0], rtol=1e-9, atol=0)                  # pressures
    assert np.allclose(re2[:, 1:1 + 2 * nF], seq[:, 1:1 + 2 * nF], rtol=1e-9, atol=1e-12)
    D, D0 = re2[:, 2 + 2 * nF:], seq[:, 2 + 2 * nF:]
    assert np.max(np.abs(D - D0)) <= 1e-9 * np.max(np.abs(D0))
