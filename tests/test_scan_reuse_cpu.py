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


def _setup_scan(vws):
    """(manager, settings, cleanup) for the Yukawa driver with the test-double engine behind the drop-in"""
    import sys
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import ref_drivers as rd
    from fake_engine import FakeEngine
    import wallgo_b200.engine as engine_mod
    from wallgo_b200 import dropin, vwscan

    rd.load_reference()
    engine_mod.BoltzmannEngine = FakeEngine
    vwscan._physicalGpus = lambda: 1
    manager, settings, _ = rd.build_manager("yukawa", 12, 7)
    cls = dropin.install(vectoriseEOM=False)
    rd.patch_load_collisions(cls)
    return manager, settings


def _rank_main(rank, world, port, vws, q):
    import logging
    import os
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    logging.disable(logging.CRITICAL)
    import torch.distributed as dist
    manager, settings = _setup_scan(vws)
    from wallgo_b200 import vwscan
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st = {}
        table = vwscan.scanWallPressure(manager, settings, vws, workers=1, stats=st, reuseFactors=True)
        q.put((rank, table, st["reuse"], st["done"]))
    finally:
        dist.destroy_process_group()


def test_scanWallPressure_world_size_2_gloo(rd):
    """The scan as the metric defines it on two ranks (gloo, CPU, test-double engine): each rank runs its own
    solve server with factor re-use on its share of the points, one all_gather returns the table; it equals the
    single-process table bit for bit."""
    import socket
    import torch.multiprocessing as mp

    vws = np.linspace(0.2, 0.5, 4)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, vws, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # the same scan in this process (fork is fine here: no process group in the parent)
    import subprocess
    import sys
    import os
    import pickle
    code = ("import sys, pickle, numpy as np, logging; logging.disable(logging.CRITICAL); sys.path.insert(0, %r); "
            "import test_scan_reuse_cpu as t; vws = np.linspace(0.2, 0.5, 4); m, s = t._setup_scan(vws); "
            "from wallgo_b200 import vwscan; "
            "tab = vwscan.scanLocal(m, s, vws, range(4), workers=1, reuseFactors=True); "
            "sys.stdout.buffer.write(b'TABLE' + pickle.dumps(tab))") % os.path.dirname(os.path.abspath(__file__))
    raw = subprocess.run([sys.executable, "-c", code], capture_output=True, timeout=600, check=True).stdout
    single = pickle.loads(raw[raw.index(b"TABLE") + 5:])
    for rank, table, reuse, done in out:
        assert done == 2 and reuse["enabled"] and reuse["slots"] == 1
        assert np.array_equal(table, single), rank
    assert sum(o[2]["stale"] for o in out) > 0
