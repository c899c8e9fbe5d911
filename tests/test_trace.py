"""Golden TRACE of the hot path as driven by the reference's own wall solver (Yukawa example,
off-equilibrium on; recorded by oracle/make_trace.py from the live reference).  Every getDeltas()
call of that solveWall run is replayed and compared to what the reference returned."""
import numpy as np
import pytest

from conftest import FixtureBackground, FixtureGrid, load_golden, relmax

TRACE = "trace_yukawa_M12_N7"


def _calls(tr):
    for i in range(int(tr["nCalls"])):
        yield i, dict(P=int(tr["P"]), M=int(tr["M"]), N=int(tr["N"]), chi=tr["chi"], rz=tr["rz"], rp=tr["rp"],
                      pz=tr["pz"], pp=tr["pp"], dpzdrz=tr["dpzdrz"], dppdrp=tr["dppdrp"],
                      dxidchi=tr["call_dxidchi"][i], T=tr["call_T"][i], v=tr["call_v"][i],
                      vw=float(tr["call_vw"][i]), msq=tr["call_msq"][i])


def test_trace_replay_oracle():
    from oracle import boltzmann_oracle as bo

    tr = load_golden(TRACE)
    tab = bo.build_tables(int(tr["M"]), int(tr["N"]))
    worst = 0.0
    for i, g in _calls(tr):
        pb = bo.Problem(tab=tab, xi_dxidchi=g["dxidchi"], pz=g["pz"], pp=g["pp"], dpzdrz=g["dpzdrz"],
                        dppdrp=g["dppdrp"], T=g["T"], v=g["v"], msq=g["msq"], vw=g["vw"],
                        statistics=tr["statistics"], collision=tr["collision"],
                        collisionMultiplier=float(tr["collisionMultiplier"]))
        out = bo.get_deltas(pb, None, "AUTO")
        assert tuple(out["truncatedTail"]) == tuple(bool(x) for x in tr["call_tail"][i])
        assert tuple(out["spectralPeaks"]) == tuple(int(x) for x in tr["call_peaks"][i])
        worst = max(worst, relmax(out["deltaF"], tr["call_deltaF"][i]), relmax(out["Deltas"], tr["call_Deltas"][i]))
    assert worst < 1e-10, worst


@pytest.mark.gpu
def test_trace_replay_gpu(cuda_device):
    """All 34 Boltzmann solves of the reference's solveWall run, through the CUDA path: deltaF and the
    four moments within 1e-10, truncation decisions and spectral peaks identical."""
    import wallgo_b200 as wb

    tr = load_golden(TRACE)
    P = int(tr["P"])
    solver = None
    worst = 0.0
    for i, g in _calls(tr):
        if solver is None:
            grid = FixtureGrid(g)
            parts = [wb.Particle(f"p{a}", a, (lambda f, a=a: f.msq[a]), None,
                                 "Fermion" if tr["statistics"][a] < 0 else "Boson", 4) for a in range(P)]
            solver = wb.BoltzmannSolver(grid, collisionMultiplier=float(tr["collisionMultiplier"]))
            solver.updateParticleList(parts)
            solver.setCollisionArray(wb.CollisionArray(grid, "Chebyshev", parts, tr["collision"]))
        solver.grid.dxidchi = g["dxidchi"]          # EOM re-maps z(chi) per wallPressure (same Grid object)
        solver.setBackground(FixtureBackground(g))
        res = solver.getDeltas()
        assert tuple(res.truncatedTail) == tuple(bool(x) for x in tr["call_tail"][i]), i
        assert tuple(res.spectralPeaks) == tuple(int(x) for x in tr["call_peaks"][i]), i
        D = np.stack([res.Deltas.Delta00.coefficients, res.Deltas.Delta02.coefficients,
                      res.Deltas.Delta20.coefficients, res.Deltas.Delta11.coefficients])
        worst = max(worst, relmax(res.deltaF, tr["call_deltaF"][i]), relmax(D, tr["call_Deltas"][i]))
        assert abs(res.truncationError - tr["call_truncErr"][i]) <= 1e-8 * tr["call_truncErr"][i]
    assert worst < 1e-10, worst
