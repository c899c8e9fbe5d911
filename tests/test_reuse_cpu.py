"""Host logic of FACTOR RE-USE in the slot pool (BoltzmannSolver.slotPool(reuseFactors=True)): a sequence of nearby
systems submitted to one slot is solved, from its third member on, by refinement steps preconditioned with the
factors of the sequence's last factorisation.  Runs on CPU with the oracle-backed TEST DOUBLE of the engine
(tests/fake_engine.py), whose refine_steps() is the same Richardson step in NumPy; the CUDA path has its own
`-m gpu` twin (tests/test_gpu_scan.py)."""
import numpy as np
import pytest


@pytest.fixture()
def solver(monkeypatch):
    from fake_engine import FakeEngine
    import wallgo_b200 as wb
    import wallgo_b200.engine as engine_mod
    from wallgo_b200 import synthetic

    monkeypatch.setattr(engine_mod, "BoltzmannEngine", FakeEngine)
    grid = synthetic.makeGrid(8, 5)
    parts = synthetic.makeParticles(2)
    s = wb.BoltzmannSolver(grid)
    s.updateParticleList(parts)
    s.setCollisionArray(synthetic.makeCollision(grid, parts))
    return s


def raw_of(solver, bg):
    """what a scan worker sends to the solve server (vwscan.RemoteBoltzmannSolver.getDeltas)"""
    solver.setBackground(bg)
    b = solver.background
    dxidchi, _, _ = solver.grid.getCompactificationDerivatives()
    return (np.array(b.temperatureProfile, float), np.array(b.velocityProfile, float), solver._msq(b.fieldProfiles),
            np.array(dxidchi, float), float(b.velocityWall))


def sequence(solver, widths, dv=0.02):
    from wallgo_b200 import synthetic
    return [raw_of(solver, synthetic.makeBackground(solver.grid, velocityMid=-0.5, wallL=w, dv=dv)) for w in widths]


def run(pool, raws, slot=None, seqs=None):
    out = {}
    for k, raw in enumerate(raws):
        assert pool.submit(k, raw=raw, slot=slot, seq=None if seqs is None else seqs[k])
        while pool.busy():
            for token, res in pool.poll():
                out[token] = res
    return [out[k] for k in range(len(raws))]


def close(a, b, tol):
    return np.max(np.abs(a - b)) <= tol * np.max(np.abs(b))


def test_sequence_converges_on_stale_factors_and_equals_fresh_solves(solver):
    raws = sequence(solver, [0.05, 0.0505, 0.0508, 0.0509, 0.05095])
    fresh = run(solver.slotPool(1), raws)
    pool = solver.slotPool(2, reuseFactors=True)
    got = run(pool, raws, slot=1, seqs=range(5))
    st = pool.reuseStats
    assert st["fresh"] == 2 and st["stale"] == 3 and st["fallbacks"] == 0
    assert 3 * pool.REUSE_CHUNK <= st["steps"] <= 3 * pool.REUSE_MAX_STEPS
    for g, f in zip(got, fresh):
        assert close(g.deltaF, f.deltaF, 1e-12)
        assert close(g.Deltas.Delta00.coefficients, f.Deltas.Delta00.coefficients, 1e-11)
        assert tuple(g.truncatedTail) == tuple(f.truncatedTail) and tuple(g.spectralPeaks) == tuple(f.spectralPeaks)
    # positions 0 and 1 always factor: bit-identical to the plain pool
    assert np.array_equal(got[0].deltaF, fresh[0].deltaF) and np.array_equal(got[1].deltaF, fresh[1].deltaF)


def test_far_system_falls_back_to_a_fresh_factorisation(solver):
    raws = sequence(solver, [0.05, 0.0505, 0.0508, 0.0509])
    for k in (2, 3):      # temperature doubled: the collision term grows fourfold, I - A_old^-1 A_new has norm ~3
        raws[k] = (2.0 * raws[k][0],) + raws[k][1:]
    fresh = run(solver.slotPool(1), raws)
    pool = solver.slotPool(1, reuseFactors=True)
    got = run(pool, raws, slot=0, seqs=range(4))
    st = pool.reuseStats
    assert st["fallbacks"] == 1 and st["fresh"] == 3 and st["stale"] == 1
    assert np.array_equal(got[2].deltaF, fresh[2].deltaF)          # the fall-back IS the fresh path
    assert close(got[3].deltaF, fresh[3].deltaF, 1e-12)            # and its factors serve the next member


def test_other_velocity_or_grid_or_no_sequence_number_always_factors(solver):
    from wallgo_b200 import synthetic
    raws = sequence(solver, [0.05, 0.0505, 0.0508])
    other = raw_of(solver, synthetic.makeBackground(solver.grid, velocityMid=-0.4, wallL=0.0508))
    pool = solver.slotPool(1, reuseFactors=True)
    run(pool, raws[:2] + [other], slot=0, seqs=[0, 1, 2])
    assert pool.reuseStats == dict(fresh=3, stale=0, fallbacks=0, steps=0)
    stretched = list(raws[2])
    stretched[3] = raws[2][3] * 1.01                               # another grid mapping (dxi/dchi)
    run(pool, raws[:2] + [tuple(stretched)], slot=0, seqs=[0, 1, 2])
    assert pool.reuseStats["stale"] == 0 and pool.reuseStats["fresh"] == 6
    run(pool, raws, slot=0, seqs=None)                             # no position: no re-use
    assert pool.reuseStats["stale"] == 0 and pool.reuseStats["fresh"] == 9
    plain = solver.slotPool(1)                                     # re-use off: sequence numbers are ignored
    run(plain, raws, slot=0, seqs=range(3))
    assert plain.reuseStats["stale"] == 0


def test_factorisation_limit_and_fixed_slots(solver):
    raws = sequence(solver, [0.05, 0.051])
    pool = solver.slotPool(2, reuseFactors=True)
    pool.maxFresh = 1
    assert pool.submit("a", raw=raws[0], slot=0, seq=0)
    assert not pool.submit("b", raw=raws[1], slot=0, seq=0)        # the slot is busy
    assert not pool.submit("b", raw=raws[1], slot=1, seq=0)        # one factorisation in flight already
    done = []
    while pool.busy():
        done += pool.poll()
    assert [t for t, _ in done] == ["a"]
    assert pool.submit("b", raw=raws[1], slot=1, seq=0)
    while pool.busy():
        done += pool.poll()
    assert [t for t, _ in done] == ["a", "b"]


def test_verdicts():
    from wallgo_b200.boltzmann import REUSE_CHUNK, REUSE_MAX_STEPS, reuseVerdict

    geo = lambda rho, k: [(rho ** (i + 1), 1.0) for i in range(k)]      # noqa: E731
    assert reuseVerdict(geo(0.5, 4)) == "more"            # the first chunk is never judged on its contraction
    assert reuseVerdict(geo(0.5, 8)) == "more"            # 0.5^4 = 0.06 per chunk
    assert reuseVerdict(geo(0.7, 8)) == "fail"            # 0.7^4 = 0.24 per chunk: too slow
    assert reuseVerdict(geo(1.1, 4)) == "more" and reuseVerdict(geo(1.1, 8)) == "fail"
    assert reuseVerdict(geo(1e-4, 4)) == "done"
    assert reuseVerdict(geo(0.5, REUSE_MAX_STEPS)) == "fail"            # out of steps
    assert reuseVerdict([(1.0, 1.0)] * 3 + [(float("nan"), 1.0)]) == "fail"
    assert reuseVerdict([(1.0, 1.0)] * 3 + [(0.0, 0.0)]) == "fail"      # 0/0
    assert REUSE_CHUNK * 2 <= REUSE_MAX_STEPS


def test_in_process_reuse_follows_getDeltas_sequences(solver):
    """solver.reuseFactors: consecutive getDeltas() with the same wall velocity and grid form a sequence."""
    from wallgo_b200 import synthetic
    bgs = [synthetic.makeBackground(solver.grid, velocityMid=-0.5, wallL=w) for w in (0.05, 0.0505, 0.0508, 0.0509)]
    plain = []
    for bg in bgs:
        solver.setBackground(bg)
        plain.append(solver.getDeltas())
    assert not hasattr(solver, "reuseStats")
    solver.reuseFactors = True
    got = []
    for bg in bgs:
        solver.setBackground(bg)
        got.append(solver.getDeltas())
    assert solver.reuseStats["fresh"] == 2 and solver.reuseStats["stale"] == 2 and solver.reuseStats["fallbacks"] == 0
    for g, f in zip(got, plain):
        assert close(g.deltaF, f.deltaF, 1e-12) and close(g.Deltas.Delta11.coefficients, f.Deltas.Delta11.coefficients, 1e-11)
    assert np.array_equal(got[0].deltaF, plain[0].deltaF)
    # another wall velocity starts a new sequence: two factorisations again
    for w in (0.05, 0.0505, 0.0508):
        solver.setBackground(synthetic.makeBackground(solver.grid, velocityMid=-0.4, wallL=w))
        solver.getDeltas()
    assert solver.reuseStats["fresh"] == 4 and solver.reuseStats["stale"] == 3
    # a copy (EOM.getBoltzmannFiniteDifference deep-copies the solver) owns no factors and starts afresh
    import copy
    c = copy.deepcopy(solver)
    c.setBackground(synthetic.makeBackground(solver.grid, velocityMid=-0.4, wallL=0.0509))
    c.getDeltas()
    assert c.reuseStats["fresh"] == 5 and c.reuseStats["stale"] == 3     # counters are copied, factors are not


def test_new_static_data_drops_the_stale_factors(solver):
    """A new collision tensor between two members of a sequence: the next member factors afresh."""
    from wallgo_b200 import synthetic
    raws = sequence(solver, [0.05, 0.0505, 0.0508, 0.0509])
    pool = solver.slotPool(1, reuseFactors=True)
    run(pool, raws[:3], slot=0, seqs=range(3))
    assert pool.reuseStats["stale"] == 1
    for s_ in pool.sibs:
        s_.setCollisionArray(synthetic.makeCollision(solver.grid, solver.offEqParticles, r=0.4))
    out = {}
    assert pool.submit(3, raw=raws[3], slot=0, seq=3)
    while pool.busy():
        out.update(dict(pool.poll()))
    assert pool.reuseStats["stale"] == 1 and pool.reuseStats["fresh"] == 3
    fresh = run(solver.slotPool(1), raws[3:])
    assert np.array_equal(out[3].deltaF, fresh[0].deltaF)


def test_other_uses_of_the_operator_buffer_drop_the_tag(solver):
    """buildLinearEquations() between two members of a getDeltas() sequence re-assembles into the operator buffer:
    the next member must not take it for factors."""
    from wallgo_b200 import synthetic
    solver.reuseFactors = True
    bgs = [synthetic.makeBackground(solver.grid, velocityMid=-0.5, wallL=w) for w in (0.05, 0.0505, 0.0508, 0.0509)]
    for bg in bgs[:3]:
        solver.setBackground(bg)
        solver.getDeltas()
    assert solver.reuseStats["stale"] == 1 and solver._engine.factorTag is not None
    solver.buildLinearEquations()
    assert solver._engine.factorTag is None
    solver.setBackground(bgs[3])
    res = solver.getDeltas()
    assert solver.reuseStats["stale"] == 1 and solver.reuseStats["fresh"] == 3
    solver.reuseFactors = False
    solver.setBackground(bgs[3])
    assert np.array_equal(res.deltaF, solver.getDeltas().deltaF)
