"""GPU parity tests: the CUDA path (through the C ABI) against the committed reference golden
vectors and against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): deltaF and the four Delta moments <= 1e-10 relative
(to the max-norm of the quantity); integer outputs (truncation flags, kept shapes, spectral
peaks) must be equal.
"""
import os

import numpy as np
import pytest

from conftest import golden_names, load_golden, oracle_problem_from_golden, relmax, solver_from_golden

pytestmark = pytest.mark.gpu

TOL = 1e-10          # north-star tolerance on deltaF and moments
TOL_ASSEMBLY = 2e-13  # operator/source: relative to the largest entry of the row / vector


@pytest.fixture(scope="module")
def oracle():
    from oracle import boltzmann_oracle as bo
    return bo


@pytest.mark.parametrize("name", [n for n in golden_names()])
def test_basis_tables_match_reference(cuda_device, name):
    g = load_golden(name)
    solver = solver_from_golden(g)
    eng = solver._getEngine()
    tabs = eng.get_tables()
    for k in ("Tchi", "Dchi", "Trz", "Trp", "Drz"):
        ref = g["ref_" + k]
        err = np.max(np.abs(tabs[k] - ref)) / np.max(np.abs(ref))
        assert err < 2e-13, (name, k, err)


@pytest.mark.parametrize("name", golden_names())
def test_operator_and_source_match_reference(cuda_device, name):
    g = load_golden(name)
    solver = solver_from_golden(g)
    op, src, liou, coll = solver.buildLinearEquations()
    n = op.shape[0]
    assert op.shape == (n, n) and src.shape == (n,)
    assert relmax(src, g["ref_source"]) < 1e-11, name
    scale = np.max(np.abs(g["ref_op_probe"]), axis=1, keepdims=True)
    rows = g["ref_op_probe_rows"]
    assert np.max(np.abs(op[rows] - g["ref_op_probe"]) / scale) < TOL_ASSEMBLY, name
    assert relmax(np.diag(op), g["ref_op_diag"]) < TOL_ASSEMBLY
    ones = np.ones(n)
    assert relmax(op @ ones, g["ref_op_rowsum"]) < 1e-12
    assert relmax(op.T @ ones, g["ref_op_colsum"]) < 1e-12
    if "ref_operator" in g:
        assert np.max(np.abs(op - g["ref_operator"])) / np.max(np.abs(g["ref_operator"])) < TOL_ASSEMBLY
    # the two parts add up to the operator (boltzmann.py:810)
    assert relmax((liou + coll).reshape(n, n), op) < 1e-15


@pytest.mark.parametrize("name", golden_names())
def test_deltaF_matches_reference(cuda_device, name):
    g = load_golden(name)
    solver = solver_from_golden(g)
    dF = solver.solveBoltzmannEquations()
    assert dF.shape == g["ref_deltaF"].shape
    assert relmax(dF, g["ref_deltaF"]) < TOL, name


@pytest.mark.parametrize("name", golden_names())
def test_getDeltas_matches_reference(cuda_device, name):
    g = load_golden(name)
    solver = solver_from_golden(g)
    res = solver.getDeltas()
    assert tuple(res.truncatedTail) == tuple(bool(x) for x in g["ref_truncatedTail"]), name
    assert tuple(res.spectralPeaks) == tuple(int(x) for x in g["ref_spectralPeaks"]), name
    assert relmax(res.deltaF, g["ref_deltaF_trunc"]) < TOL
    D = np.stack([res.Deltas.Delta00.coefficients, res.Deltas.Delta02.coefficients,
                  res.Deltas.Delta20.coefficients, res.Deltas.Delta11.coefficients])
    for i in range(4):
        assert relmax(D[i], g["ref_Deltas"][i]) < TOL, (name, i)
    assert abs(res.truncationError - float(g["ref_truncationError"])) <= 1e-8 * float(g["ref_truncationError"])
    # moments of a GIVEN deltaF (no solve): isolates the quadrature kernels
    res2 = solver.getDeltas(g["ref_deltaF"])
    D2 = np.stack([res2.Deltas.Delta00.coefficients, res2.Deltas.Delta02.coefficients,
                   res2.Deltas.Delta20.coefficients, res2.Deltas.Delta11.coefficients])
    assert relmax(D2, g["ref_Deltas"]) < 1e-12
    assert relmax(res2.deltaF, g["ref_deltaF_trunc"]) < 1e-12


@pytest.mark.parametrize("name", golden_names())
def test_feedback_matches_reference(cuda_device, name):
    """Device feedback kernel vs the reference's own EOM.deltaToTmunu and EOM source terms, computed from
    the reference's Deltas (isolates the kernel) and from the device solve (the production sequence)."""
    g = load_golden(name)
    solver = solver_from_golden(g)
    solver.getDeltas(g["ref_deltaF"])
    fb = solver.getFeedback()
    for key, ref in (("T30", "ref_T30"), ("T33", "ref_T33"), ("potentialOut", "ref_potOut"), ("dVout", "ref_dVout")):
        assert relmax(fb[key], g[ref]) < 1e-12, (name, key)
    solver.getDeltas()
    fb = solver.getFeedback()
    for key, ref in (("T30", "ref_T30"), ("T33", "ref_T33"), ("potentialOut", "ref_potOut"), ("dVout", "ref_dVout")):
        assert relmax(fb[key], g[ref]) < 1e-10, (name, key)


@pytest.mark.parametrize("name", ["tiny_P1_M6_N5", "small_P2_M12_N7"])
def test_checkLinearization_matches_reference(cuda_device, name):
    g = load_golden(name)
    solver = solver_from_golden(g)
    c1, c2 = solver.checkLinearization()
    assert abs(c1 - g["ref_linearization"][0]) < 1e-9 * abs(g["ref_linearization"][0])
    assert abs(c2 - g["ref_linearization"][1]) < 1e-8 * abs(g["ref_linearization"][1])


@pytest.mark.parametrize("P,M,N,kw", [
    (1, 16, 9, {}),
    (2, 10, 9, dict(r=0.05, seed=5)),
    (3, 8, 7, dict(collisionMultiplier=1e-2, seed=7)),
    (1, 24, 13, dict(seed=2)),
    (4, 8, 7, dict(seed=11)),              # four species (configs 4 and 5): off-diagonal blocks are collision only
    (1, 6, 31, dict(seed=3)),              # N = 31, the momentum grid of config 5 (q = 900)
    (2, 5, 21, dict(r=0.05, seed=4)),      # N = 21 with two species
])
def test_against_oracle_on_seeded_inputs(cuda_device, oracle, P, M, N, kw):
    """Sizes without a golden file: CUDA vs the oracle on the same synthetic inputs."""
    import wallgo_b200 as wb
    from conftest import FixtureBackground, FixtureGrid

    pb = oracle.synthetic_problem(P, M, N, **kw)
    g = dict(P=P, M=M, N=N, chi=pb.tab.chi, rz=pb.tab.rz, rp=pb.tab.rp, pz=pb.pz, pp=pb.pp,
             dxidchi=pb.xi_dxidchi, dpzdrz=pb.dpzdrz, dppdrp=pb.dppdrp, vw=pb.vw, v=pb.v, T=pb.T, msq=pb.msq)
    grid = FixtureGrid(g)
    parts = [wb.Particle(f"p{a}", a, (lambda f, a=a: f.msq[a]), None,
                         "Fermion" if pb.statistics[a] < 0 else "Boson", 12) for a in range(P)]
    solver = wb.BoltzmannSolver(grid, collisionMultiplier=pb.collisionMultiplier)
    solver.updateParticleList(parts)
    solver.setBackground(FixtureBackground(g))
    solver.setCollisionArray(wb.CollisionArray(grid, "Chebyshev", parts, pb.collision))
    ref = oracle.get_deltas(pb, None, "AUTO")
    res = solver.getDeltas()
    assert tuple(res.truncatedTail) == tuple(ref["truncatedTail"])
    assert tuple(res.spectralPeaks) == tuple(ref["spectralPeaks"])
    assert relmax(res.deltaF, ref["deltaF"]) < TOL
    D = np.stack([res.Deltas.Delta00.coefficients, res.Deltas.Delta02.coefficients,
                  res.Deltas.Delta20.coefficients, res.Deltas.Delta11.coefficients])
    for i in range(4):
        assert relmax(D[i], ref["Deltas"][i]) < TOL


@pytest.mark.parametrize("M,N,a,b,c,d,e,f", [(25, 19, 1, 2, 3, 4, 5, 6), (5, 5, 1, 1, 2, 3, 5, 8)])
def test_Delta00_known_answer(cuda_device, M, N, a, b, c, d, e, f):
    """Re-statement of the reference's analytic test (tests/test_Boltzmann.py:14-85), rtol 1e-14."""
    import wallgo_b200 as wb

    grid = wb.Grid(M, N, 1, 100)
    fieldv = np.ones(M + 1)
    fieldv[M // 2:] = 0
    fieldv += 0.1 * np.sin(7 * 2 * np.pi * np.arange(M + 1) + 6)
    v = -np.ones(M + 1) / np.sqrt(3) + 0.01 * np.sin(10 * 2 * np.pi * np.arange(M + 1))
    T = 100 * np.ones(M + 1)

    class F:
        def __init__(self, x):
            self.x = x

    particle = wb.Particle("top", 0, lambda fl: 0.5 * fl.x**2, None, "Fermion", 12)
    bg = wb.BoltzmannBackground(0.5 * (v[0] + v[-1]), v, F(fieldv), T)
    solver = wb.BoltzmannSolver(grid, "Cardinal", "Cardinal", "Spectral")
    solver.truncationOption = wb.ETruncationOption.NONE
    solver.updateParticleList([particle])
    solver.setBackground(bg)
    solver.setCollisionArray(wb.CollisionArray(grid, "Cardinal", [particle],
                                               np.zeros((1, N - 1, N - 1, 1, N - 1, N - 1))))
    rz = grid.rzValues[None, :, None]
    rp = grid.rpValues[None, None, :]
    pz = grid.pzValues[None, :, None]
    pp = grid.ppValues[None, None, :]
    msq = 0.5 * fieldv[1:-1, None, None] ** 2
    E = np.sqrt(msq + pz**2 + pp**2)
    eps = 2e-16
    integrand = (2 * E * (1 - rz**2) * (1 - rp**2)
                 * np.sqrt((1 - rz**2) * (1 - rp) ** 2 / (1 - rp**2 + eps)) / (np.log(2 / (1 - rp)) + eps))
    integrand = integrand * (a + b * rz + c * rz**2) * (d + e * rp + f * rp**2)
    res = solver.getDeltas(integrand[None, ...])
    exact = (4 * a + c) * (4 * d + f) * T**3 / 64
    np.testing.assert_allclose(res.Deltas.Delta00.coefficients[0], exact[1:-1], rtol=1e-14, atol=0)


@pytest.mark.parametrize("M,N", [(3, 3), (5, 5)])
def test_solution_residual(cuda_device, M, N):
    """Re-statement of tests/test_Boltzmann.py:88-125: ||A deltaF - S|| / ||S|| < 1e-14."""
    import wallgo_b200 as wb

    grid = wb.Grid(M, N, 1, 1)
    fieldv = np.ones(M + 1)
    fieldv[M // 2:] = 0
    fieldv += 0.1 * np.sin(7 * 2 * np.pi * np.arange(M + 1) + 6)
    v = -np.ones(M + 1) / np.sqrt(3) + 0.01 * np.sin(10 * 2 * np.pi * np.arange(M + 1))

    class F:
        def __init__(self, x):
            self.x = x

    particle = wb.Particle("top", 0, lambda fl: 0.5 * fl.x**2, None, "Fermion", 12)
    solver = wb.BoltzmannSolver(grid)
    solver.updateParticleList([particle])
    solver.setBackground(wb.BoltzmannBackground(0.5 * (v[0] + v[-1]), v, F(fieldv), 100 * np.ones(M + 1)))
    rng = np.random.default_rng(11)
    q = (N - 1) ** 2
    coll = (0.5 * np.identity(q) + 1e-2 * rng.standard_normal((q, q))).reshape(1, N - 1, N - 1, 1, N - 1, N - 1)
    solver.setCollisionArray(wb.CollisionArray(grid, "Chebyshev", [particle], coll))
    dF = solver.solveBoltzmannEquations()
    op, src, _, _ = solver.buildLinearEquations()
    ratio = np.linalg.norm(op @ dF.flatten() - src) / np.linalg.norm(src)
    assert ratio < 1e-14


def test_singular_operator_raises(cuda_device):
    """A singular operator must surface as numpy.linalg.LinAlgError like boltzmann.py:283 would."""
    g = load_golden("tiny_P1_M6_N5")
    solver = solver_from_golden(g)
    eng = solver._getEngine()
    solver._uploadBackground(eng)
    eng.assemble(float(g["vw"]), 1.0)
    eng.A.zero_()
    eng.factor()
    assert eng.info() == 1


def test_getDeltasBatch_equals_sequential(cuda_device):
    """Several independent systems in flight on one GPU give exactly the sequential results, in order."""
    import wallgo_b200 as wb
    from wallgo_b200 import synthetic

    P, M, N = 2, 12, 7
    grid = synthetic.makeGrid(M, N)
    parts = synthetic.makeParticles(P)
    solver = wb.BoltzmannSolver(grid)
    solver.updateParticleList(parts)
    solver.setCollisionArray(synthetic.makeCollision(grid, parts))
    bgs = [synthetic.makeBackground(grid, velocityMid=-0.3 - 0.05 * i) for i in range(7)]
    seq = []
    for bg in bgs:
        solver.setBackground(bg)
        seq.append(solver.getDeltas())
    for conc in (1, 2, 3):
        out = solver.getDeltasBatch(bgs, concurrency=conc)
        assert len(out) == len(bgs)
        for a, b in zip(seq, out):
            assert np.array_equal(a.deltaF, b.deltaF)
            assert np.array_equal(a.Deltas.Delta11.coefficients, b.Deltas.Delta11.coefficients)
            assert a.truncatedTail == b.truncatedTail and a.spectralPeaks == b.spectralPeaks
    assert solver.getDeltasBatch([], concurrency=2) == []


@pytest.mark.parametrize("P,M,N,option", [(2, 12, 7, "AUTO"), (1, 20, 11, "AUTO"), (3, 8, 9, "THIRD"), (1, 10, 7, "NONE")])
def test_getDeltasStrided_equals_sequential(cuda_device, P, M, N, option):
    """The lock-step strided-batch path (wg_batch_*: one launch sequence for a whole batch of small systems)
    returns, for every system, exactly what getDeltas() returns -- deltaF, all four moments, truncation
    decisions, peaks and error estimate -- for full, short (padded) and single batches."""
    import wallgo_b200 as wb
    from wallgo_b200 import synthetic
    from wallgo_b200.containers import ETruncationOption

    grid = synthetic.makeGrid(M, N)
    parts = synthetic.makeParticles(P)
    solver = wb.BoltzmannSolver(grid, truncationOption=ETruncationOption[option])
    solver.updateParticleList(parts)
    solver.setCollisionArray(synthetic.makeCollision(grid, parts))
    bgs = [synthetic.makeBackground(grid, velocityMid=-0.25 - 0.045 * i) for i in range(11)]
    seq = []
    for bg in bgs:
        solver.setBackground(bg)
        seq.append(solver.getDeltas())

    def same(a, b):
        assert np.array_equal(a.deltaF, b.deltaF)
        for name in ("Delta00", "Delta02", "Delta20", "Delta11"):
            assert np.array_equal(getattr(a.Deltas, name).coefficients, getattr(b.Deltas, name).coefficients)
        assert a.truncatedTail == b.truncatedTail and a.spectralPeaks == b.spectralPeaks
        assert a.truncationError == b.truncationError

    for batch in (4, 11, 1, None):
        out = solver.getDeltasStrided(bgs, batch=batch) if batch else solver.getDeltasBatch(bgs)
        assert len(out) == len(bgs)
        for a, b in zip(seq, out):
            same(a, b)
    # the single-system path still works on the same solver afterwards
    solver.setBackground(bgs[3])
    same(seq[3], solver.getDeltas())


def test_batched_c_abi_equals_single_calls(cuda_device):
    """wg_assemble_factor_solve_batched (B systems, B streams, one C call) gives exactly the deltaF of the
    single-system calls."""
    import torch
    import wallgo_b200 as wb
    from wallgo_b200 import synthetic
    from wallgo_b200.engine import BoltzmannEngine

    P, M, N = 2, 10, 7
    grid = synthetic.makeGrid(M, N)
    parts = synthetic.makeParticles(P)
    solver = wb.BoltzmannSolver(grid)
    solver.updateParticleList(parts)
    solver.setCollisionArray(synthetic.makeCollision(grid, parts))
    bgs = [synthetic.makeBackground(grid, velocityMid=-0.3 - 0.07 * i) for i in range(3)]
    sibs = solver._siblings(3)
    engines, vws, single = [], [], []
    for s_, bg in zip(sibs, bgs):
        s_.setBackground(bg)
        eng = s_._getEngine()
        s_._uploadBackground(eng)
        vw = float(s_.background.velocityWall)
        eng.assemble(vw, 1.0)
        eng.factor()
        eng.solve()
        single.append(eng.S.clone())
        engines.append(eng)
        vws.append(vw)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in engines]
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
    BoltzmannEngine.solve_batched(engines, vws, [1.0] * 3, streams)
    torch.cuda.synchronize()
    for eng, ref in zip(engines, single):
        assert eng.info() == 0
        assert torch.equal(eng.S, ref)


def test_overflow_guard_matches_oracle(cuda_device, oracle):
    """A cold plasma (T ~ 0.5 on a momentum grid scaled for T = 100) drives E/T beyond MAX_EXPONENT ~ 709.78 on
    part of the grid: the equilibrium-distribution guard of boltzmann.py:864-886 must act on the same points
    (source exactly -0 there) and the solve must agree with the oracle."""
    import wallgo_b200 as wb
    from conftest import FixtureBackground, FixtureGrid

    P, M, N = 2, 8, 7
    pb = oracle.synthetic_problem(P, M, N, seed=9)
    pb.T = pb.T / 200.0
    rq = oracle.row_quantities(pb)
    x = rq["Epl"] / rq["T"]
    assert np.any(x > 709.782712893384) and np.any(x < 700.0)      # the guard is exercised, and not everywhere
    g = dict(P=P, M=M, N=N, chi=pb.tab.chi, rz=pb.tab.rz, rp=pb.tab.rp, pz=pb.pz, pp=pb.pp,
             dxidchi=pb.xi_dxidchi, dpzdrz=pb.dpzdrz, dppdrp=pb.dppdrp, vw=pb.vw, v=pb.v, T=pb.T, msq=pb.msq)
    grid = FixtureGrid(g)
    parts = [wb.Particle(f"p{a}", a, (lambda f, a=a: f.msq[a]), None,
                         "Fermion" if pb.statistics[a] < 0 else "Boson", 12) for a in range(P)]
    solver = wb.BoltzmannSolver(grid)
    solver.updateParticleList(parts)
    solver.setBackground(FixtureBackground(g))
    solver.setCollisionArray(wb.CollisionArray(grid, "Chebyshev", parts, pb.collision))
    op_ref, src_ref = oracle.assemble(pb)
    op, src, _, _ = solver.buildLinearEquations()
    assert np.all(np.isfinite(src)) and np.all(np.isfinite(op))
    assert relmax(src, src_ref) < 1e-12
    assert np.array_equal(src == 0.0, src_ref == 0.0)
    assert relmax(op, op_ref) < 1e-12
    ref = oracle.get_deltas(pb, None, "AUTO")
    res = solver.getDeltas()
    assert tuple(res.truncatedTail) == tuple(ref["truncatedTail"])
    assert relmax(res.deltaF, ref["deltaF"]) < TOL


@pytest.mark.parametrize("tool,args", [("parity_fuzz.py", ("20", "11")), ("lu_fuzz.py", ("44", "5"))])
def test_fuzz_subsets(cuda_device, tool, args):
    """Seeded subsets of the two fuzzers: random sizes / bases / derivative and truncation options against the
    oracle (1e-10, equal truncation flags and peaks), and structured matrices (sparse, low rank, duplicated rows,
    300 orders of magnitude, integer ties, zero blocks, non-finite entries) against LAPACK."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", tool), *args], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " 0 bad" in r.stdout.splitlines()[-1]


def test_two_solvers_interleaved(cuda_device):
    """Two solvers of different sizes used alternately in one process (each with its own handle, LU plan and
    cached graphs) give what each gives alone."""
    import wallgo_b200 as wb
    from wallgo_b200 import synthetic

    def make(P, M, N):
        grid = synthetic.makeGrid(M, N)
        parts = synthetic.makeParticles(P)
        s = wb.BoltzmannSolver(grid)
        s.updateParticleList(parts)
        s.setCollisionArray(synthetic.makeCollision(grid, parts))
        return s, grid

    a, ga = make(1, 8, 5)
    b, gb = make(2, 10, 7)
    vms = (-0.3, -0.45, -0.6)
    alone_a, alone_b = [], []
    for vm in vms:
        a.setBackground(synthetic.makeBackground(ga, velocityMid=vm))
        alone_a.append(a.getDeltas())
    for vm in vms:
        b.setBackground(synthetic.makeBackground(gb, velocityMid=vm))
        alone_b.append(b.getDeltas())
    for i, vm in enumerate(vms):
        a.setBackground(synthetic.makeBackground(ga, velocityMid=vm))
        b.setBackground(synthetic.makeBackground(gb, velocityMid=vm))
        rb = b.getDeltas()
        ra = a.getDeltas()
        assert np.array_equal(ra.deltaF, alone_a[i].deltaF) and np.array_equal(rb.deltaF, alone_b[i].deltaF)
        assert np.array_equal(ra.Deltas.Delta20.coefficients, alone_a[i].Deltas.Delta20.coefficients)
        assert np.array_equal(rb.Deltas.Delta11.coefficients, alone_b[i].Deltas.Delta11.coefficients)


@pytest.mark.parametrize("name", ["interp_P1_N9_to_N5", "interp_P2_N7_to_N5"])
def test_collision_ingest_interpolates_on_the_device(cuda_device, name, tmp_path):
    """File ingest with N_f > N: the down-interpolation kernel (wg_interpolate_collision) against the live
    reference's CollisionArray.interpolateCollisionArray (golden), through loadCollisions."""
    import wallgo_b200 as wb

    g = load_golden(name)
    P, M, Nf, Nt = int(g["P"]), int(g["M"]), int(g["Nf"]), int(g["Nt"])
    parts = [wb.Particle(f"p{a}", a, None, None, "Fermion", 1) for a in range(P)]
    for a in range(P):
        for b in range(P):
            np.savez(tmp_path / f"collisions_p{a}_p{b}.npz", **{"data": g["source"][a, :, :, b], "Basis Size": Nf,
                                                                 "Basis Type": "Chebyshev"})
    solver = wb.BoltzmannSolver(wb.Grid(M, Nt, 0.05, 100.0), basisN=str(g["basis"]))
    solver.updateParticleList(parts)
    solver.loadCollisions(tmp_path)
    assert solver.collisionArray[...].shape == g["ref_interpolated"].shape
    assert relmax(solver.collisionArray[...], g["ref_interpolated"]) < 1e-12
    src = wb.CollisionArray(wb.Grid(M, Nf, 0.05, 100.0), "Chebyshev", parts, g["source"])
    host = wb.CollisionArray.interpolateCollisionArray(src, wb.Grid(M, Nt, 0.05, 100.0))
    assert relmax(solver.collisionArray[...], host[...]) < 1e-13


def test_wg_moments_requires_its_spectra_call(cuda_device):
    """wg_moments consumes the Chebyshev coefficients of the preceding wg_spectra on the same deltaF: calling it
    twice, or on another vector, is a state error instead of a silent result from stale data."""
    from wallgo_b200 import _lib

    g = load_golden("tiny_P1_M6_N5")
    solver = solver_from_golden(g)
    solver.getDeltas()
    eng = solver._getEngine()
    with pytest.raises(_lib.WallGoB200Error):
        eng.moments_async(eng.S, (5, 4, 4), True)          # second wg_moments without a new wg_spectra
    eng.spectra_async(eng.S)
    with pytest.raises(_lib.WallGoB200Error):
        eng.moments_async(eng.dF_in, (5, 4, 4), True)      # other vector than the one wg_spectra saw
    eng.spectra_async(eng.S)
    eng.moments_async(eng.S, (5, 4, 4), True)


def test_solvers_on_two_devices_in_one_process(cuda_device):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = load_golden("small_P2_M12_N7")
    a, b = solver_from_golden(g, device=0), solver_from_golden(g, device=1)
    ra, rb = a.getDeltas(), b.getDeltas()
    assert torch.cuda.current_device() == 0
    assert np.array_equal(ra.deltaF, rb.deltaF)
    assert relmax(rb.deltaF, g["ref_deltaF_trunc"]) < TOL


def test_per_solver_tuning_does_not_leak(cuda_device):
    """wg_solver_set_tuning touches one solver; results are identical for every schedule."""
    g = load_golden("cfg1_P1_M20_N11")
    a, b = solver_from_golden(g), solver_from_golden(g)
    ea, eb = a._getEngine(), b._getEngine()
    from wallgo_b200 import _lib
    _lib.check(ea.lib.wg_solver_set_tuning(ea.h, 64, 0, 0, 0, 0), "wg_solver_set_tuning")
    ra, rb = a.getDeltas(), b.getDeltas()
    assert relmax(ra.deltaF, rb.deltaF) < 1e-12 and relmax(ra.deltaF, g["ref_deltaF_trunc"]) < TOL


def _fuzz_solver(P, M, N, **kw):
    """Solver + oracle problem of one tools/parity_fuzz.py case."""
    import wallgo_b200 as wb
    from conftest import FixtureBackground, FixtureGrid
    from oracle import boltzmann_oracle as bo

    option = kw.pop("option", "AUTO")
    pb = bo.synthetic_problem(P, M, N, **kw)
    g = dict(P=P, M=M, N=N, chi=pb.tab.chi, rz=pb.tab.rz, rp=pb.tab.rp, pz=pb.pz, pp=pb.pp,
             dxidchi=pb.xi_dxidchi, dpzdrz=pb.dpzdrz, dppdrp=pb.dppdrp, vw=pb.vw, v=pb.v, T=pb.T, msq=pb.msq)
    grid = FixtureGrid(g)
    parts = [wb.Particle(f"p{a}", a, (lambda f, a=a: f.msq[a]), None, "Fermion" if pb.statistics[a] < 0 else "Boson", 12)
             for a in range(P)]
    solver = wb.BoltzmannSolver(grid, kw.get("basisM", "Cardinal"), kw.get("basisN", "Chebyshev"),
                                kw.get("derivatives", "Spectral"), pb.collisionMultiplier,
                                getattr(wb.ETruncationOption, option))
    solver.updateParticleList(parts)
    solver.setBackground(FixtureBackground(g))
    solver.setCollisionArray(wb.CollisionArray(grid, kw.get("basisN", "Chebyshev"), parts, pb.collision))
    return solver, pb


@pytest.mark.parametrize("case", [
    # the round-1 fuzz outlier (cond ~ 1.2e7: LAPACK itself is 3.6e-11 off the true solution)
    dict(P=3, M=10, N=13, seed=220, r=0.05, collisionMultiplier=0.01, vw_mid=-0.7303480839249901, basisM="Cardinal",
         basisN="Cardinal", derivatives="Finite Difference", grid3=False, option="THIRD"),
    dict(P=2, M=12, N=9, seed=7, r=0.05, collisionMultiplier=0.01, vw_mid=-0.6, grid3=True),
    dict(P=1, M=14, N=11, seed=3, r=0.5, collisionMultiplier=1.0, vw_mid=-0.3, grid3=True),
])
def test_refinement_reaches_the_extended_precision_solution(cuda_device, case):
    """wg_refine (residual with regenerated operator entries in double-double + one solve with the existing
    factors): deltaF agrees with the extended-precision solution of the assembled FP64 system to 1e-12 -- far
    inside the 1e-10 gate also where cond(A) * eps is not -- and the reported correction is the forward error
    the unrefined solve had."""
    from oracle import boltzmann_oracle as bo

    case = dict(case)
    P, M, N = case.pop("P"), case.pop("M"), case.pop("N")
    solver, pb = _fuzz_solver(P, M, N, **case)
    # yardstick 1: the extended-precision solution of the operator the DEVICE assembled (what wg_refine converges to)
    Adev, bdev, _, _ = solver.buildLinearEquations()
    xDev, _ = bo.solve_extended(Adev, bdev)
    # yardstick 2: the same for the oracle's operator (differs from the device's in the last bits of every entry)
    A, b = bo.assemble(pb)
    xTrue, xLapack = bo.solve_extended(A, b)
    solver.refinementSteps = 0
    x0 = solver.solveBoltzmannEquations().reshape(-1)
    solver.refinementSteps = 1
    x1 = solver.solveBoltzmannEquations().reshape(-1)
    e0, e1, eL = relmax(x0, xDev), relmax(x1, xDev), relmax(xLapack, xTrue)
    print(f"n={A.shape[0]} cond~{np.linalg.cond(A):.1e}: unrefined {e0:.2e}, refined {e1:.2e} (own operator); LAPACK on the "
          f"oracle operator {eL:.2e}; refined vs oracle-operator truth {relmax(x1, xTrue):.2e}")
    assert e1 < 1e-14
    assert relmax(x1, xTrue) < 1e-12                      # the gate is 1e-10
    assert relmax(x1, xLapack) < 1e-10 or eL > 5e-11      # the reference path's own answer, where it is accurate
    solver.getDeltas()
    dmax, xmax = solver.lastRefinement
    assert abs(dmax / xmax - e0) <= 0.5 * e0 + 1e-15      # the correction IS the forward error of the first solve
    assert solver.conditionEstimate() == dmax / xmax / 2.0 ** -53


def test_refinement_in_every_path_is_the_same(cuda_device):
    """getDeltas, the slot pool and the strided batch all apply the same refinement step: bit-identical results,
    and switching it off changes them (so the step really runs)."""
    import wallgo_b200 as wb
    from wallgo_b200 import synthetic

    P, M, N = 2, 10, 7
    grid = synthetic.makeGrid(M, N)
    parts = synthetic.makeParticles(P)
    solver = wb.BoltzmannSolver(grid)
    solver.updateParticleList(parts)
    solver.setCollisionArray(synthetic.makeCollision(grid, parts))
    bgs = [synthetic.makeBackground(grid, velocityMid=-0.3 - 0.05 * i) for i in range(5)]
    seq = []
    for bg in bgs:
        solver.setBackground(bg)
        seq.append(solver.getDeltas().deltaF)
    for out in (solver.getDeltasBatch(bgs, concurrency=2), solver.getDeltasStrided(bgs, batch=3)):
        for a, r in zip(seq, out):
            assert np.array_equal(a, r.deltaF)
    solver.refinementSteps = 0
    solver.setBackground(bgs[0])
    plain = solver.getDeltas().deltaF
    assert not np.array_equal(plain, seq[0])
    assert relmax(plain, seq[0]) < 1e-10


@pytest.mark.parametrize("P,M,N,G", [(2, 9, 7, 3), (3, 6, 5, 4), (1, 12, 9, 11)])
def test_assemble_rows_equals_full_assembly(cuda_device, P, M, N, G):
    """Row-panel assembly (wg_assemble_rows, the out-of-core form for operators larger than HBM) writes exactly
    the bytes of the resident assembly, panel by panel, for full, ragged and single-group panels."""
    import ctypes as C
    import torch
    import wallgo_b200 as wb
    from wallgo_b200 import synthetic
    from wallgo_b200._lib import check

    grid = synthetic.makeGrid(M, N)
    parts = synthetic.makeParticles(P)
    solver = wb.BoltzmannSolver(grid)
    solver.updateParticleList(parts)
    solver.setCollisionArray(synthetic.makeCollision(grid, parts))
    solver.setBackground(synthetic.makeBackground(grid, velocityMid=-0.4))
    eng = solver._getEngine()
    solver._uploadBackground(eng)
    vw = float(solver.background.velocityWall)
    full = eng.assemble(vw, 1.0).clone()
    S0 = eng.S.clone()
    n, q, groups = eng.n, eng.q, P * (M - 1)
    for g in (G, 1, groups):
        got = torch.zeros_like(full)
        panel = torch.empty((g * q, n), dtype=torch.float64, device=full.device)
        for g0 in range(0, groups, g):
            ng = min(g, groups - g0)
            panel.fill_(float("nan"))
            check(eng.lib.wg_assemble_rows(eng.h, eng.prof.data_ptr(), vw, 1.0, 3, g0, ng, panel.data_ptr(), n,
                                           eng.S.data_ptr(), eng._stream()), "wg_assemble_rows")
            got[g0 * q:(g0 + ng) * q] = panel[:ng * q]
        assert torch.equal(got, full)
        assert torch.equal(eng.S, S0)
    h = C.c_void_p(eng.h.value)
    assert eng.lib.wg_assemble_rows(h, eng.prof.data_ptr(), vw, 1.0, 3, groups, 1, panel.data_ptr(), n,
                                    eng.S.data_ptr(), eng._stream()) == -1


def test_strided_batches_survive_an_engine_change(cuda_device):
    """Batch contexts borrow the solver's engine handle: when the engine is rebuilt (another species count) they
    are released first, and the next batch builds new ones; empty input returns an empty list."""
    import gc
    import wallgo_b200 as wb
    from wallgo_b200 import synthetic

    M, N = 10, 7
    grid = synthetic.makeGrid(M, N)
    solver = wb.BoltzmannSolver(grid)
    assert solver.getDeltasStrided([]) == []
    for P in (1, 3, 2):
        parts = synthetic.makeParticles(P)
        solver.updateParticleList(parts)
        solver.setCollisionArray(synthetic.makeCollision(grid, parts))
        bgs = [synthetic.makeBackground(grid, velocityMid=-0.3 - 0.05 * i) for i in range(5)]
        out = solver.getDeltasStrided(bgs, batch=4)
        solver.setBackground(bgs[2])
        one = solver.getDeltas()
        assert np.array_equal(out[2].deltaF, one.deltaF)
        assert out[2].deltaF.shape == (P, M - 1, N - 1, N - 1)
    del solver
    gc.collect()
