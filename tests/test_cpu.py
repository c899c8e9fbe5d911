"""CPU-only tests: the oracle against the reference golden vectors, the host-side logic, and the
C-ABI surface (library loads and exports every symbol of include/wallgo_b200.h; no compute)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden, oracle_problem_from_golden, relmax


# ----------------------------------------------------------------------------- oracle vs golden
@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_reference_golden(name):
    from oracle import boltzmann_oracle as bo

    g = load_golden(name)
    pb = oracle_problem_from_golden(g)
    for k, ref in (("Tchi", "ref_Tchi"), ("Dchi", "ref_Dchi"), ("Trz", "ref_Trz"), ("Trp", "ref_Trp"),
                   ("Drz", "ref_Drz")):
        assert relmax(getattr(pb.tab, k), g[ref]) < 1e-12, (name, k)
    op, src = bo.assemble(pb)
    assert relmax(src, g["ref_source"]) < 1e-11
    assert relmax(op[g["ref_op_probe_rows"]], g["ref_op_probe"]) < 1e-13
    assert relmax(op @ np.ones(op.shape[0]), g["ref_op_rowsum"]) < 1e-12
    dF = np.linalg.solve(op, src).reshape(g["ref_deltaF"].shape)
    assert relmax(dF, g["ref_deltaF"]) < 1e-10
    out = bo.get_deltas(pb, g["ref_deltaF"], str(g["option"]))
    assert tuple(out["truncatedTail"]) == tuple(bool(x) for x in g["ref_truncatedTail"])
    assert tuple(out["spectralPeaks"]) == tuple(int(x) for x in g["ref_spectralPeaks"])
    assert relmax(out["Deltas"], g["ref_Deltas"]) < 1e-12
    assert relmax(out["deltaF"], g["ref_deltaF_trunc"]) < 1e-12
    assert abs(out["truncationError"] - float(g["ref_truncationError"])) < 1e-9 * float(g["ref_truncationError"])
    T30, T33, pot, dV = bo.feedback(g["ref_Deltas"], g["msq"][:, 1:-1], g["dofs"], float(g["velocityMid"]),
                                    g["dmsq_interior"])
    assert relmax(T30, g["ref_T30"]) < 1e-13 and relmax(T33, g["ref_T33"]) < 1e-13
    assert relmax(pot, g["ref_potOut"]) < 1e-13 and relmax(dV, g["ref_dVout"]) < 1e-13
    if "ref_linearization" in g:
        c1, c2 = bo.check_linearization(pb, g["ref_deltaF"])
        assert abs(c1 - g["ref_linearization"][0]) < 1e-9 * abs(g["ref_linearization"][0])
        assert abs(c2 - g["ref_linearization"][1]) < 1e-8 * abs(g["ref_linearization"][1])


def test_oracle_delta00_known_answer():
    """Reference known-answer test tests/test_Boltzmann.py:14-85 through the oracle."""
    from oracle import boltzmann_oracle as bo

    for (M, N, a, b, c, d, e, f) in [(25, 19, 1, 2, 3, 4, 5, 6), (5, 5, 1, 1, 2, 3, 5, 8)]:
        tab = bo.build_tables(M, N, "Cardinal", "Cardinal", "Spectral")
        pz, pp, dpzdrz, dppdrp = bo.momentum_maps(tab.rz, tab.rp, 100.0)
        fieldv = np.ones(M + 1)
        fieldv[M // 2:] = 0
        fieldv += 0.1 * np.sin(7 * 2 * np.pi * np.arange(M + 1) + 6)
        msq = 0.5 * fieldv**2
        T = 100 * np.ones(M + 1)
        rz, rp = tab.rz[None, :, None], tab.rp[None, None, :]
        E = np.sqrt(msq[1:-1, None, None] + pz[None, :, None] ** 2 + pp[None, None, :] ** 2)
        eps = 2e-16
        integrand = (2 * E * (1 - rz**2) * (1 - rp**2)
                     * np.sqrt((1 - rz**2) * (1 - rp) ** 2 / (1 - rp**2 + eps)) / (np.log(2 / (1 - rp)) + eps))
        integrand = integrand * (a + b * rz + c * rz**2) * (d + e * rp + f * rp**2)
        pb = bo.Problem(tab=tab, xi_dxidchi=np.ones(M - 1), pz=pz, pp=pp, dpzdrz=dpzdrz, dppdrp=dppdrp, T=T,
                        v=np.zeros(M + 1), msq=msq[None, :], vw=0.0, statistics=np.array([-1.0]),
                        collision=np.zeros((1, N - 1, N - 1, 1, N - 1, N - 1)))
        D = bo.moments(integrand[None, ...], pb)
        exact = (4 * a + c) * (4 * d + f) * T[1:-1] ** 3 / 64
        np.testing.assert_allclose(D[0, 0], exact, rtol=1e-14, atol=0)


@pytest.mark.parametrize("M,N,slope", [(7, 9, -0.1), (9, 9, -0.1), (11, 9, 0.1)])
def test_oracle_truncation_decisions(M, N, slope):
    """Reference test tests/test_Boltzmann.py:128-174 (Chebyshev/Chebyshev, AUTO) through the oracle."""
    from oracle import boltzmann_oracle as bo

    tab = bo.build_tables(M, N, "Chebyshev", "Chebyshev", "Spectral")
    z, pz, pp = np.meshgrid(np.arange(M - 1), np.arange(N - 1), np.arange(N - 1), indexing="ij")
    dF = (np.exp(slope * z - 1e-2 * pz - 1e-3 * pp) / (1 + pz) / (1 + pp) ** 2)[None, ...]
    out, shape, _ = bo.check_spectral_convergence(dF, tab, "AUTO")
    nT = (M - 1) // 3
    if slope > 0:
        assert shape == (1, M - 1 - nT, N - 1, N - 1)
        assert np.allclose(out[:, -nT:], 0, atol=1e-2)
    else:
        assert shape == (1, M - 1, N - 1, N - 1)
        assert not np.allclose(out[:, -nT:], 0)


def test_lu_c_oracle_against_lapack():
    so = os.path.join(ROOT, "oracle", "_build", "liblu_oracle.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(so)
    import scipy.linalg as sl
    rng = np.random.default_rng(0)
    for n in (1, 7, 64, 301):
        A = rng.standard_normal((n, n))
        b = rng.standard_normal(n)
        LU = A.copy()
        piv = np.zeros(n, dtype=np.int32)
        assert lib.wgo_lu_factor(LU.ctypes.data_as(C.c_void_p), n, n, piv.ctypes.data_as(C.c_void_p)) == 0
        lu_ref, piv_ref = sl.lu_factor(A)
        assert np.array_equal(piv, piv_ref)
        assert relmax(LU, lu_ref) < 1e-11
        x = b.copy()
        lib.wgo_lu_solve(LU.ctypes.data_as(C.c_void_p), n, n, piv.ctypes.data_as(C.c_void_p),
                         x.ctypes.data_as(C.c_void_p))
        assert relmax(x, np.linalg.solve(A, b)) < 1e-10
    Z = np.zeros((3, 3))
    assert lib.wgo_lu_factor(Z.ctypes.data_as(C.c_void_p), 3, 3, np.zeros(3, np.int32).ctypes.data_as(C.c_void_p)) == 1


# ----------------------------------------------------------------------------- host logic
@pytest.mark.parametrize("name", golden_names())
def test_host_tables_match_reference(name):
    from wallgo_b200 import spectral

    g = load_golden(name)
    Tchi, Dchi, DchiFull, Trz, Drz, Trp = spectral.buildTables(int(g["M"]), int(g["N"]), str(g["basisM"]),
                                                               str(g["basisN"]), str(g["derivatives"]))
    for mine, ref in ((Tchi, "ref_Tchi"), (Dchi, "ref_Dchi"), (Trz, "ref_Trz"), (Trp, "ref_Trp"), (Drz, "ref_Drz")):
        assert relmax(mine, g[ref]) < 2e-13, (name, ref)
    if str(g["basisM"]) == "Cardinal":
        assert relmax(DchiFull[1:-1, 1:-1], Dchi) == 0.0


@pytest.mark.parametrize("name", ["cfg1_P1_M20_N11", "small_P2_M12_N7", "simplegrid_P3_M9_N5"])
def test_host_grid_matches_reference(name):
    import wallgo_b200 as wb

    g = load_golden(name)
    M, N = int(g["M"]), int(g["N"])
    grid = wb.Grid3Scales(M, N, 0.2, 0.2, 0.05, 100.0, 0.5, 0.1) if bool(g["grid3"]) else wb.Grid(M, N, 0.05, 100.0)
    chi, rz, rp = grid.getCompactCoordinates()
    assert relmax(chi, g["chi"]) < 1e-15 and relmax(rz, g["rz"]) < 1e-15 and relmax(rp, g["rp"]) < 1e-15
    dxidchi, dpzdrz, dppdrp = grid.getCompactificationDerivatives()
    assert relmax(dxidchi, g["dxidchi"]) < 1e-14
    assert relmax(dpzdrz, g["dpzdrz"]) < 1e-15 and relmax(dppdrp, g["dppdrp"]) < 1e-15
    _, pz, pp = grid.getCoordinates()
    assert relmax(pz, g["pz"]) < 1e-15 and relmax(pp, g["pp"]) < 1e-15
    assert grid.getCompactCoordinates(True, "z").size == M + 1
    assert grid.getCoordinates(True)[2][-1] == np.inf


def test_host_grid3scales_position_map():
    """xi(chi) of Grid3Scales against the oracle's restatement (grid3Scales.py:198-264)."""
    import wallgo_b200 as wb
    from oracle import boltzmann_oracle as bo

    grid = wb.Grid3Scales(30, 11, 0.3, 0.15, 0.04, 90.0, 0.4, 0.2, wallCenter=0.01)
    xi, d = bo.position_map_3scales(grid.chiValues, 0.3, 0.15, 0.04, 0.4, 0.2, 0.01)
    assert relmax(grid.xiValues, xi) < 1e-13
    assert relmax(grid.dxidchi, d) < 1e-14
    grid.changePositionFalloffScale(0.5, 0.6, 0.1, 0.0)
    xi, d = bo.position_map_3scales(grid.chiValues, 0.5, 0.6, 0.1, 0.4, 0.2, 0.0)
    assert relmax(grid.xiValues, xi) < 1e-13


def test_host_collision_change_basis():
    import wallgo_b200 as wb
    from oracle import boltzmann_oracle as bo

    M, N, P = 6, 7, 2
    tab = bo.build_tables(M, N)
    rng = np.random.default_rng(4)
    C0 = rng.standard_normal((P, N - 1, N - 1, P, N - 1, N - 1))
    grid = wb.Grid(M, N, 1.0, 1.0)
    ca = wb.CollisionArray(grid, "Cardinal", [None] * P, C0.copy())
    assert ca.changeBasis("Cardinal") is ca
    ca.changeBasis("Chebyshev")
    assert relmax(ca[...], bo.collision_to_chebyshev(C0, tab)) < 1e-12
    ca.changeBasis("Cardinal")
    assert relmax(ca[...], C0) < 1e-10
    assert ca.getBasisType() == "Cardinal" and ca.getBasisSize() == N - 1


@pytest.mark.parametrize("name", ["interp_P1_N9_to_N5", "interp_P2_N7_to_N5"])
def test_host_collision_interpolation_matches_reference(name):
    """CollisionArray.interpolateCollisionArray vs the live reference's output (golden)."""
    import wallgo_b200 as wb

    g = load_golden(name)
    P, M, Nf, Nt = int(g["P"]), int(g["M"]), int(g["Nf"]), int(g["Nt"])
    src = wb.CollisionArray(wb.Grid(M, Nf, 0.05, 100.0), "Chebyshev", [None] * P, g["source"])
    out = wb.CollisionArray.interpolateCollisionArray(src, wb.Grid(M, Nt, 0.05, 100.0))
    assert out.getBasisType() == str(g["basis"]) and out[...].shape == g["ref_interpolated"].shape
    assert relmax(out[...], g["ref_interpolated"]) < 1e-12
    with pytest.raises(AssertionError):
        wb.CollisionArray.interpolateCollisionArray(src, wb.Grid(M, Nf + 2, 0.05, 100.0))


@pytest.mark.parametrize("name", golden_names())
def test_host_truncation_decision(name):
    """The host half of checkSpectralConvergence, fed with spectra computed by the oracle."""
    import wallgo_b200 as wb
    from oracle import boltzmann_oracle as bo
    from conftest import FixtureGrid

    g = load_golden(name)
    pb = oracle_problem_from_golden(g)
    c = np.abs(bo.to_chebyshev(g["ref_deltaF"], pb.tab))
    spectra = np.concatenate([c.sum(axis=(0, 2, 3)), c.sum(axis=(0, 1, 3)), c.sum(axis=(0, 1, 2))])
    solver = wb.BoltzmannSolver(FixtureGrid(g), str(g["basisM"]), str(g["basisN"]), str(g["derivatives"]),
                                truncationOption=getattr(wb.ETruncationOption, str(g["option"])))
    keep, peaks = solver._decideTruncation(spectra)
    M1, N1 = int(g["M"]) - 1, int(g["N"]) - 1
    if N1 // 3 > 0 and M1 // 3 > 0:
        assert (keep[0] != M1, keep[1] != N1, keep[2] != N1) == tuple(bool(x) for x in g["ref_truncatedTail"])
        assert tuple(peaks) == tuple(int(x) for x in g["ref_spectralPeaks"])


def test_results_arithmetic_mirrors_reference():
    """scalar*res, res+res, res-res as used for under-relaxation (equationOfMotion.py:1352-1355)."""
    import wallgo_b200 as wb

    def mk(x):
        p = [wb.Polynomial(np.full((1, 3), x * (i + 1)), None, ("Array", "Cardinal"), ("Array", "z"), False)
             for i in range(4)]
        return wb.BoltzmannResults(np.full((1, 3, 2, 2), x), wb.BoltzmannDeltas(*p), 0.1 * x, (False,) * 3, (1, 2, 3))

    a, b = mk(1.0), mk(2.0)
    r = 0.25 * a + (1 - 0.25) * b
    assert np.allclose(r.deltaF, 1.75) and np.allclose(r.Deltas.Delta02.coefficients, 3.5)
    d = b - a
    assert np.allclose(d.Deltas.Delta11.coefficients, 4.0) and np.allclose(d.deltaF, 1.0)
    assert (-2 * a).truncationError == pytest.approx(0.2)


def test_solver_validates_arguments_like_reference():
    import wallgo_b200 as wb

    grid = wb.Grid(5, 5, 1.0, 1.0)
    with pytest.raises(AssertionError):
        wb.BoltzmannSolver(grid, basisM="Legendre")
    with pytest.raises(AssertionError):
        wb.BoltzmannSolver(grid, derivatives="Upwind")
    with pytest.raises(AssertionError):
        wb.BoltzmannSolver(grid, "Cardinal", "Chebyshev", "Finite Difference")
    with pytest.raises(ValueError):
        wb.Particle("x", 0, None, None, "Anyon", 1)
    s = wb.BoltzmannSolver(grid)
    with pytest.raises(AssertionError):
        s.updateParticleList([object()])


# ----------------------------------------------------------------------------- C ABI surface
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "wallgo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wg_[a-z_0-9]+)\s*\(", text)))


def test_capi_exports_every_declared_symbol():
    from wallgo_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libwallgo_b200.so not built (run __graft_entry__.build())")
    declared = _header_symbols()
    assert len(declared) >= 25
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/wallgo_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes binding and header disagree"
    assert _lib.load().wg_version() == 100


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    from wallgo_b200 import _lib
    from wallgo_b200.engine import BoltzmannEngine

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.WallGoB200Error):
        BoltzmannEngine(1, 6, 5)
    if os.path.exists(_lib.LIB_PATH):
        lib = _lib.load()
        h = C.c_void_p()
        assert lib.wg_create(1, 6, 5, 0, C.byref(h)) == -2  # WG_ERR_CUDA
        assert b"no CPU path" in lib.wg_last_error()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wallgo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower().replace(
                    "c oracle", ""), f"{f} mentions the oracle"


def test_tail_decision_equals_polyfit():
    """The truncation decision is the sign of a least-squares slope (polynomial.py:994-1033); the closed form
    used on the host path must decide exactly as np.polyfit does, including degenerate spectra."""
    import warnings
    from wallgo_b200.spectral import tailConverging

    def reference(coeffs, offset):
        m = len(coeffs)
        if m < 2:
            return False
        with warnings.catch_warnings(), np.errstate(divide="ignore", invalid="ignore"):
            warnings.simplefilter("ignore")
            fit = np.polyfit(np.arange(m) + offset, np.log(np.abs(coeffs)), 1, cov=(m >= 3))
        slope = fit[0][0] if m >= 3 else fit[0]
        return bool(slope < 0)

    rng = np.random.default_rng(0)
    cases = []
    for _ in range(1500):
        m = int(rng.integers(1, 14))
        decay = rng.uniform(-1.5, 1.5)
        cases.append(np.exp(decay * np.arange(m) + rng.normal(0, rng.uniform(0, 2), m)) * rng.choice([-1, 1], m))
    cases += [np.ones(6), np.zeros(5), np.array([1.0, 0.0, 1e-3, 1e-6]), np.array([2.0, 2.0]), np.array([1e-300, 1e300]),
              np.array([1.0, 1.0 + 1e-15, 1.0, 1.0 - 1e-15])]
    with warnings.catch_warnings(), np.errstate(divide="ignore", invalid="ignore"):
        warnings.simplefilter("ignore")
        for c in cases:
            for off in (0, 7):
                assert tailConverging(c, off) == reference(c, off), (c, off)


def test_auto_concurrency_policy():
    import wallgo_b200 as wb

    ac = wb.BoltzmannSolver.autoConcurrency
    assert ac(15600) == 4 and ac(11600) == 8 and ac(8700) == 8 and ac(3800) == 8 and ac(1900) == 32
    assert ac(31200, freeBytes=20 * 2**30) == 1          # one 7.8 GB operator + workspace in 60 % of 20 GB
    assert ac(15600, freeBytes=170 * 2**30) == 4
    assert ac(1900, freeBytes=2**30) >= 1


@pytest.mark.parametrize("M,N,option", [(3, 7, "AUTO"), (3, 7, "THIRD"), (8, 3, "AUTO"), (8, 3, "THIRD"), (3, 3, "THIRD"),
                                        (3, 3, "NONE"), (4, 5, "AUTO")])
def test_truncation_on_grids_with_fewer_than_three_coefficients(M, N, option):
    """With fewer than three coefficients along an axis the reference's cut is 0 and its slice `[-0:]` is the
    whole axis: a truncation zeroes everything along it while the reported shape does not change
    (boltzmann.py:400-446).  The host decision must tell the device to keep 0 coefficients there and report the
    reference's shape; checked against the oracle's restatement of the same lines."""
    import wallgo_b200 as wb
    from oracle import boltzmann_oracle as bo
    from conftest import FixtureGrid

    rng = np.random.default_rng(M * 100 + N)
    P = 2
    pb = bo.synthetic_problem(P, M, N, seed=1)
    dF = rng.standard_normal((P, M - 1, N - 1, N - 1)) * np.exp(0.3 * np.arange(M - 1))[None, :, None, None]
    ref_dF, ref_shape, ref_peaks = bo.check_spectral_convergence(dF, pb.tab, option)
    c = np.abs(bo.to_chebyshev(dF, pb.tab))
    spectra = np.concatenate([c.sum(axis=(0, 2, 3)), c.sum(axis=(0, 1, 3)), c.sum(axis=(0, 1, 2))])
    g = dict(P=P, M=M, N=N, chi=pb.tab.chi, rz=pb.tab.rz, rp=pb.tab.rp, pz=pb.pz, pp=pb.pp,
             dxidchi=pb.xi_dxidchi, dpzdrz=pb.dpzdrz, dppdrp=pb.dppdrp)
    solver = wb.BoltzmannSolver(FixtureGrid(g), truncationOption=getattr(wb.ETruncationOption, option))
    keep, peaks = solver._decideTruncation(spectra)
    assert tuple(keep) == tuple(ref_shape[1:])
    assert tuple(peaks) == tuple(ref_peaks)
    # what the device is told to keep reproduces the oracle's truncated coefficients
    cc = bo.to_chebyshev(dF, pb.tab).copy()
    cc[:, keep.dev[0]:, :, :] = 0
    cc[:, :, keep.dev[1]:, :] = 0
    cc[:, :, :, keep.dev[2]:] = 0
    ref_c = bo.to_chebyshev(ref_dF, pb.tab) if option != "NONE" else bo.to_chebyshev(dF, pb.tab)
    if option != "NONE":
        assert np.max(np.abs(cc - ref_c)) <= 1e-12 * max(1.0, np.max(np.abs(ref_c)))


def test_collision_directory_ingest_and_errors(tmp_path):
    """CollisionArray.newFromDirectory / BoltzmannSolver.loadCollisions (boltzmann.py:822-844,
    collisionArray.py:179-354): one file per ordered pair, CollisionLoadError for a missing file, for a grid
    larger than the files and for a size mismatch without interpolation."""
    import wallgo_b200 as wb

    g = load_golden("interp_P2_N7_to_N5")
    P, M, Nf = int(g["P"]), int(g["M"]), int(g["Nf"])
    parts = [wb.Particle(f"p{a}", a, None, None, "Fermion", 1) for a in range(P)]
    for a in range(P):
        for b in range(P):
            if (a + b) % 2:
                np.save(tmp_path / f"collisions_p{a}_p{b}.npy", g["source"][a, :, :, b])
            else:
                np.savez(tmp_path / f"collisions_p{a}_p{b}.npz", **{"data": g["source"][a, :, :, b], "Basis Size": Nf,
                                                                     "Basis Type": "Chebyshev"})
    grid = wb.Grid(M, Nf, 0.05, 100.0)
    ca = wb.CollisionArray.newFromDirectory(tmp_path, grid, "Chebyshev", parts)
    assert np.array_equal(ca[...], g["source"]) and ca.getBasisType() == "Chebyshev"
    solver = wb.BoltzmannSolver(grid)
    solver.updateParticleList(parts)
    solver.loadCollisions(tmp_path)
    assert np.array_equal(solver.collisionArray[...], g["source"])
    with pytest.raises(wb.CollisionLoadError):
        wb.CollisionArray.newFromDirectory(tmp_path, wb.Grid(M, Nf + 2, 0.05, 100.0), "Chebyshev", parts)
    with pytest.raises(wb.CollisionLoadError):
        wb.CollisionArray.newFromDirectory(tmp_path, wb.Grid(M, Nf - 2, 0.05, 100.0), "Chebyshev", parts, bInterpolate=False)
    (tmp_path / "collisions_p1_p0.npy").unlink()
    with pytest.raises(wb.CollisionLoadError):
        solver.loadCollisions(tmp_path)
    (tmp_path / "collisions_p1_p0.hdf5").write_text("version https://git-lfs.github.com/spec/v1\n")   # an LFS pointer
    with pytest.raises(wb.CollisionLoadError):
        solver.loadCollisions(tmp_path)


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Argument checks of the round-2 entry points run before any CUDA call: NULL handles and bad ranges return
    WG_ERR_ARG (-1) with a message; nothing computes on the CPU."""
    from wallgo_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    lib = _lib.load()
    null = C.c_void_p()
    out = C.c_void_p()
    assert lib.wg_batch_create(null, 4, C.byref(out)) == -1
    assert b"wg_batch_create" in lib.wg_last_error()
    assert lib.wg_batch_assemble(null, None, None, 1.0, 3, None, 8, 64, None, None) == -1
    assert lib.wg_batch_factor(null, None, 8, 64, None) == -1
    assert lib.wg_batch_solve(null, None, 8, 64, None, None) == -1
    assert lib.wg_batch_refine(null, None, None, 1.0, None, 8, 64, None, None, None) == -1
    assert lib.wg_batch_spectra(null, None, None, None) == -1
    assert lib.wg_batch_moments(null, None, None, None, 1, None, None, None, None) == -1
    assert lib.wg_batch_info(null, None, None) == -1
    assert lib.wg_batch_destroy(null) == 0
    assert lib.wg_refine(null, None, 0.5, 1.0, None, 8, None, None, None) == -1
    assert b"wg_refine" in lib.wg_last_error()
    assert lib.wg_assemble_rows(null, None, 0.5, 1.0, 3, 0, 1, None, 8, None, None) == -1
    assert lib.wg_lu_create_batched(0, 4, 0, C.byref(out)) == -1
    assert lib.wg_lu_create_batched(8, 0, 0, C.byref(out)) == -1
    assert lib.wg_lu_factor_batched(null, None, 8, 64, None, None, None) == -1
    assert lib.wg_lu_solve_batched(null, None, 8, 64, None, 8, None) == -1


def test_bench_workloads_and_profile_traffic():
    """bench.py: the workloads are BASELINE.json's configs, and the roofline `traffic` figures come out of the
    committed ncu summaries under profiles/ (not literals): the parser must find them."""
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench

    assert bench.WORKLOADS["cfg1"] == (1, 20, 11) and bench.WORKLOADS["cfg2a"] == (1, 40, 21)
    assert bench.WORKLOADS["cfg3"] == (3, 30, 11) and bench.WORKLOADS["cfg4"] == (4, 30, 11)
    assert bench.WORKLOADS["cfg5fit"] == (4, 32, 31) and bench.WORKLOADS["cfg5stream"] == (4, 64, 31)
    n = 15600
    gemm, src = bench.profile_traffic("dgemm_128x64_2cta")
    assert src and 3.8e9 < gemm < 5.0e9                     # 3.83e9 algorithmic at the first trailing shape
    asm, src = bench.profile_traffic("assemble")
    assert src and 0.98 * 8 * n * n < asm < 1.05 * 8 * n * n
    tri, src = bench.profile_traffic("trisolve_wave")
    assert src and 0.98 * 4 * n * n < tri < 1.05 * 4 * n * n  # one direction reads half the factors once
    assert bench.profile_traffic("no_such_kernel") == (None, None)


def test_scan_row_packing_round_trip():
    """vwscan rows: [pressure | widths | offsets | solves | Delta00,02,20,11]; unpackRow inverts the packing."""
    from wallgo_b200 import vwscan

    nF, P, M = 2, 3, 6
    width = vwscan.rowWidth(nF, P, M)
    assert width == 2 + 2 * nF + 4 * P * (M - 1)
    rng = np.random.default_rng(0)
    row = rng.standard_normal(width)
    d = vwscan.unpackRow(row, nF, P, M)
    assert d["pressure"] == row[0] and np.array_equal(d["widths"], row[1:3]) and np.array_equal(d["offsets"], row[3:5])
    assert d["solves"] == row[5] and d["Deltas"].shape == (4, P, M - 1)
    assert np.array_equal(d["Deltas"].ravel(), row[6:])
    with pytest.raises(AssertionError):
        vwscan.unpackRow(row[:-1], nF, P, M)
    assert 1 <= vwscan.defaultWorkers(1) <= 12 and 1 <= vwscan.defaultWorkers(8) <= 12


def test_extended_precision_yardstick():
    """oracle.solve_extended (the yardstick of the refinement tests and of tools/parity_fuzz.py): on integer systems
    whose exact solution is known (A, x integer, b = A x exact in FP64) it is exact to the last bits while LAPACK's
    own solve carries cond * eps."""
    from oracle import boltzmann_oracle as bo

    worstLap = 0.0
    for seed in range(4):
        rng = np.random.default_rng(seed)
        n = 150
        A = rng.integers(-3, 4, size=(n, n)).astype(float)
        A[:, 0] = A[:, 1] + rng.integers(-1, 2, size=n) * (rng.random(n) < 0.08)      # two nearly dependent columns
        xExact = rng.integers(-50, 51, size=n).astype(float)
        b = A @ xExact                                           # exact: small integers
        if np.linalg.matrix_rank(A) < n:
            continue
        xExt, xLap = bo.solve_extended(A, b)
        assert relmax(xExt, xExact) < 5e-16, (seed, relmax(xExt, xExact))
        worstLap = max(worstLap, relmax(xLap, xExact))
    assert worstLap > 1e-15          # the plain FP64 solve is measurably off on at least one of them
