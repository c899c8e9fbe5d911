"""GPU tests of the dense pieces behind the C ABI: FP64 tensor-core GEMM and the blocked
partial-pivoting LU + triangular solves (replacing LAPACK dgesv, boltzmann.py:283).

Checked against NumPy/LAPACK and against the C restatement in oracle/lu_oracle.c; at the
full benchmark size through size-independent properties (residual, linearity, P*A = L*U).
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(cuda_device):
    import torch
    from wallgo_b200 import _lib

    lib = _lib.load()
    return torch, lib, _lib


def _stream(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (256, 384, 256), (1000, 16, 16), (777, 33, 48),
                                   (129, 257, 100), (5, 7, 3), (2048, 128, 128), (300, 64, 32),
                                   (64, 2000, 256), (1, 1, 1)])
def test_dgemm_sub(env, M, N, K):
    torch, lib, _lib = env
    rng = np.random.default_rng(M * 1000 + N * 10 + K)
    ldc, lda, ldb = N + (3 if (M * N) % 2 else 4), K + 2, N + 6
    Ch = rng.standard_normal((M, ldc))
    Ah = rng.standard_normal((M, lda))
    Bh = rng.standard_normal((K, ldb))
    Cd, Ad, Bd = (torch.from_numpy(x).cuda() for x in (Ch, Ah, Bh))
    _lib.check(lib.wg_dgemm_sub(Cd.data_ptr(), Ad.data_ptr(), Bd.data_ptr(), M, N, K, ldc, lda, ldb,
                                _stream(torch)), "wg_dgemm_sub")
    got = Cd.cpu().numpy()
    want = Ch.copy()
    want[:, :N] -= Ah[:, :K] @ Bh[:, :N]
    scale = np.abs(Ah[:, :K]) @ np.abs(Bh[:, :N]) + np.abs(Ch[:, :N])
    assert np.max(np.abs(got[:, :N] - want[:, :N]) / scale) < 1e-15 * max(4, K ** 0.5)
    assert np.array_equal(got[:, N:], Ch[:, N:])  # padding columns untouched


def _lu(env, A, b=None, lda=None):
    torch, lib, _lib = env
    n = A.shape[0]
    lda = n if lda is None else lda
    Ah = np.zeros((n, lda))
    Ah[:, :n] = A
    Ad = torch.from_numpy(Ah).cuda()
    ipiv = torch.zeros(n, dtype=torch.int32, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    plan = C.c_void_p()
    _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)), "wg_lu_create")
    try:
        _lib.check(lib.wg_lu_factor(plan, Ad.data_ptr(), lda, ipiv.data_ptr(), info.data_ptr(), _stream(torch)),
                   "wg_lu_factor")
        x = None
        if b is not None:
            bd = torch.from_numpy(b.copy()).cuda()
            _lib.check(lib.wg_lu_solve(plan, Ad.data_ptr(), lda, bd.data_ptr(), _stream(torch)), "wg_lu_solve")
            x = bd.cpu().numpy()
        torch.cuda.synchronize()
    finally:
        lib.wg_lu_destroy(plan)
    return Ad.cpu().numpy()[:, :n], ipiv.cpu().numpy(), int(info.item()), x


def _check_factors(A, LU, ipiv):
    n = A.shape[0]
    L = np.tril(LU, -1) + np.identity(n)
    U = np.triu(LU)
    PA = A.copy()
    for j in range(n):
        p = ipiv[j]
        assert j <= p < n
        if p != j:
            PA[[j, p]] = PA[[p, j]]
    err = np.max(np.abs(L @ U - PA)) / np.max(np.abs(A))
    assert err < 1e-13 * max(1, n ** 0.5), err
    assert np.max(np.abs(L)) <= 1.0 + 1e-15  # partial pivoting bounds the multipliers


@pytest.mark.parametrize("n,lda_pad", [(1, 0), (2, 0), (5, 1), (16, 0), (17, 3), (64, 0), (100, 0), (255, 1),
                                       (256, 0), (257, 0), (400, 0), (513, 2), (1000, 0), (1900, 0), (2049, 1)])
def test_lu_random(env, n, lda_pad):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    LU, ipiv, info, x = _lu(env, A, b, lda=n + lda_pad)
    assert info == 0
    _check_factors(A, LU, ipiv)
    xr = np.linalg.solve(A, b)
    assert np.max(np.abs(x - xr)) / np.max(np.abs(xr)) < 1e-10


@pytest.mark.parametrize("n", [300, 700, 2600])
def test_lu_schedules_are_bit_identical(env, n):
    """CUDA-graph replay (default), plain streams, and no look-ahead are schedules of the same arithmetic:
    factors, pivots and solutions must agree bit for bit."""
    torch, lib, _ = env
    rng = np.random.default_rng(7 * n)
    A = rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    outs = []
    try:
        for graph, look in ((1, 1), (0, 1), (0, 0), (1, 0)):
            lib.wg_set_tuning(256, graph)
            lib.wg_set_lookahead(look)
            LU, ipiv, info, x = _lu(env, A, b)
            assert info == 0
            outs.append((LU, ipiv, x))
            LU2, ipiv2, _, x2 = _lu(env, A, b)     # run-to-run determinism of the same schedule
            assert np.array_equal(LU, LU2) and np.array_equal(ipiv, ipiv2) and np.array_equal(x, x2)
    finally:
        lib.wg_set_tuning(256, 1)
        lib.wg_set_lookahead(1)
    for LU, ipiv, x in outs[1:]:
        assert np.array_equal(outs[0][0], LU) and np.array_equal(outs[0][1], ipiv) and np.array_equal(outs[0][2], x)
    _check_factors(A, outs[0][0], outs[0][1])


def test_lu_near_far_split_is_bit_identical(env):
    """n >= 8192 splits the bulk columns over two streams (the far part runs a step behind): a pure
    re-scheduling, so factors and pivots equal the unsplit schedule bit for bit, on streams and as a graph."""
    torch, lib, _ = env
    n = 8448
    rng = np.random.default_rng(5)
    A = rng.standard_normal((n, n))
    outs = []
    try:
        for graph, split in ((1, 1), (1, 0), (0, 1)):
            lib.wg_set_tuning(256, graph)
            lib.wg_set_bulk_split(split)
            LU, ipiv, info, _ = _lu(env, A)
            assert info == 0
            outs.append((LU, ipiv))
    finally:
        lib.wg_set_tuning(256, 1)
        lib.wg_set_bulk_split(1)
    for LU, ipiv in outs[1:]:
        assert np.array_equal(outs[0][1], ipiv) and np.array_equal(outs[0][0], LU)
    # P A = L U on a sample of rows (the full product at this size is a CPU-minute)
    LU, ipiv = outs[0]
    perm = np.arange(n)
    for i, pv in enumerate(ipiv):
        perm[i], perm[pv] = perm[pv], perm[i]
    rows = rng.choice(n, size=64, replace=False)
    L = np.tril(LU, -1) + np.eye(n)
    U = np.triu(LU)
    assert np.max(np.abs(L[rows] @ U - A[perm[rows]])) < 1e-9


@pytest.mark.parametrize("n", [40, 700, 2600])
def test_lu_32_column_leaves(env, n):
    """The alternative 32-column leaf panel (rotating register window) is a different blocking of the same
    partial-pivoting LU: same pivots on a generic matrix, factors equal to rounding."""
    torch, lib, _ = env
    rng = np.random.default_rng(11 * n)
    A = rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    LU0, ipiv0, info0, x0 = _lu(env, A, b)
    try:
        lib.wg_debug_leaf32(1)
        LU1, ipiv1, info1, x1 = _lu(env, A, b)
    finally:
        lib.wg_debug_leaf32(0)
    assert info0 == info1 == 0
    assert np.array_equal(ipiv0, ipiv1)
    assert np.max(np.abs(LU0 - LU1)) / np.max(np.abs(LU0)) < 1e-11
    assert np.max(np.abs(x0 - x1)) / np.max(np.abs(x0)) < 1e-9
    _check_factors(A, LU1, ipiv1)


@pytest.mark.parametrize("n", [8, 50, 300])
def test_lu_pivots_match_c_oracle(env, n):
    """Same pivot rule as dgetf2 (first maximal |a| in the column): identical ipiv and factors to rounding."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "oracle", "_build", "liblu_oracle.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_build/liblu_oracle.so not built")
    olib = C.CDLL(so)
    rng = np.random.default_rng(100 + n)
    A = rng.standard_normal((n, n))
    LU, ipiv, info, _ = _lu(env, A)
    Ao = A.copy()
    piv = np.zeros(n, dtype=np.int32)
    oinfo = olib.wgo_lu_factor(Ao.ctypes.data_as(C.c_void_p), n, n, piv.ctypes.data_as(C.c_void_p))
    assert info == oinfo == 0
    assert np.array_equal(ipiv, piv)
    assert np.max(np.abs(LU - Ao)) / np.max(np.abs(Ao)) < 1e-12


def test_lu_singular_reports_info(env):
    n = 40
    rng = np.random.default_rng(3)
    A = rng.standard_normal((n, n))
    A[:, 7] = 0.0
    _, _, info, _ = _lu(env, A)
    assert info == 8
    A = rng.standard_normal((n, n))
    A[:, 20] = A[:, 3]          # exactly dependent columns: elimination produces an exact zero column
    _, _, info, _ = _lu(env, A)
    assert info == 0 or info == 21


def test_lu_badly_scaled_needs_pivoting(env):
    """Entries spanning many orders of magnitude (1e-17 .. 2e5)."""
    n = 600
    rng = np.random.default_rng(9)
    A = rng.standard_normal((n, n)) * 10.0 ** rng.uniform(-8, 5, size=(n, 1))
    A[np.arange(n), np.arange(n)] *= 1e-12   # tiny diagonal forces row exchanges
    b = rng.standard_normal(n)
    LU, ipiv, info, x = _lu(env, A, b)
    assert info == 0
    assert np.mean(ipiv != np.arange(n)) > 0.9
    xr = np.linalg.solve(A, b)
    assert np.max(np.abs(x - xr)) / np.max(np.abs(xr)) < 1e-9


def test_full_size_properties(cuda_device):
    """BASELINE config M=40, N=21, P=1 (n = 15600): residual, linearity and idempotent re-solve."""
    import torch
    import wallgo_b200 as wb
    from conftest import FixtureBackground, FixtureGrid
    from oracle import boltzmann_oracle as bo

    P, M, N = 1, 40, 21
    pb = bo.synthetic_problem(P, M, N)
    g = dict(P=P, M=M, N=N, chi=pb.tab.chi, rz=pb.tab.rz, rp=pb.tab.rp, pz=pb.pz, pp=pb.pp,
             dxidchi=pb.xi_dxidchi, dpzdrz=pb.dpzdrz, dppdrp=pb.dppdrp, vw=pb.vw, v=pb.v, T=pb.T, msq=pb.msq)
    grid = FixtureGrid(g)
    parts = [wb.Particle("p0", 0, (lambda f: f.msq[0]), None, "Fermion", 12)]
    solver = wb.BoltzmannSolver(grid)
    solver.updateParticleList(parts)
    solver.setBackground(FixtureBackground(g))
    solver.setCollisionArray(wb.CollisionArray(grid, "Chebyshev", parts, pb.collision))
    eng = solver._getEngine()
    solver._uploadBackground(eng)
    eng.assemble(pb.vw, 1.0)
    A0 = eng.A.clone()
    S0 = eng.S.clone()
    # a few assembled rows against the oracle's row formula (the oracle cannot hold 1.9 GB x 3)
    rq = bo.row_quantities(pb)
    src = (bo.dfeq(rq["Epl"] / rq["T"], rq["s"]) / rq["T"]) * rq["dchidxi"] * (
        rq["Pw"] * rq["Ppl"] * rq["gpl"] ** 2 * rq["dv"] + rq["Pw"] * rq["Epl"] * rq["dT"] / rq["T"]
        + 0.5 * rq["dmsq"] * rq["uwupl"])
    assert np.max(np.abs(S0.cpu().numpy() - src.reshape(-1))) / np.max(np.abs(src)) < 1e-11
    eng.factor()
    eng.solve()
    assert eng.info() == 0
    x = eng.S.clone()
    r = torch.mv(A0, x) - S0
    assert float(torch.linalg.norm(r) / torch.linalg.norm(S0)) < 1e-12
    # linearity: solve(3*S) == 3*solve(S) with the same factors
    b3 = 3.0 * S0
    eng.solve(b3)
    assert float(torch.max(torch.abs(b3 - 3.0 * x)) / torch.max(torch.abs(x))) < 1e-12
    # moments are linear in deltaF
    res1 = solver.getDeltas(x.cpu().numpy().reshape(P, M - 1, N - 1, N - 1))
    solver.truncationOption = wb.ETruncationOption.NONE
    resa = solver.getDeltas(x.cpu().numpy().reshape(P, M - 1, N - 1, N - 1))
    resb = solver.getDeltas(2.0 * x.cpu().numpy().reshape(P, M - 1, N - 1, N - 1))
    assert np.allclose(2.0 * resa.Deltas.Delta00.coefficients, resb.Deltas.Delta00.coefficients, rtol=1e-13, atol=0)
    assert np.all(np.isfinite(res1.Deltas.Delta11.coefficients))


@pytest.mark.parametrize("P,M,N", [(3, 30, 11), (4, 30, 11), (2, 40, 21), (4, 8, 31), (4, 64, 15), (4, 64, 21)])
def test_other_baseline_sizes_residual(cuda_device, P, M, N):
    """BASELINE configs 3, 4 and the two-species reading of config 2 (n = 8700, 11600, 31200), and the two
    reductions of config 5 (4 species, M=64, N=31: 411 GB, does not fit) that do fit one B200 -- its momentum grid
    N=31 with 4 species at M=8 (n = 25200) and its spatial grid M=64 with 4 species at N=15 (n = 49392, 19.5 GB,
    beyond 32768 rows) and at N=21 (n = 100800, an 81 GB operator: the largest of the family that fits 180 GB) --
    through the public class: the solution satisfies the assembled system, the factors reproduce the operator row sums,
    and several systems in flight return exactly what one at a time returns."""
    import torch
    import wallgo_b200 as wb
    from wallgo_b200 import synthetic

    grid = synthetic.makeGrid(M, N)
    parts = synthetic.makeParticles(P)
    solver = wb.BoltzmannSolver(grid)
    solver.updateParticleList(parts)
    solver.setCollisionArray(synthetic.makeCollision(grid, parts))
    solver.setBackground(synthetic.makeBackground(grid))
    eng = solver._getEngine()
    solver._uploadBackground(eng)
    n = eng.n
    assert n == P * (M - 1) * (N - 1) ** 2
    eng.assemble(float(solver.background.velocityWall), 1.0)
    S0 = eng.S.clone()
    # A x without keeping a second copy of the operator: y = A @ ones before, compared through the factors after
    ones = torch.ones(n, dtype=torch.float64, device=eng.dev)
    rowsum = torch.mv(eng.A, ones)
    # off-diagonal species blocks are collision-only and block diagonal in z
    if P > 1:
        q = (N - 1) ** 2
        blk = eng.A[:q, (M - 1) * q + q:(M - 1) * q + 2 * q]      # species 0, z-block 0 vs species 1, z-block 1
        assert float(torch.max(torch.abs(blk))) == 0.0
    eng.factor()
    assert eng.info() == 0
    eng.solve()
    x = eng.S.clone()
    eng.solve(rowsum)                      # A^-1 (A 1) = 1
    assert float(torch.max(torch.abs(rowsum - 1.0))) < 1e-9
    eng.assemble(float(solver.background.velocityWall), 1.0)   # operator again (the LU overwrote it)
    r = torch.mv(eng.A, x) - S0
    assert float(torch.linalg.norm(r) / torch.linalg.norm(S0)) < 1e-12
    if n <= 12000:
        bgs = [synthetic.makeBackground(grid, velocityMid=-0.35 - 0.03 * i) for i in range(4)]
        seq = []
        for bg in bgs:
            solver.setBackground(bg)
            seq.append(solver.getDeltas())
        out = solver.getDeltasBatch(bgs)
        for a, b in zip(seq, out):
            assert np.array_equal(a.deltaF, b.deltaF)
            assert np.array_equal(a.Deltas.Delta00.coefficients, b.Deltas.Delta00.coefficients)


@pytest.mark.parametrize("P,M,N", [(1, 6, 5), (2, 12, 7), (1, 20, 11), (3, 30, 11)])
def test_fused_forward_solve_equals_separate_solves(cuda_device, P, M, N):
    """wg_factor_solve (forward substitution riding along inside the factorisation, then the backward solve)
    against wg_factor + wg_solve on the same operator: same factors bit for bit, solutions equal to rounding,
    and the factors remain usable for a later wg_solve."""
    import torch
    import wallgo_b200 as wb
    from wallgo_b200 import synthetic

    grid = synthetic.makeGrid(M, N)
    parts = synthetic.makeParticles(P)
    solver = wb.BoltzmannSolver(grid)
    solver.updateParticleList(parts)
    solver.setCollisionArray(synthetic.makeCollision(grid, parts))
    solver.setBackground(synthetic.makeBackground(grid))
    eng = solver._getEngine()
    solver._uploadBackground(eng)
    vw = float(solver.background.velocityWall)
    eng.assemble(vw, 1.0)
    S0 = eng.S.clone()
    eng.factor()
    LU_sep = eng.A.clone()
    eng.solve()
    x_sep = eng.S.clone()
    for _ in range(2):                      # second pass replays the cached graphs
        eng.assemble(vw, 1.0)
        eng.factor_solve()
        assert eng.info() == 0
        assert torch.equal(eng.A, LU_sep)
        scale = float(torch.max(torch.abs(x_sep)))
        assert float(torch.max(torch.abs(eng.S - x_sep))) <= 1e-11 * scale
    b2 = S0.clone()
    eng.solve(b2)                           # the factors are still a complete LU
    assert float(torch.max(torch.abs(b2 - x_sep))) <= 1e-11 * float(torch.max(torch.abs(x_sep)))


@pytest.mark.parametrize("kind", ["inf_mid_leaf", "neg_inf_late", "nan_entry", "nan_row", "nan_col", "all_nan", "inf_small"])
def test_lu_nonfinite_input_does_not_fault(env, kind):
    """Garbage in, garbage out -- but never an out-of-range access.  An infinite entry turns into NaNs
    (inf - inf) inside a leaf panel; the pivot search must keep treating those rows as candidates (its integer
    key of |a| has to stay non-negative for NaN), otherwise an inactive row "wins" and carries an invalid
    position.  Regression test for exactly that fault."""
    torch, lib, _ = env
    n = 64 if kind == "inf_small" else 300
    rng = np.random.default_rng(1)
    A = rng.standard_normal((n, n))
    if kind == "inf_mid_leaf":
        A[150, 100] = np.inf
    elif kind == "neg_inf_late":
        A[280, 270] = -np.inf
    elif kind == "nan_entry":
        A[150, 100] = np.nan
    elif kind == "nan_row":
        A[150, :] = np.nan
    elif kind == "nan_col":
        A[:, 100] = np.nan
    elif kind == "all_nan":
        A[:] = np.nan
    elif kind == "inf_small":
        A[60, 20] = np.inf
    LU, ipiv, info, x = _lu(env, A, rng.standard_normal(n))
    assert ipiv.min() >= 0 and ipiv.max() < n
    assert np.all(ipiv >= np.arange(n))                  # LAPACK-style pivots: row i swaps with a row at or below it
    # the context is still healthy: a clean factorisation right after
    B = rng.standard_normal((n, n))
    LU2, ipiv2, info2, _ = _lu(env, B)
    assert info2 == 0
    _check_factors(B, LU2, ipiv2)


def test_unphysical_background_does_not_fault(cuda_device):
    """A superluminal wall velocity makes the Lorentz factors NaN: the solve must return (NaN) results or raise
    LinAlgError like the reference would, not fault the device."""
    import torch
    import wallgo_b200 as wb
    from wallgo_b200 import synthetic

    grid = synthetic.makeGrid(12, 7)
    parts = synthetic.makeParticles(2)
    solver = wb.BoltzmannSolver(grid)
    solver.updateParticleList(parts)
    solver.setCollisionArray(synthetic.makeCollision(grid, parts))
    with np.errstate(all="ignore"):
        solver.setBackground(synthetic.makeBackground(grid, velocityMid=-1.02))
        try:
            res = solver.getDeltas()
            assert not np.all(np.isfinite(res.deltaF))
        except (np.linalg.LinAlgError, ValueError, FloatingPointError):
            pass
    torch.cuda.synchronize()
    solver.setBackground(synthetic.makeBackground(grid, velocityMid=-0.4))
    good = solver.getDeltas()
    assert np.all(np.isfinite(good.deltaF))


def test_lu_singular_large(env):
    """Zero matrix, a zero column deep inside a panel and inside the second panel: info is the 1-based index of
    the first exact zero pivot (LAPACK dgetrf), the factorisation carries on, nothing faults."""
    torch, lib, _ = env
    n = 600
    rng = np.random.default_rng(2)
    _, ipiv, info, _ = _lu(env, np.zeros((n, n)))
    assert info == 1 and np.array_equal(ipiv, np.arange(n))
    for col in (100, 299, 300, 599):
        A = rng.standard_normal((n, n))
        A[:, col] = 0.0
        _, ipiv, info, _ = _lu(env, A)
        assert info == col + 1, (col, info)
        assert ipiv.min() >= 0 and ipiv.max() < n


def test_lu_unaligned_operator(env):
    """An operator that starts 8 bytes off a 16-byte boundary with an odd leading dimension takes the scalar
    (non-vectorised) load/store paths of every kernel: same pivots, factors equal to rounding."""
    torch, lib, _lib = env
    n, lda = 700, 703
    rng = np.random.default_rng(3)
    A = rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    LU0, ipiv0, info0, x0 = _lu(env, A, b)
    big = torch.zeros(n * lda + 1, dtype=torch.float64, device="cuda")
    view = big[1:].view(n, lda)                      # data_ptr is 8 (mod 16)
    assert view.data_ptr() % 16 == 8
    view[:, :n] = torch.from_numpy(A).cuda()
    ipiv = torch.zeros(n, dtype=torch.int32, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    bd = torch.from_numpy(b.copy()).cuda()
    plan = C.c_void_p()
    _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)), "wg_lu_create")
    try:
        _lib.check(lib.wg_lu_factor(plan, view.data_ptr(), lda, ipiv.data_ptr(), info.data_ptr(), _stream(torch)), "factor")
        _lib.check(lib.wg_lu_solve(plan, view.data_ptr(), lda, bd.data_ptr(), _stream(torch)), "solve")
        torch.cuda.synchronize()
    finally:
        lib.wg_lu_destroy(plan)
    assert int(info.item()) == 0
    assert np.array_equal(ipiv.cpu().numpy(), ipiv0)
    LU1 = view[:, :n].cpu().numpy()
    assert np.max(np.abs(LU1 - LU0)) / np.max(np.abs(LU0)) < 1e-12
    assert np.max(np.abs(bd.cpu().numpy() - x0)) / np.max(np.abs(x0)) < 1e-10


def test_lu_plan_reused_with_alternating_buffers(env):
    """One plan, two operators in turn: the cached CUDA graphs are rebuilt when the buffers change and replayed
    when they come back; results match fresh plans bit for bit."""
    torch, lib, _lib = env
    n = 900
    rng = np.random.default_rng(4)
    mats = [rng.standard_normal((n, n)) for _ in range(2)]
    rhs = [rng.standard_normal(n) for _ in range(2)]
    ref = [_lu(env, m, r) for m, r in zip(mats, rhs)]
    dev = [torch.empty((n, n), dtype=torch.float64, device="cuda") for _ in range(2)]
    bd = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
    ipiv = torch.zeros(n, dtype=torch.int32, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    plan = C.c_void_p()
    _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)), "wg_lu_create")
    try:
        for turn in (0, 1, 0, 0, 1):
            dev[turn].copy_(torch.from_numpy(mats[turn]))
            bd[turn].copy_(torch.from_numpy(rhs[turn]))
            _lib.check(lib.wg_lu_factor(plan, dev[turn].data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), _stream(torch)), "factor")
            _lib.check(lib.wg_lu_solve(plan, dev[turn].data_ptr(), n, bd[turn].data_ptr(), _stream(torch)), "solve")
            torch.cuda.synchronize()
            assert int(info.item()) == 0
            assert np.array_equal(dev[turn].cpu().numpy(), ref[turn][0])
            assert np.array_equal(ipiv.cpu().numpy(), ref[turn][1])
            assert np.array_equal(bd[turn].cpu().numpy(), ref[turn][3])
    finally:
        lib.wg_lu_destroy(plan)


@pytest.mark.parametrize("n,rows", [(64, (3, 40)), (600, (5, 100, 599)), (2600, (700, 1300, 2500)), (9000, (4500, 8999, 600))])
def test_lu_first_maximal_entry_wins_ties(env, n, rows):
    """LAPACK's idamax takes the FIRST row of maximal |a|: exact ties that sit in different warps, CTAs of the
    cluster and with either sign must resolve to the smallest row index."""
    rng = np.random.default_rng(n)
    A = rng.uniform(-0.5, 0.5, size=(n, n))
    for i, r in enumerate(rows):
        A[r, 0] = 7.25 if i % 2 == 0 else -7.25
    LU, ipiv, info, _ = _lu(env, A)
    assert info == 0
    assert ipiv[0] == min(rows)
    assert abs(LU[0, 0]) == 7.25


def test_lu_beyond_32768_rows(env):
    """Panels taller than 32768 rows use 8 rows x 4 columns per thread in the leaf cluster (capacity 65536):
    a 33 280-row system (8.9 GB) is factored and solved on the device, the solution satisfies the system."""
    torch, lib, _lib = env
    n = 33280
    gen = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=gen)
    A0 = A.clone()
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
    x = b.clone()
    ipiv = torch.zeros(n, dtype=torch.int32, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    plan = C.c_void_p()
    _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)), "wg_lu_create")
    try:
        _lib.check(lib.wg_lu_factor(plan, A.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), _stream(torch)), "factor")
        _lib.check(lib.wg_lu_solve(plan, A.data_ptr(), n, x.data_ptr(), _stream(torch)), "solve")
        torch.cuda.synchronize()
    finally:
        lib.wg_lu_destroy(plan)
    assert int(info.item()) == 0
    p = ipiv.cpu().numpy()
    assert np.all(p >= np.arange(n)) and p.max() < n
    r = torch.mv(A0, x) - b
    rel = float(torch.linalg.norm(r) / (torch.linalg.norm(A0) * torch.linalg.norm(x)))
    assert rel < 1e-14, rel
    assert float(torch.max(torch.abs(torch.tril(A, -1)))) <= 1.0 + 1e-12      # partial pivoting bounds the multipliers


@pytest.mark.parametrize("wb", [2, 4, 8])
def test_lu_tall_panel_leaf_variants_at_small_size(env, wb):
    """The leaf variants meant for very tall panels (8 columns x up to 4 rows per thread beyond 16384 rows,
    4 columns x 8 rows beyond 32768) forced on a small matrix: same pivots as the default blocking, factors
    equal to rounding (the blocking differs, so not bit for bit)."""
    torch, lib, _ = env
    n = 700
    rng = np.random.default_rng(17)
    A = rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    LU0, ipiv0, info0, x0 = _lu(env, A, b)
    try:
        lib.wg_debug_force_wb(wb)
        LU1, ipiv1, info1, x1 = _lu(env, A, b)
    finally:
        lib.wg_debug_force_wb(0)
    assert info0 == info1 == 0
    assert np.array_equal(ipiv0, ipiv1)
    assert np.max(np.abs(LU0 - LU1)) / np.max(np.abs(LU0)) < 1e-11
    assert np.max(np.abs(x0 - x1)) / np.max(np.abs(x0)) < 1e-9
    _check_factors(A, LU1, ipiv1)


# ---------------------------------------------------------------- strided batches (lock-step LU of many systems)
def _lu_batched(env, As, bs, pad=0, graph=None):
    """Factor + solve a list of equally sized systems with ONE batched plan; returns (LU[B], ipiv[B], info[B], x[B])."""
    torch, lib, _lib = env
    B, n = len(As), As[0].shape[0]
    lda = n + pad
    strideA = n * lda + 4 * (pad > 0)
    buf = np.zeros(B * strideA)
    for i, A in enumerate(As):
        buf[i * strideA:i * strideA + n * lda].reshape(n, lda)[:, :n] = A
    Ad = torch.from_numpy(buf).cuda()
    bd = torch.from_numpy(np.stack(bs).copy()).cuda()
    ipiv = torch.zeros(B * n, dtype=torch.int32, device="cuda")
    info = torch.zeros(B, dtype=torch.int32, device="cuda")
    plan = C.c_void_p()
    _lib.check(lib.wg_lu_create_batched(n, B, 0, C.byref(plan)), "wg_lu_create_batched")
    try:
        for _ in range(2 if graph else 1):      # second pass replays the captured graph on restored data
            Ad.copy_(torch.from_numpy(buf))
            bd.copy_(torch.from_numpy(np.stack(bs)))
            _lib.check(lib.wg_lu_factor_batched(plan, Ad.data_ptr(), lda, strideA, ipiv.data_ptr(), info.data_ptr(),
                                                _stream(torch)), "wg_lu_factor_batched")
            _lib.check(lib.wg_lu_solve_batched(plan, Ad.data_ptr(), lda, strideA, bd.data_ptr(), n, _stream(torch)),
                       "wg_lu_solve_batched")
            torch.cuda.synchronize()
    finally:
        lib.wg_lu_destroy(plan)
    out = Ad.cpu().numpy()
    LUs = [out[i * strideA:i * strideA + n * lda].reshape(n, lda)[:, :n].copy() for i in range(B)]
    return LUs, ipiv.cpu().numpy().reshape(B, n), info.cpu().numpy(), bd.cpu().numpy()


@pytest.mark.parametrize("n,B,pad", [(37, 5, 0), (300, 7, 0), (513, 3, 2), (1900, 6, 0), (700, 40, 0), (1100, 60, 0)])
def test_lu_batched_is_bit_identical_to_single(env, n, B, pad):
    """Every system of a lock-step batch gets exactly the factors, pivots and solution of the single-system
    path (same kernels, the batch is only a grid dimension) and agrees with LAPACK."""
    rng = np.random.default_rng(n * 31 + B)
    As = [rng.standard_normal((n, n)) + (i % 3) * np.eye(n) for i in range(B)]
    bs = [rng.standard_normal(n) for _ in range(B)]
    LUs, ipivs, infos, xs = _lu_batched(env, As, bs, pad=pad, graph=True)
    assert not infos.any()
    for i in (0, B // 2, B - 1):
        LU1, ipiv1, info1, x1 = _lu(env, As[i], bs[i], lda=n + pad)
        assert info1 == 0
        assert np.array_equal(ipivs[i], ipiv1)
        assert np.array_equal(LUs[i], LU1)
        assert np.array_equal(xs[i], x1)
    for i in range(B):
        ref = np.linalg.solve(As[i], bs[i])
        assert np.max(np.abs(xs[i] - ref)) <= 1e-9 * np.max(np.abs(ref))


def test_lu_batched_singular_member(env):
    """One singular system in a batch: its info is set LAPACK-style, the others are unaffected."""
    n, B = 200, 4
    rng = np.random.default_rng(5)
    As = [rng.standard_normal((n, n)) for _ in range(B)]
    As[2][:, 17] = 0.0
    bs = [rng.standard_normal(n) for _ in range(B)]
    LUs, ipivs, infos, xs = _lu_batched(env, As, bs)
    assert list(infos) == [0, 0, 18, 0]
    for i in (0, 1, 3):
        ref = np.linalg.solve(As[i], bs[i])
        assert np.max(np.abs(xs[i] - ref)) <= 1e-9 * np.max(np.abs(ref))


@pytest.mark.parametrize("n", [1, 31, 128, 129, 300, 1900, 5000])
def test_trisolve_wavefront_equals_block_steps(env, n):
    """The persistent wavefront form of the triangular solves (two launches, ticket + published counter, diagonal
    blocks applied through their inverses) is reproducible bit for bit on streams and as a graph, run after run;
    it agrees with the one-launch-per-block substitution form to rounding, and solves the system."""
    torch, lib, _ = env
    rng = np.random.default_rng(11 * n + 1)
    A = rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    outs = {}
    try:
        for key, (wave, graph) in enumerate(((1, 1), (0, 1), (1, 0), (1, 1))):
            lib.wg_debug_wave_solve(wave)
            lib.wg_set_tuning(256, graph)
            LU, ipiv, info, x = _lu(env, A, b)
            assert info == 0
            outs[key] = x
    finally:
        lib.wg_debug_wave_solve(1)
        lib.wg_set_tuning(256, 1)
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[0], outs[3])
    xr = np.linalg.solve(A, b)
    scale = np.max(np.abs(xr))
    assert np.max(np.abs(outs[0] - outs[1])) / scale < 1e-9
    assert np.max(np.abs(outs[0] - xr)) / scale < 1e-9
    assert np.max(np.abs(outs[1] - xr)) / scale < 1e-9


def test_trisolve_wavefront_many_right_hand_sides(env):
    """Repeated solves with one set of factors (checkLinearization re-uses them): the counters are reset by every
    call, results repeat exactly."""
    torch, lib, _lib = env
    n = 2000
    rng = np.random.default_rng(3)
    A = rng.standard_normal((n, n))
    Ad = torch.from_numpy(A.copy()).cuda()
    ipiv = torch.zeros(n, dtype=torch.int32, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    plan = C.c_void_p()
    _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)), "wg_lu_create")
    try:
        _lib.check(lib.wg_lu_factor(plan, Ad.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), _stream(torch)), "factor")
        first = None
        for rep in range(5):
            bs = [rng.standard_normal(n) for _ in range(3)] if rep == 0 else bs
            xs = []
            for b in bs:
                bd = torch.from_numpy(b.copy()).cuda()
                _lib.check(lib.wg_lu_solve(plan, Ad.data_ptr(), n, bd.data_ptr(), _stream(torch)), "solve")
                xs.append(bd.cpu().numpy())
            if first is None:
                first = xs
                for b, x in zip(bs, xs):
                    assert np.max(np.abs(A @ x - b)) < 1e-8 * np.max(np.abs(x)) * n ** 0.5
            else:
                for x0, x in zip(first, xs):
                    assert np.array_equal(x0, x)
    finally:
        lib.wg_lu_destroy(plan)


def test_solve_before_factor_is_a_state_error(env):
    """The triangular solves need the diagonal-block inverses a factorisation of the same plan leaves behind:
    solving first is reported (WG_ERR_STATE), not computed from uninitialised memory."""
    torch, lib, _lib = env
    n = 64
    A = torch.eye(n, dtype=torch.float64, device="cuda")
    b = torch.ones(n, dtype=torch.float64, device="cuda")
    plan = C.c_void_p()
    _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)), "wg_lu_create")
    try:
        assert lib.wg_lu_solve(plan, A.data_ptr(), n, b.data_ptr(), _stream(torch)) == -3
        assert b"no factorisation" in lib.wg_last_error()
        ipiv = torch.zeros(n, dtype=torch.int32, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(lib.wg_lu_factor(plan, A.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), _stream(torch)), "factor")
        _lib.check(lib.wg_lu_solve(plan, A.data_ptr(), n, b.data_ptr(), _stream(torch)), "solve")
        torch.cuda.synchronize()
        assert torch.equal(b, torch.ones(n, dtype=torch.float64, device="cuda"))
    finally:
        lib.wg_lu_destroy(plan)
