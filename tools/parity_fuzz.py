"""Developer fuzz: the CUDA path against the CPU oracle over random sizes, bases, derivative and truncation
options (tests/ and tools/ may use the oracle; the product never does)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import wallgo_b200 as wb
from conftest import FixtureBackground, FixtureGrid, relmax
from oracle import boltzmann_oracle as bo
ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng0 = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
worst = 0.0
for case in range(ncases):
    seed = int(rng0.integers(1 << 30)); rng = np.random.default_rng(seed)
    P = int(rng.integers(1, 5)); M = int(rng.integers(3, 15)); N = int(rng.choice([3, 5, 7, 9, 11, 13]))
    deriv = str(rng.choice(["Spectral", "Spectral", "Finite Difference"]))
    if deriv == "Finite Difference":
        basisM = basisN = "Cardinal"
    else:
        basisM = str(rng.choice(["Cardinal", "Chebyshev"])); basisN = str(rng.choice(["Chebyshev", "Cardinal"]))
    option = str(rng.choice(["AUTO", "AUTO", "NONE", "THIRD"]))
    kw = dict(seed=seed % 1000, r=float(rng.choice([0.5, 0.05])), collisionMultiplier=float(rng.choice([1.0, 1e-2, 10.0])),
              vw_mid=-float(rng.uniform(0.05, 0.9)), basisM=basisM, basisN=basisN, derivatives=deriv, grid3=bool(rng.random() < 0.7))
    print(f"case {case}: P={P} M={M} N={N} {basisM}/{basisN}/{deriv} {option} {kw}", flush=True)
    pb = bo.synthetic_problem(P, M, N, **kw)
    g = dict(P=P, M=M, N=N, chi=pb.tab.chi, rz=pb.tab.rz, rp=pb.tab.rp, pz=pb.pz, pp=pb.pp,
             dxidchi=pb.xi_dxidchi, dpzdrz=pb.dpzdrz, dppdrp=pb.dppdrp, vw=pb.vw, v=pb.v, T=pb.T, msq=pb.msq)
    grid = FixtureGrid(g)
    parts = [wb.Particle(f"p{a}", a, (lambda f, a=a: f.msq[a]), None, "Fermion" if pb.statistics[a] < 0 else "Boson", 12)
             for a in range(P)]
    solver = wb.BoltzmannSolver(grid, basisM, basisN, deriv, pb.collisionMultiplier, getattr(wb.ETruncationOption, option))
    solver.updateParticleList(parts)
    solver.setBackground(FixtureBackground(g))
    solver.setCollisionArray(wb.CollisionArray(grid, basisN, parts, pb.collision))
    # reference answer: the EXTENDED-PRECISION solution of the assembled system pushed through the oracle's
    # truncation + moments (LAPACK's own FP64 solve is off by cond * eps, up to 4e-11 in this fuzz: a yardstick
    # no more accurate than the thing measured).  The device applies one mixed-precision refinement step (wg_refine).
    A_ref, b_ref = bo.assemble(pb)
    xExt, xLap = bo.solve_extended(A_ref, b_ref)
    ref = bo.get_deltas(pb, xExt.reshape(P, M - 1, N - 1, N - 1), option)
    res = solver.getDeltas()
    e = relmax(res.deltaF, ref["deltaF"])
    D = np.stack([res.Deltas.Delta00.coefficients, res.Deltas.Delta02.coefficients, res.Deltas.Delta20.coefficients,
                  res.Deltas.Delta11.coefficients])
    ed = max(relmax(D[i], ref["Deltas"][i]) for i in range(4))
    same = tuple(res.truncatedTail) == tuple(ref["truncatedTail"]) and tuple(res.spectralPeaks) == tuple(ref["spectralPeaks"])
    worst = max(worst, e, ed)
    # operator / source, the spectral queries on a foreign deltaF, and the linearisation criteria
    extra = []
    if case % 2 == 0:
        op_ref, src_ref = bo.assemble(pb)
        op, src, _, _ = solver.buildLinearEquations()
        if relmax(op, op_ref) > 1e-12 or relmax(src, src_ref) > 1e-12:
            extra.append(f"operator {relmax(op, op_ref):.1e} source {relmax(src, src_ref):.1e}")
        dFx = rng.standard_normal(ref["deltaF"].shape) * np.exp(rng.uniform(-0.5, 0.5) * np.arange(M - 1))[None, :, None, None]
        r_dF, r_shape, r_peaks = bo.check_spectral_convergence(dFx, pb.tab, option)
        g_dF, g_shape, g_peaks = solver.checkSpectralConvergence(dFx)
        if tuple(g_shape) != tuple(r_shape) or tuple(g_peaks) != tuple(r_peaks) or relmax(g_dF, r_dF) > 1e-11:
            extra.append(f"checkSpectralConvergence shape {g_shape} vs {r_shape} peaks {g_peaks} vs {r_peaks} dF {relmax(g_dF, r_dF):.1e}")
        with np.errstate(all="ignore"):
            te_ref = bo.estimate_truncation_error(r_dF, r_shape, pb.tab)
            te = solver.estimateTruncationError(g_dF, g_shape)
        if np.isfinite(te_ref) and abs(te - te_ref) > 1e-9 * max(abs(te_ref), 1e-300):
            extra.append(f"estimateTruncationError {te} vs {te_ref}")
    if case % 3 == 0 and min(M, N) >= 5:
        c_ref = bo.check_linearization(pb)
        c_gpu = solver.checkLinearization()
        if any(abs(a - b) > 1e-7 * abs(b) for a, b in zip(c_gpu, c_ref)):
            extra.append(f"checkLinearization {c_gpu} vs {c_ref}")
    for msg in extra:
        bad += 1
        print("  MISMATCH", msg)
    if not same or e > 1e-10 or ed > 1e-10:
        bad += 1
        print(f"  MISMATCH deltaF {e:.2e} Deltas {ed:.2e} flags {res.truncatedTail} vs {ref['truncatedTail']} peaks {res.spectralPeaks} vs {ref['spectralPeaks']}")
print(f"{ncases} cases, {bad} bad, worst relative deviation {worst:.2e}")
sys.exit(1 if bad else 0)
