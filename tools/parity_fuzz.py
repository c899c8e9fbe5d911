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
    ref = bo.get_deltas(pb, None, option)
    res = solver.getDeltas()
    e = relmax(res.deltaF, ref["deltaF"])
    D = np.stack([res.Deltas.Delta00.coefficients, res.Deltas.Delta02.coefficients, res.Deltas.Delta20.coefficients,
                  res.Deltas.Delta11.coefficients])
    ed = max(relmax(D[i], ref["Deltas"][i]) for i in range(4))
    same = tuple(res.truncatedTail) == tuple(ref["truncatedTail"]) and tuple(res.spectralPeaks) == tuple(ref["spectralPeaks"])
    worst = max(worst, e, ed)
    if not same or e > 1e-10 or ed > 1e-10:
        bad += 1
        print(f"  MISMATCH deltaF {e:.2e} Deltas {ed:.2e} flags {res.truncatedTail} vs {ref['truncatedTail']} peaks {res.spectralPeaks} vs {ref['spectralPeaks']}")
print(f"{ncases} cases, {bad} bad, worst relative deviation {worst:.2e}")
sys.exit(1 if bad else 0)
