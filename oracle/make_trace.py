"""
Records a golden TRACE of the Boltzmann hot path as the reference's own wall solver drives it.

Runs ONLY in the build container (needs /root/reference).  The unmodified reference
(`WallGoManager.solveWall` on the shipped Yukawa example model, off-equilibrium ON) is run with a
synthetic relaxation-time collision tensor injected in place of the Git-LFS HDF5 files (SURVEY.md
Appendix B).  Every `BoltzmannSolver.getDeltas()` call is recorded: the boosted background handed to
the solver (inputs of the hot path) and deltaF / the four Delta moments / truncation outputs it
returned.  tests/ replay the inputs through the CUDA path (and the oracle) and compare call by call.

usage: python oracle/make_trace.py [M] [N]
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, "/root/reference/src")
sys.dont_write_bytecode = True

import WallGo  # noqa: E402
from WallGo import Polynomial  # noqa: E402
from WallGo.collisionArray import CollisionArray  # noqa: E402


def load_yukawa():
    spec = importlib.util.spec_from_file_location("yukawa_example", "/root/reference/Models/Yukawa/yukawa.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def rta_collisions(solver, r=0.05):
    """C = r * identity in the Cardinal basis (a relaxation-time operator), changed to the solver's basis."""
    grid, parts = solver.grid, solver.offEqParticles
    P, n1 = len(parts), grid.N - 1
    C = (r * np.identity(P * n1 * n1)).reshape(P, n1, n1, P, n1, n1)
    poly = Polynomial(C, grid, ("Array", "Cardinal", "Cardinal", "Array", "Cardinal", "Cardinal"),
                      CollisionArray.AXIS_TYPES, endpoints=False)
    return CollisionArray.newFromPolynomial(poly, parts).changeBasis(solver.basisN)


def build_manager(M, N, solverClass=None):
    yk = load_yukawa()
    if solverClass is not None:
        import WallGo.manager as _m
        _m.BoltzmannSolver = solverClass
    manager = WallGo.WallGoManager()
    manager.config.configGrid.spatialGridSize = M
    manager.config.configGrid.momentumGridSize = N
    manager.config.configEOM.maxIterations = 25
    manager.config.configThermodynamics.phaseTracerTol = 1e-8
    manager.setPathToCollisionData(os.path.join(ROOT, "does_not_exist"))
    model = yk.YukawaModel()
    manager.registerModel(model)
    model.modelParameters.update({"sigma": 0.0, "msq": 1.0, "gamma": -1.2, "lam": 0.10, "y": 0.55, "mf": 0.30})
    manager.setupThermodynamicsHydrodynamics(
        WallGo.PhaseInfo(temperature=8.0, phaseLocation1=WallGo.Fields([0.4]), phaseLocation2=WallGo.Fields([27.0])),
        WallGo.VeffDerivativeSettings(temperatureVariationScale=1.0, fieldValueVariationScale=[100.0]))
    settings = WallGo.WallSolverSettings(bIncludeOffEquilibrium=True, meanFreePathScale=100.0,
                                         wallThicknessGuess=10.0)
    return manager, settings


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    ref = WallGo.boltzmann.BoltzmannSolver
    calls = []
    holder = {}
    orig_load, orig_get = ref.loadCollisions, ref.getDeltas

    def load(self, path):
        self.setCollisionArray(rta_collisions(self))

    def get(self, deltaF=None):
        res = orig_get(self, deltaF)
        if deltaF is None and self.derivatives == "Spectral":
            holder["solver"] = self
            bg, g = self.background, self.grid
            calls.append(dict(
                T=np.array(bg.temperatureProfile), v=np.array(bg.velocityProfile), vw=float(bg.velocityWall),
                msq=np.array([p.msqVacuum(bg.fieldProfiles) for p in self.offEqParticles]),
                dxidchi=np.array(g.getCompactificationDerivatives()[0]),
                deltaF=np.array(res.deltaF),
                Deltas=np.stack([res.Deltas.Delta00.coefficients, res.Deltas.Delta02.coefficients,
                                 res.Deltas.Delta20.coefficients, res.Deltas.Delta11.coefficients]),
                truncErr=float(res.truncationError), tail=np.array(res.truncatedTail),
                peaks=np.array(res.spectralPeaks)))
        return res

    ref.loadCollisions, ref.getDeltas = load, get
    try:
        manager, settings = build_manager(M, N)
        results = manager.solveWall(settings)
    finally:
        ref.loadCollisions, ref.getDeltas = orig_load, orig_get
    solver = holder["solver"]
    g = solver.grid
    print(f"solveWall: success={results.success} vw={results.wallVelocity!r} widths={results.wallWidths} "
          f"calls recorded={len(calls)}")
    out = dict(
        M=M, N=N, P=len(solver.offEqParticles), nCalls=len(calls),
        chi=g.chiValues, rz=g.rzValues, rp=g.rpValues, pz=g.pzValues, pp=g.ppValues, dpzdrz=g.dpzdrz,
        dppdrp=g.dppdrp, momentumFalloffT=g.momentumFalloffT,
        statistics=np.array([-1.0 if p.statistics == "Fermion" else 1.0 for p in solver.offEqParticles]),
        collision=np.array(solver.collisionArray[...]), collisionMultiplier=float(solver.collisionMultiplier),
        wallVelocity=float(results.wallVelocity), wallWidths=np.array(results.wallWidths),
        wallOffsets=np.array(results.wallOffsets), success=bool(results.success))
    for key in ("T", "v", "vw", "msq", "dxidchi", "deltaF", "Deltas", "truncErr", "tail", "peaks"):
        out["call_" + key] = np.stack([c[key] for c in calls])
    path = os.path.join(ROOT, "tests", "golden", f"trace_yukawa_M{M}_N{N}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
