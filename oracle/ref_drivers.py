"""Test infrastructure (never shipped, never imported by wallgo_b200/): loads the UNMODIFIED reference
from oracle/_ref (a copy made by oracle/fetch_ref.sh; /root/reference/src when that copy is absent and the
build container's tree is there) and builds the four example drivers BASELINE.json's configs name, the way
the reference's own `WallGoExampleBase.runExample` does (Models/wallGoExampleBase.py:218-428), minus the
command line and the equilibrium-only pre-run.

Every shipped collision .hdf5 is a Git-LFS pointer and there is no HDF5 library in this image, so
`BoltzmannSolver.loadCollisions` (boltzmann.py:822-844, called from manager.py:619) is replaced by an in-memory
relaxation-time tensor C = r*identity built with the reference's own `CollisionArray.newFromPolynomial`
 -- for BOTH arms of every comparison.
"""
from __future__ import annotations

import argparse
import copy
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

EXAMPLES = {  # name -> (directory, file, example class or None for the plain Yukawa script)
    "yukawa": ("Yukawa", "yukawa.py", None),
    "singlet": ("SingletStandardModel_Z2", "singletStandardModelZ2.py", "SingletStandardModelExample"),
    "sm": ("StandardModel", "standardModel.py", "StandardModelExample"),
    "idm": ("InertDoubletModel", "inertDoubletModel.py", "InertDoubletModelExample"),
}


def reference_root():
    """Directory holding src/WallGo and Models/, or None."""
    for cand in (os.path.join(HERE, "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "src", "WallGo")):
            return cand
    return None


def load_reference():
    """import WallGo from the reference copy (h5py replaced by the empty stand-in of oracle/stubs)."""
    root = reference_root()
    if root is None:
        raise ImportError("the reference is not available: run oracle/fetch_ref.sh in the build container")
    try:
        import h5py  # noqa: F401  pylint: disable=unused-import,import-outside-toplevel
    except ImportError:
        stubs = os.path.join(HERE, "stubs")
        if stubs not in sys.path:
            sys.path.insert(0, stubs)
    src = os.path.join(root, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    models = os.path.join(root, "Models")
    if models not in sys.path:
        sys.path.insert(0, models)    # the examples do `from wallGoExampleBase import ...`
    sys.dont_write_bytecode = True
    import WallGo  # pylint: disable=import-outside-toplevel
    return WallGo


def _load_module(name):
    root = reference_root()
    d, f, _ = EXAMPLES[name]
    path = os.path.join(root, "Models", d, f)
    modname = "wgref_example_" + name
    if modname in sys.modules:
        return sys.modules[modname]
    exdir = os.path.dirname(path)
    if exdir not in sys.path:
        sys.path.insert(0, exdir)     # exampleCollisionDefs.py next to the example
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def rta_collisions(solver, r=0.05):
    """C = r * identity in the Cardinal basis (relaxation-time operator), changed to the solver's basis with
    the reference's own CollisionArray (collisionArray.py:137-174, 356-388)."""
    WallGo = load_reference()
    from WallGo.collisionArray import CollisionArray  # pylint: disable=import-outside-toplevel

    grid, parts = solver.grid, solver.offEqParticles
    P, n1 = len(parts), grid.N - 1
    C = (r * np.identity(P * n1 * n1)).reshape(P, n1, n1, P, n1, n1)
    poly = WallGo.Polynomial(C, grid, ("Array", "Cardinal", "Cardinal", "Array", "Cardinal", "Cardinal"),
                             CollisionArray.AXIS_TYPES, endpoints=False)
    return CollisionArray.newFromPolynomial(poly, parts).changeBasis(solver.basisN)


def patch_load_collisions(cls, r=0.05):
    """Replace cls.loadCollisions by the in-memory tensor; returns the original for restoring."""
    orig = cls.__dict__.get("loadCollisions")

    def load(self, directoryPath):  # pylint: disable=unused-argument
        self.setCollisionArray(rta_collisions(self, r))

    cls.loadCollisions = load
    return orig


def build_manager(name, M, N, benchmarkIndex=None):
    """(manager, wallSolverSettings, model) for one example, thermodynamics and hydrodynamics set up, ready
    for manager.solveWall(settings) / manager.setupWallSolver(settings)."""
    WallGo = load_reference()
    mod = _load_module(name)
    _, _, clsName = EXAMPLES[name]
    if clsName is None:      # Models/Yukawa/yukawa.py:149-224 (a plain script with a main())
        manager = WallGo.WallGoManager()
        manager.config.configGrid.spatialGridSize = M
        manager.config.configGrid.momentumGridSize = N
        manager.config.configEOM.maxIterations = 25
        manager.config.configThermodynamics.phaseTracerTol = 1e-8
        manager.setPathToCollisionData(os.path.join(ROOT, "does_not_exist"))
        model = mod.YukawaModel()
        manager.registerModel(model)
        model.modelParameters.update({"sigma": 0.0, "msq": 1.0, "gamma": -1.2, "lam": 0.10, "y": 0.55, "mf": 0.30})
        manager.setupThermodynamicsHydrodynamics(
            WallGo.PhaseInfo(temperature=8.0, phaseLocation1=WallGo.Fields([0.4]),
                             phaseLocation2=WallGo.Fields([27.0])),
            WallGo.VeffDerivativeSettings(temperatureVariationScale=1.0, fieldValueVariationScale=[100.0]))
        settings = WallGo.WallSolverSettings(bIncludeOffEquilibrium=True, meanFreePathScale=100.0,
                                             wallThicknessGuess=10.0)
        return manager, settings, model
    ex = getattr(mod, clsName)()
    ex.cmdArgs = argparse.Namespace(momentumGridSize=N, verbose=30, recalculateMatrixElements=False,
                                    recalculateCollisions=False, includeDetonations=False,
                                    skipEquilibriumEOM=True, outOfEquilibriumGluon=False)
    manager = WallGo.WallGoManager()
    ex.configureManager(manager)
    manager.config.configGrid.spatialGridSize = M
    manager.config.configGrid.momentumGridSize = N
    model = ex.initWallGoModel()
    manager.registerModel(model)
    points = ex.getBenchmarkPoints()
    if benchmarkIndex is None:
        benchmarkIndex = 2 if name == "sm" else 0     # StandardModel benchmark #3
    bm = points[benchmarkIndex]
    ex.updateModelParameters(model, bm.inputParameters)
    manager.setupThermodynamicsHydrodynamics(bm.phaseInfo, bm.veffDerivativeScales)
    manager.setPathToCollisionData(os.path.join(ROOT, "does_not_exist"))
    settings = copy.deepcopy(bm.wallSolverSettings)
    settings.bIncludeOffEquilibrium = True
    return manager, settings, model


def result_vector(res):
    """[vw, widths..., offsets...] of a WallGoResults."""
    return np.concatenate([[float(res.wallVelocity)], np.ravel(res.wallWidths), np.ravel(res.wallOffsets)])


def synthetic_solver(P, M, N, basisM="Cardinal", basisN="Chebyshev", derivatives="Spectral", option="AUTO",
                     r=0.5, seed=0, mult=1.0, grid3=True, solverClass=None, velocityMid=-0.5):
    """The deterministic synthetic system of SURVEY.md section 8(d) built from the reference's OWN classes
    (grid, particles, background, collision array); `solverClass` defaults to the reference BoltzmannSolver.
    Returns (solver, grid, background, particles)."""
    WallGo = load_reference()
    from WallGo.boltzmann import ETruncationOption  # pylint: disable=import-outside-toplevel
    from WallGo.collisionArray import CollisionArray  # pylint: disable=import-outside-toplevel

    if grid3:
        grid = WallGo.Grid3Scales(M, N, 0.2, 0.2, 0.05, 100.0, 0.5, 0.1)
    else:
        grid = WallGo.Grid(M, N, 0.05, 100.0)
    L = 0.05
    xi = grid.getCoordinates(endpoints=True)[0]
    th = np.tanh(xi / L)
    phi = 50 * (1 - th)
    T = 100 + th
    v = velocityMid + 0.02 * th
    particles = []
    for a in range(P):
        particles.append(WallGo.Particle(
            name=f"p{a}", index=a,
            msqVacuum=(lambda f, a=a: 0.5 * (1 + 0.3 * a) * f.getField(0) ** 2),
            msqDerivative=(lambda f, a=a: np.transpose([(1 + 0.3 * a) * f.getField(0)])),
            statistics="Fermion" if a % 2 == 0 else "Boson", totalDOFs=12))
    bg = WallGo.BoltzmannBackground(velocityMid, v, WallGo.Fields(phi[:, None]), T)
    cls = WallGo.BoltzmannSolver if solverClass is None else solverClass
    solver = cls(grid, basisM, basisN, derivatives, mult, getattr(ETruncationOption, option))
    solver.updateParticleList(particles)
    solver.setBackground(bg)
    rng = np.random.default_rng(seed)
    q = (N - 1) ** 2
    C = (r * np.identity(P * q) + 1e-3 * rng.standard_normal((P * q, P * q))).reshape(
        P, N - 1, N - 1, P, N - 1, N - 1)
    poly = WallGo.Polynomial(C.copy(), grid,
                             ("Array", "Cardinal", "Cardinal", "Array", "Cardinal", "Cardinal"),
                             CollisionArray.AXIS_TYPES, endpoints=False)
    ca = CollisionArray.newFromPolynomial(poly, particles).changeBasis(basisN)
    solver.setCollisionArray(ca)
    return solver, grid, bg, particles


def deltas_table(res):
    """(4, P, M-1) array of the four moments of a BoltzmannResults."""
    d = res.Deltas
    return np.stack([d.Delta00.coefficients, d.Delta02.coefficients, d.Delta20.coefficients, d.Delta11.coefficients])
