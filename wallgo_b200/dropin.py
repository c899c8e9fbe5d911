"""Drop-in for `WallGo.BoltzmannSolver` (only importable where WallGo itself is installed).

    import WallGo, wallgo_b200.dropin as b200
    b200.install()            # WallGoManager.setupWallSolver now builds the B200 solver
    manager = WallGo.WallGoManager(); ...; manager.solveWall(settings)

`install()` rebinds the name `BoltzmannSolver` in `WallGo` and `WallGo.manager`
(the single construction site, manager.py:605-611).  The class is a real subclass of the
reference class, so `isinstance` (equationOfMotion.py:106), the shared-grid assertion (:110-112)
and `copy.deepcopy` + attribute rewrites in `getBoltzmannFiniteDifference` (:2135-2143) keep working,
and results are WallGo's own `BoltzmannResults` / `BoltzmannDeltas` / `Polynomial`.
"""
from __future__ import annotations

from .boltzmann import BoltzmannSolverB200Mixin

_cached = None


def makeDropInClass():
    """Returns the subclass of WallGo.BoltzmannSolver backed by the B200 engine."""
    global _cached
    if _cached is not None:
        return _cached
    import WallGo  # pylint: disable=import-outside-toplevel
    from WallGo.boltzmann import BoltzmannSolver as _RefSolver, ETruncationOption
    from WallGo.collisionArray import CollisionArray
    import WallGo.results  # noqa: F401  pylint: disable=import-outside-toplevel

    class _WallGoTypes:
        Polynomial = WallGo.Polynomial
        BoltzmannDeltas = WallGo.BoltzmannDeltas
        BoltzmannResults = WallGo.results.BoltzmannResults

    class B200BoltzmannSolver(BoltzmannSolverB200Mixin, _RefSolver):
        """WallGo.BoltzmannSolver with every numerical step on a B200."""

        _types = _WallGoTypes

        def __init__(self, grid, basisM="Cardinal", basisN="Chebyshev", derivatives="Spectral",
                     collisionMultiplier=1.0, truncationOption=ETruncationOption.AUTO, device=None):
            self._initB200(grid, basisM, basisN, derivatives, collisionMultiplier, truncationOption, device)

        def loadCollisions(self, directoryPath):
            """File ingest stays the reference's own host code (collisionArray.py:178-354); the
            tensor it returns is uploaded to the GPU at the next solve."""
            self.collisionArray = CollisionArray.newFromDirectory(
                directoryPath, self.grid, self.basisN, self.offEqParticles)

    _cached = B200BoltzmannSolver
    return _cached


def install(vectoriseEOM: bool = False, reuseFactors: bool = False):
    """Make WallGoManager construct the B200 solver (manager.py:605-611).

    reuseFactors=True: solvers built from now on re-use LU factors across the Boltzmann solves of one
    `EOM.wallPressure` (same wall velocity and grid mapping): from the third solve on, refinement steps
    preconditioned with the last factorisation replace the LU (BoltzmannSolverB200Mixin._solveMaybeStale);
    deltaF then agrees with the plain path to ~1e-13 instead of bit for bit.

    vectoriseEOM=True additionally rebinds `WallGo.EOM.findPlasmaProfile` to the lock-step vector form of
    `wallgo_b200.eomhost` (the consumer of the moments and, once the Boltzmann solve
    runs on the GPU, the largest host cost of a solveWall).  It follows SciPy's scalar solvers iterate for
    iterate; profiles agree with the reference to ~1e-9, the wall velocity to better than 1e-6."""
    import WallGo  # pylint: disable=import-outside-toplevel
    import WallGo.manager as _manager

    cls = makeDropInClass()
    cls.reuseFactorsDefault = bool(reuseFactors)
    _manager.BoltzmannSolver = cls
    WallGo.BoltzmannSolver = cls
    if vectoriseEOM:
        from . import eomhost
        if not hasattr(WallGo.EOM, "_findPlasmaProfileReference"):
            eomhost.install(WallGo.EOM)
    return cls
