"""Host-side mirror of the reference's boundary types for the Boltzmann path.

When WallGo itself is importable the drop-in (wallgo_b200/dropin.py) returns WallGo's own
`Polynomial`, `BoltzmannDeltas`, `BoltzmannResults`; these stand-ins are what the standalone
class returns on a machine without WallGo (e.g. the benchmark box).  They keep the same field
names and the same `scalar*res`, `res+res`, `res-res` arithmetic the EOM loop relies on
(reference containers.py:167-227, results.py:13-85, equationOfMotion.py:1352-1355).
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum, auto

import numpy as np


class CollisionLoadError(Exception):
    """Raised when collision integrals fail to load (mirror of WallGo.exceptions.CollisionLoadError,
    exceptions.py:52-55; the drop-in raises WallGo's own class)."""


class ETruncationOption(Enum):
    """Mirror of WallGo.boltzmann.ETruncationOption (boltzmann.py:24-34)."""
    NONE = auto()
    AUTO = auto()
    THIRD = auto()


class Polynomial:
    """Minimal stand-in: coefficient array + basis/direction/endpoints labels (polynomial.py:24-85)."""

    def __init__(self, coefficients, grid, basis="Cardinal", direction="z", endpoints=False):
        self.coefficients = np.asanyarray(coefficients)
        self.rank = self.coefficients.ndim
        self.grid = grid
        self.basis = (basis,) * self.rank if isinstance(basis, str) else tuple(basis)
        self.direction = (direction,) * self.rank if isinstance(direction, str) else tuple(direction)
        self.endpoints = (endpoints,) * self.rank if isinstance(endpoints, bool) else tuple(endpoints)

    def _like(self, coeffs):
        return Polynomial(coeffs, self.grid, self.basis, self.direction, self.endpoints)

    def __mul__(self, other):
        if isinstance(other, Polynomial):
            return self._like(self.coefficients * other.coefficients)
        return self._like(other * self.coefficients)

    __rmul__ = __mul__

    def __add__(self, other):
        if isinstance(other, Polynomial):
            return self._like(self.coefficients + other.coefficients)
        return self._like(other + self.coefficients)

    __radd__ = __add__

    def __sub__(self, other):
        return self.__add__((-1.0) * other)


@dataclass
class BoltzmannDeltas:
    Delta00: Polynomial  # pylint: disable=invalid-name
    Delta02: Polynomial  # pylint: disable=invalid-name
    Delta20: Polynomial  # pylint: disable=invalid-name
    Delta11: Polynomial  # pylint: disable=invalid-name

    def __mul__(self, number):
        return BoltzmannDeltas(number * self.Delta00, number * self.Delta02, number * self.Delta20,
                               number * self.Delta11)

    __rmul__ = __mul__

    def __add__(self, other):
        return BoltzmannDeltas(other.Delta00 + self.Delta00, other.Delta02 + self.Delta02,
                               other.Delta20 + self.Delta20, other.Delta11 + self.Delta11)

    def __sub__(self, other):
        return self.__add__((-1) * other)


@dataclass
class BoltzmannResults:
    deltaF: np.ndarray
    Deltas: BoltzmannDeltas  # pylint: disable=invalid-name
    truncationError: float
    truncatedTail: tuple
    spectralPeaks: tuple
    linearizationCriterion1: float = 0
    linearizationCriterion2: float = 0

    def __mul__(self, number):
        return BoltzmannResults(
            deltaF=number * self.deltaF, Deltas=number * self.Deltas,
            truncationError=abs(number) * self.truncationError, truncatedTail=self.truncatedTail,
            spectralPeaks=self.spectralPeaks,
            linearizationCriterion1=abs(number) * self.linearizationCriterion1,
            linearizationCriterion2=self.linearizationCriterion2)

    __rmul__ = __mul__

    def __add__(self, other):
        return BoltzmannResults(
            deltaF=other.deltaF + self.deltaF, Deltas=other.Deltas + self.Deltas,
            truncationError=other.truncationError + self.truncationError,
            truncatedTail=self.truncatedTail, spectralPeaks=self.spectralPeaks,
            linearizationCriterion1=other.linearizationCriterion1 + self.linearizationCriterion1,
            linearizationCriterion2=other.linearizationCriterion2 + self.linearizationCriterion2)

    def __sub__(self, other):
        return self.__add__((-1) * other)


class Particle:
    """Off-equilibrium species (reference particle.py:19-67); msqVacuum is a host callback."""

    def __init__(self, name, index, msqVacuum, msqDerivative, statistics, totalDOFs):
        if statistics not in ("Fermion", "Boson"):
            raise ValueError(f"{statistics=} not in ['Fermion', 'Boson']")
        self.name, self.index = name, index
        self.msqVacuum, self.msqDerivative = msqVacuum, msqDerivative
        self.statistics, self.totalDOFs = statistics, totalDOFs


class BoltzmannBackground:
    """Wall-frame background profiles on the M+1 Lobatto nodes (reference containers.py:112-164)."""

    def __init__(self, velocityMid, velocityProfile, fieldProfiles, temperatureProfile,
                 polynomialBasis="Cardinal"):
        self.velocityWall = 0
        self.velocityProfile = np.asarray(velocityProfile)
        self.fieldProfiles = fieldProfiles
        self.temperatureProfile = np.asarray(temperatureProfile)
        self.polynomialBasis = polynomialBasis
        self.velocityMid = velocityMid

    def boostToPlasmaFrame(self):
        vm = self.velocityMid
        self.velocityProfile = (self.velocityProfile - vm) / (1.0 - self.velocityProfile * vm)
        self.velocityWall = (self.velocityWall - vm) / (1.0 - self.velocityWall * vm)

    def boostToWallFrame(self):
        vw = self.velocityWall
        self.velocityProfile = (self.velocityProfile - vw) / (1.0 - self.velocityProfile * vw)
        self.velocityWall = 0


class CollisionArray:
    """In-memory collision tensor (P,N-1,N-1,P,N-1,N-1) with its momentum-polynomial basis
    (reference collisionArray.py:95-174, 356-388)."""

    def __init__(self, grid, basisType, particles, data):
        n1 = grid.N - 1
        data = np.asarray(data, dtype=np.float64)
        assert data.shape == (len(particles), n1, n1, len(particles), n1, n1)
        assert basisType in ("Cardinal", "Chebyshev")
        self.grid, self.basisType, self.particles = grid, basisType, particles
        self.size = n1
        self.data = data

    def __getitem__(self, key):
        return self.data[key]

    def getBasisSize(self):
        return self.size

    def getBasisType(self):
        return self.basisType

    # ------------------------------------------------------------------ file ingest
    @staticmethod
    def readPairFile(directoryPath, name1, name2):
        """(data (N_f-1)^4, basisSize N_f, basisType) of one ordered pair.  The reference's on-disk format is
        `collisions_{a}_{b}.hdf5` with attrs "Basis Size", "Basis Type" on the group "metadata" and a dataset
        "{a}, {b}" (collisionArray.py:225-249).  This image has no HDF5 library and every shipped .hdf5 is a
        Git-LFS pointer, so an HDF5 reader could not be verified; until then the same content is accepted as
          collisions_{a}_{b}.npz   keys: "data" (or "{a}, {b}"), "Basis Size", "Basis Type"
          collisions_{a}_{b}.npy   raw (N_f-1)^4 tensor, Chebyshev basis, N_f from the shape
        (`python -c "import h5py, numpy ..."` on a machine with h5py converts a directory in three lines,
        INTEGRATION.md).  An .hdf5 next to them is tried with h5py when that module exists."""
        import os

        base = os.path.join(str(directoryPath), f"collisions_{name1}_{name2}")
        if os.path.exists(base + ".npz"):
            with np.load(base + ".npz", allow_pickle=False) as z:
                key = "data" if "data" in z.files else f"{name1}, {name2}"
                if key not in z.files:
                    raise CollisionLoadError(f"CollisionArray error: {base}.npz holds no dataset 'data' or '{key}'.")
                data = np.array(z[key], dtype=np.float64)
                size = int(z["Basis Size"]) if "Basis Size" in z.files else data.shape[0] + 1
                btype = str(z["Basis Type"]) if "Basis Type" in z.files else "Chebyshev"
            return data, size, btype
        if os.path.exists(base + ".npy"):
            data = np.array(np.load(base + ".npy", allow_pickle=False), dtype=np.float64)
            return data, data.shape[0] + 1, "Chebyshev"
        if os.path.exists(base + ".hdf5"):
            try:
                import h5py  # pylint: disable=import-outside-toplevel
                with h5py.File(base + ".hdf5", "r") as f:
                    meta = f["metadata"]
                    btype = meta.attrs["Basis Type"]
                    btype = btype.decode() if isinstance(btype, bytes) else str(btype)
                    return np.array(f[f"{name1}, {name2}"][:], dtype=np.float64), int(meta.attrs["Basis Size"]), btype
            except (ImportError, OSError, AttributeError) as exc:
                raise CollisionLoadError(
                    f"CollisionArray error: {base}.hdf5 cannot be read here ({exc}); convert it to .npz "
                    "(keys 'data', 'Basis Size', 'Basis Type') -- see INTEGRATION.md.") from exc
        raise CollisionLoadError(f"CollisionArray error: {base}.hdf5 not found.")

    @staticmethod
    def newFromDirectory(directoryPath, grid, basisType, particles, bInterpolate=True, device=None):
        """Mirror of CollisionArray.newFromDirectory (collisionArray.py:179-354): one file per ordered pair of
        out-of-equilibrium species, all with the same basis size N_f and basis type; N_f > grid.N is brought
        down to the grid by `interpolateCollisionArray`, whose contraction runs ON THE DEVICE
        (wg_interpolate_collision) when a GPU is present.  Raises CollisionLoadError like the reference."""
        P = len(particles)
        tensor, sizeFile, typeFile = None, None, None
        for i, p1 in enumerate(particles):
            for j, p2 in enumerate(particles):
                data, size, btype = CollisionArray.readPairFile(directoryPath, p1.name, p2.name)
                if grid.N > size:
                    raise CollisionLoadError(
                        f"Target collision grid size ({grid.N}) must be smaller or equal to the grid size "
                        f"recorded in collision files ({size}).")
                if btype not in ("Cardinal", "Chebyshev"):
                    raise CollisionLoadError(f"CollisionArray error: unknown basis type {btype!r} in the collision file.")
                if data.shape != (size - 1,) * 4:
                    raise CollisionLoadError(
                        f"CollisionArray error: dataset '{p1.name}, {p2.name}' has shape {data.shape}, "
                        f"expected {(size - 1,) * 4}.")
                if tensor is None:
                    tensor = np.zeros((P, size - 1, size - 1, P, size - 1, size - 1))
                    sizeFile, typeFile = size, btype
                else:
                    assert size == sizeFile, "CollisionArray error: All the collision files must have the same basis size."
                    assert btype == typeFile, "CollisionArray error: All the collision files must have the same basis type."
                tensor[i, :, :, j, :, :] = data
        if sizeFile == grid.N:
            return CollisionArray(grid, typeFile, particles, tensor).changeBasis(basisType)
        if not bInterpolate:
            raise CollisionLoadError(
                f"Grid size mismatch when loading collision directory: {directoryPath}. "
                "Consider using bInterpolate=True in CollisionArray.loadFromFile().")
        from .grid import Grid  # pylint: disable=import-outside-toplevel
        dummyGrid = Grid(grid.M, sizeFile, getattr(grid, "positionFalloff", 1.0), grid.momentumFalloffT)
        src = CollisionArray(dummyGrid, typeFile, particles, tensor)
        return CollisionArray.interpolateCollisionArrayDevice(src, grid, device=device).changeBasis(basisType)

    @staticmethod
    def interpolationMatrices(srcGrid, targetGrid, nF):
        """Lagrange cardinals of the source momentum nodes evaluated on the target nodes: Lz, Lp [nT x nF]
        (what Polynomial.evaluate does along the Cardinal axes, polynomial.py:300-443)."""
        from .spectral import lobattoNodes

        _, rzF, rpF = lobattoNodes(srcGrid.M, srcGrid.N)
        _, rzT, rpT = lobattoNodes(targetGrid.M, targetGrid.N)

        def lagrange(xT, xFull, idx):
            out = np.ones((xT.size, idx.size))
            for n_, i in enumerate(idx):
                for k in range(xFull.size):
                    if k != i:
                        out[:, n_] *= (xT - xFull[k]) / (xFull[i] - xFull[k])
            return out

        Lz = lagrange(rzT, np.concatenate([[-1.0], rzF, [1.0]]), np.arange(1, nF + 1))   # [beta_T, beta_F]
        Lp = lagrange(rpT, np.concatenate([rpF, [1.0]]), np.arange(0, nF))               # [gamma_T, gamma_F]
        return Lz, Lp

    @staticmethod
    def interpolateCollisionArrayDevice(srcCollision, targetGrid, device=None):
        """interpolateCollisionArray with the contraction on the GPU (wg_interpolate_collision): what the file
        ingest uses.  No CPU path: raises without a CUDA device, like every other product call."""
        import copy

        nT = targetGrid.N - 1
        assert targetGrid.N <= srcCollision.getBasisSize(), (
            f"CollisionArray interpolation error: target grid size ({targetGrid.N}) must be smaller than or "
            f"equal to the source grid size ({srcCollision.getBasisSize() + 1}).")
        src = copy.deepcopy(srcCollision)
        src.changeBasis("Chebyshev")
        nF = src.size
        Lz, Lp = CollisionArray.interpolationMatrices(src.grid, targetGrid, nF)
        P = len(src.particles)
        from . import _lib  # pylint: disable=import-outside-toplevel
        from .engine import default_device  # pylint: disable=import-outside-toplevel

        lib = _lib.load()
        data = np.empty((P, nT, nT, P, nT, nT))
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (Lz, Lp, src.data)]
        _lib.check(lib.wg_interpolate_collision(default_device() if device is None else int(device), P, nF + 1, nT + 1,
                                                _lib.host_ptr(arrs[0]), _lib.host_ptr(arrs[1]), _lib.host_ptr(arrs[2]),
                                                _lib.host_ptr(data)), "wg_interpolate_collision")
        out = CollisionArray(targetGrid, "Chebyshev", src.particles, data)
        return out.changeBasis(srcCollision.getBasisType())

    @staticmethod
    def interpolateCollisionArray(srcCollision, targetGrid):
        """Collision tensor computed on a finer momentum grid (N_f) brought to `targetGrid` (N <= N_f - 1):
        the Chebyshev series in the polynomial axes is truncated and the collocation axes are evaluated at
        the target nodes with the Lagrange cardinals of the source grid (reference
        collisionArray.py:390-478, polynomial.py:300-443).  The result has the source's basis type.
        Note: like the reference, the evaluated array (points, a, b, j, k) is reshaped without a transpose,
        which is only an identity re-labelling for a single species."""
        import copy

        nT = targetGrid.N - 1
        assert targetGrid.N <= srcCollision.getBasisSize(), (
            f"CollisionArray interpolation error: target grid size ({targetGrid.N}) must be smaller than or "
            f"equal to the source grid size ({srcCollision.getBasisSize() + 1}).")
        src = copy.deepcopy(srcCollision)
        src.changeBasis("Chebyshev")
        nF = src.size
        Lz, Lp = CollisionArray.interpolationMatrices(src.grid, targetGrid, nF)
        P = len(src.particles)
        ev = np.einsum("xb,yg,abgcjk->xyacjk", Lz, Lp, src.data)                          # (bT, gT, a, b, j, k)
        ev = ev.reshape(nT * nT, P, P, nF, nF)[..., :nT, :nT]
        data = np.ascontiguousarray(ev).reshape(P, nT, nT, P, nT, nT)
        out = CollisionArray(targetGrid, "Chebyshev", src.particles, data)
        return out.changeBasis(srcCollision.getBasisType())

    def changeBasis(self, newBasisType):
        """Inverse-transpose transform of the two polynomial axes (polynomial.py:282-297)."""
        from .spectral import chebyshevMatrix, lobattoNodes

        if newBasisType == self.basisType:
            return self
        assert newBasisType in ("Cardinal", "Chebyshev"), f"collisionarray error: unknown basis {newBasisType}"
        _, rz, rp = lobattoNodes(self.grid.M, self.grid.N)
        out = self.data
        for axis, T in ((4, chebyshevMatrix(rz, self.size, "full")), (5, chebyshevMatrix(rp, self.size, "partial"))):
            mat = T.T if newBasisType == "Chebyshev" else np.linalg.inv(T).T
            out = np.moveaxis(np.tensordot(mat, out, axes=([1], [axis])), 0, axis)
        self.data = np.ascontiguousarray(out)
        self.basisType = newBasisType
        return self
