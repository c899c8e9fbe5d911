"""Deterministic synthetic inputs of the Boltzmann path: the reference's own
benchmark grid, a tanh wall background and a relaxation-time-like collision tensor.  Used by
bench.py, smoke() and the scan example so that every arm is fed the same bytes."""
from __future__ import annotations

import numpy as np

from .containers import BoltzmannBackground, CollisionArray, Particle
from .grid import Grid3Scales


class _Fields:
    """Single-field stand-in for WallGo.Fields: phi on the M+1 profile points."""

    def __init__(self, phi):
        self.phi = np.asarray(phi)

    def getField(self, i):
        assert i == 0
        return self.phi


def makeParticles(P: int):
    return [Particle(f"p{a}", a, (lambda f, a=a: 0.5 * (1 + 0.3 * a) * f.getField(0) ** 2),
                     (lambda f, a=a: (1 + 0.3 * a) * f.getField(0)),
                     "Fermion" if a % 2 == 0 else "Boson", 12) for a in range(P)]


def makeGrid(M: int, N: int):
    return Grid3Scales(M, N, 0.2, 0.2, 0.05, 100.0, 0.5, 0.1)


def makeBackground(grid, velocityMid: float = -0.5, wallL: float = 0.05, dv: float = 0.02):
    xi = grid.getCoordinates(endpoints=True)[0]
    th = np.tanh(xi / wallL)
    return BoltzmannBackground(velocityMid, velocityMid + dv * th, _Fields(50 * (1 - th)), 100 + th)


def makeCollision(grid, particles, r: float = 0.5, noise: float = 1e-3, seed: int = 0, basis="Chebyshev"):
    P, n1 = len(particles), grid.N - 1
    q = n1 * n1
    rng = np.random.default_rng(seed)
    C = (r * np.identity(P * q) + noise * rng.standard_normal((P * q, P * q))).reshape(P, n1, n1, P, n1, n1)
    return CollisionArray(grid, "Cardinal", particles, C).changeBasis(basis)
