"""Host-side grid: Gauss-Lobatto nodes and compactification maps.

Mirror of the reference's `Grid` / `Grid3Scales` interface (grid.py:37-378,
grid3Scales.py:45-307) limited to what the Boltzmann path reads: `M`, `N`, the node arrays,
physical coordinates and Jacobians.  These are O(M+N) numbers recomputed once per
`wallPressure`, so they stay on the host; the solver uploads them with the background.
Any object with these attributes (in particular WallGo's own grids) can be used instead.
"""
from __future__ import annotations

import numpy as np

from .spectral import lobattoNodes


class Grid:
    def __init__(self, M: int, N: int, positionFalloff: float, momentumFalloffT: float):
        assert N % 2 == 1, "N must be odd"
        self.M, self.N = int(M), int(N)
        self.positionFalloff = positionFalloff
        self.momentumFalloffT = momentumFalloffT
        self.chiValues, self.rzValues, self.rpValues = lobattoNodes(self.M, self.N)
        self._cacheCoordinates()

    # --- maps (z part overridden by Grid3Scales)
    def _positionMap(self, chi):
        L = self.positionFalloff
        return L * chi / np.sqrt(1 - chi**2), L / (1 - chi**2) ** 1.5

    def _cacheCoordinates(self):
        T0 = self.momentumFalloffT
        self.xiValues, self.dxidchi = self._positionMap(self.chiValues)
        self.pzValues = 2 * T0 * np.arctanh(self.rzValues)
        self.ppValues = -T0 * np.log((1 - self.rpValues) / 2)
        self.dpzdrz = 2 * T0 / (1 - self.rzValues**2)
        self.dppdrp = T0 / (1 - self.rpValues)

    def changeMomentumFalloffScale(self, newScale):
        self.momentumFalloffT = newScale
        self._cacheCoordinates()

    def changePositionFalloffScale(self, newScale):
        self.positionFalloff = newScale
        self._cacheCoordinates()

    # --- accessors with the reference's names
    def getCompactCoordinates(self, endpoints=False, direction=None):
        chi, rz, rp = self.chiValues, self.rzValues, self.rpValues
        if endpoints:
            chi = np.concatenate([[-1.0], chi, [1.0]])
            rz = np.concatenate([[-1.0], rz, [1.0]])
            rp = np.concatenate([rp, [1.0]])
        return {"z": chi, "pz": rz, "pp": rp}.get(direction, (chi, rz, rp))

    def getCoordinates(self, endpoints=False):
        if endpoints:
            inf = np.inf
            return (np.concatenate([[-inf], self.xiValues, [inf]]),
                    np.concatenate([[-inf], self.pzValues, [inf]]),
                    np.concatenate([self.ppValues, [inf]]))
        return self.xiValues, self.pzValues, self.ppValues

    def getCompactificationDerivatives(self, endpoints=False):
        if endpoints:
            inf = np.inf
            return (np.concatenate([[inf], self.dxidchi, [inf]]),
                    np.concatenate([[inf], self.dpzdrz, [inf]]),
                    np.concatenate([self.dppdrp, [inf]]))
        return self.dxidchi, self.dpzdrz, self.dppdrp


class Grid3Scales(Grid):
    """z(chi) with separate tail lengths inside/outside and a linear core (grid3Scales.py)."""

    def __init__(self, M, N, tailLengthInside, tailLengthOutside, wallThickness, momentumFalloffT,
                 ratioPointsWall=0.5, smoothing=0.1, wallCenter=0.0):
        self._updateParameters(tailLengthInside, tailLengthOutside, wallThickness, ratioPointsWall,
                               smoothing, wallCenter)
        super().__init__(M, N, wallThickness, momentumFalloffT)

    def changePositionFalloffScale(self, tailLengthInside, tailLengthOutside, wallThickness, wallCenter):
        self._updateParameters(tailLengthInside, tailLengthOutside, wallThickness, self.ratioPointsWall,
                               self.smoothing, wallCenter)
        self._cacheCoordinates()

    def _updateParameters(self, tailIn, tailOut, L, r, smoothing, center):
        assert L > 0 and smoothing > 0 and 0 < r < 1
        assert tailIn > L * (0.5 + smoothing) / r and tailOut > L * (0.5 + smoothing) / r
        self.tailLengthInside, self.tailLengthOutside = tailIn, tailOut
        self.wallThickness, self.ratioPointsWall = L, r
        self.smoothing, self.wallCenter = smoothing, center

        def width(tail):
            return (np.sqrt(4 * smoothing * L * r**2 * (2 * r * tail - L * (1 + smoothing)))
                    / abs(2 * r * tail - L * (1 + 2 * smoothing)))

        self.aIn, self.aOut = width(tailIn), width(tailOut)

    def _positionMap(self, chi):
        L, r = self.wallThickness, self.ratioPointsWall
        tIn, tOut, aIn, aOut = self.tailLengthInside, self.tailLengthOutside, self.aIn, self.aOut

        def atanhReal(y):
            return np.arctanh(np.asarray(y) + 0j).real

        def mapping(x):
            x = np.asarray(x, dtype=float)
            rootOut = np.sqrt(aOut**2 + (x - r) ** 2)
            rootIn = np.sqrt(aIn**2 + (x + r) ** 2)
            dOutM, dOutP = np.sqrt(aOut**2 + (1 - r) ** 2), np.sqrt(aOut**2 + (1 + r) ** 2)
            dInM, dInP = np.sqrt(aIn**2 + (1 - r) ** 2), np.sqrt(aIn**2 + (1 + r) ** 2)
            cOut, cIn = (2 * r * tOut - L) / r, (2 * r * tIn - L) / r
            total = (1 - r) * cOut * atanhReal((1 - x + rootOut) / dOutM) / dOutM
            total = total - (1 + r) * cOut * atanhReal((1 + x - rootOut) / dOutP) / dOutP
            total = total + (1 - r) * cIn * atanhReal((1 + x - rootIn) / dInM) / dInM
            total = total - (1 + r) * cIn * atanhReal((1 - x + rootIn) / dInP) / dInP
            total = total + (2 * tIn + 2 * tOut - 4 * self.smoothing * L / r) * np.arctanh(x)
            return total / 2

        xi = mapping(chi) - mapping(0.0) + self.wallCenter
        d = (2 * tIn - L / r) * (1 - (chi + r) / np.sqrt(aIn**2 + (chi + r) ** 2)) / 2
        d = d + (2 * tOut - L / r) * (1 + (chi - r) / np.sqrt(aOut**2 + (chi - r) ** 2)) / 2
        d = d + (1 - 2 * self.smoothing) * L / r
        return xi, d / (1 - chi**2)
