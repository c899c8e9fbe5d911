"""Shared fixtures.  `-m gpu` tests need a B200 and call the CUDA path through the C ABI;
everything else runs on CPU.  Nothing here reads /root/reference (it does not exist on the GPU box)."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                  if not n.startswith(("trace_", "interp_")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


class FixtureGrid:
    """Grid object carrying exactly the reference's arrays from a golden file (duck-typed Grid)."""

    def __init__(self, g):
        self.M, self.N = int(g["M"]), int(g["N"])
        self.chiValues, self.rzValues, self.rpValues = g["chi"], g["rz"], g["rp"]
        self.pzValues, self.ppValues = g["pz"], g["pp"]
        self.dxidchi, self.dpzdrz, self.dppdrp = g["dxidchi"], g["dpzdrz"], g["dppdrp"]
        self.xiValues = np.zeros(self.M - 1)
        self.momentumFalloffT = 100.0

    def getCompactCoordinates(self, endpoints=False, direction=None):
        assert not endpoints
        return self.chiValues, self.rzValues, self.rpValues

    def getCoordinates(self, endpoints=False):
        return self.xiValues, self.pzValues, self.ppValues

    def getCompactificationDerivatives(self, endpoints=False):
        return self.dxidchi, self.dpzdrz, self.dppdrp


class FixtureFields:
    """Stand-in for WallGo.Fields: msq is taken from the golden file, not recomputed."""

    def __init__(self, msq, dmsq_interior=None):
        self.msq = msq
        self.dmsq = None
        if dmsq_interior is not None:   # [P, M-1, F] -> padded to the M+1 profile points
            P, m1, F = dmsq_interior.shape
            self.dmsq = np.zeros((P, m1 + 2, F))
            self.dmsq[:, 1:-1] = dmsq_interior


class FixtureBackground:
    """Already boosted background (the golden files store the plasma-frame profiles)."""

    def __init__(self, g):
        self.velocityWall = float(g["vw"])
        self.velocityProfile = g["v"]
        self.temperatureProfile = g["T"]
        self.fieldProfiles = FixtureFields(g["msq"], g.get("dmsq_interior"))
        self.velocityMid = float(g["velocityMid"]) if "velocityMid" in g else 0.0

    def boostToPlasmaFrame(self):
        pass


def solver_from_golden(g, device=0):
    """wallgo_b200.BoltzmannSolver fed with the reference's input bytes."""
    import wallgo_b200 as wb

    grid = FixtureGrid(g)
    P = int(g["P"])
    particles = [wb.Particle(f"p{a}", a, (lambda f, a=a: f.msq[a]),
                             ((lambda f, a=a: f.dmsq[a]) if "dmsq_interior" in g else None),
                             "Fermion" if g["statistics"][a] < 0 else "Boson", int(g["dofs"][a]))
                 for a in range(P)]
    solver = wb.BoltzmannSolver(grid, str(g["basisM"]), str(g["basisN"]), str(g["derivatives"]),
                                float(g["collisionMultiplier"]),
                                getattr(wb.ETruncationOption, str(g["option"])), device=device)
    solver.updateParticleList(particles)
    solver.setBackground(FixtureBackground(g))
    solver.setCollisionArray(wb.CollisionArray(grid, str(g["basisN"]), particles, g["collision"]))
    return solver


def oracle_problem_from_golden(g):
    from oracle import boltzmann_oracle as bo

    tab = bo.build_tables(int(g["M"]), int(g["N"]), str(g["basisM"]), str(g["basisN"]), str(g["derivatives"]))
    return bo.Problem(tab=tab, xi_dxidchi=g["dxidchi"], pz=g["pz"], pp=g["pp"], dpzdrz=g["dpzdrz"],
                      dppdrp=g["dppdrp"], T=g["T"], v=g["v"], msq=g["msq"], vw=float(g["vw"]),
                      statistics=g["statistics"], collision=g["collision"],
                      collisionMultiplier=float(g["collisionMultiplier"]), dofs=g["dofs"])


def relmax(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return 0
