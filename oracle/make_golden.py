"""
Pins the oracle against the LIVE reference and writes golden fixtures.

Runs ONLY in the build container (needs /root/reference).  It imports the
unmodified reference (`WallGo` from /root/reference/src, with an empty `h5py`
stub because every shipped .hdf5 is a Git-LFS pointer and h5py is not
installed), drives `WallGo.BoltzmannSolver` on deterministic synthetic inputs
 and

  1. asserts that `oracle/boltzmann_oracle.py` reproduces the reference's
     tables, operator, source, deltaF, truncation decisions and moments;
  2. stores inputs + reference outputs as small .npz files under tests/golden/
     so that GPU-box tests (no reference there) compare against reference bytes.

usage:  python oracle/make_golden.py            (from the repo root)
"""
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
from WallGo.boltzmann import ETruncationOption  # noqa: E402

from oracle import boltzmann_oracle as bo  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def make_reference_solver(P, M, N, basisM, basisN, derivatives, option, r, seed, mult, grid3=True):
    if grid3:
        grid = WallGo.Grid3Scales(M, N, 0.2, 0.2, 0.05, 100.0, 0.5, 0.1)
    else:
        grid = WallGo.Grid(M, N, 0.05, 100.0)
    L = 0.05
    xi = grid.getCoordinates(endpoints=True)[0]
    th = np.tanh(xi / L)
    phi = 50 * (1 - th)
    T = 100 + th
    v = -0.5 + 0.02 * th
    particles = []
    for a in range(P):
        particles.append(WallGo.Particle(
            name=f"p{a}", index=a,
            msqVacuum=(lambda f, a=a: 0.5 * (1 + 0.3 * a) * f.getField(0) ** 2),
            msqDerivative=(lambda f, a=a: np.transpose([(1 + 0.3 * a) * f.getField(0)])),
            statistics="Fermion" if a % 2 == 0 else "Boson", totalDOFs=12))
    bg = WallGo.BoltzmannBackground(-0.5, v, WallGo.Fields(phi[:, None]), T)
    solver = WallGo.BoltzmannSolver(grid, basisM, basisN, derivatives, mult,
                                    getattr(ETruncationOption, option))
    solver.updateParticleList(particles)
    solver.setBackground(bg)
    rng = np.random.default_rng(seed)
    q = (N - 1) ** 2
    C = (r * np.identity(P * q) + 1e-3 * rng.standard_normal((P * q, P * q))).reshape(
        P, N - 1, N - 1, P, N - 1, N - 1)
    poly = Polynomial(C.copy(), grid,
                      ("Array", "Cardinal", "Cardinal", "Array", "Cardinal", "Cardinal"),
                      CollisionArray.AXIS_TYPES, endpoints=False)
    ca = CollisionArray.newFromPolynomial(poly, particles).changeBasis(basisN)
    solver.setCollisionArray(ca)
    return solver, grid, bg, particles


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def one_case(name, P, M, N, basisM="Cardinal", basisN="Chebyshev", derivatives="Spectral",
             option="AUTO", r=0.5, seed=0, mult=1.0, grid3=True, store_operator=False,
             linearization=False):
    solver, grid, bg, particles = make_reference_solver(P, M, N, basisM, basisN, derivatives,
                                                        option, r, seed, mult, grid3)
    op, src, _, _ = solver.buildLinearEquations()
    dF = solver.solveBoltzmannEquations()
    res = solver.getDeltas(dF.copy())
    # --- the same thing through the oracle ---
    pb = bo.synthetic_problem(P, M, N, r=r, seed=seed, collisionMultiplier=mult, basisM=basisM,
                              basisN=basisN, derivatives=derivatives, grid3=grid3)
    # identical input bytes for the collision tensor
    checks = {}
    checks["collision"] = rel(pb.collision, solver.collisionArray[...])
    pb.collision = np.array(solver.collisionArray[...])
    poly = Polynomial(np.zeros(M + 1), grid, "Cardinal", "z", True)
    checks["Dchi"] = rel(pb.tab.Dchi, poly.derivMatrix(basisM, "z")[1:-1]) if derivatives == "Spectral" else 0.0
    checks["Trz"] = rel(pb.tab.Trz, poly.matrix(basisN, "pz")) if derivatives == "Spectral" else 0.0
    checks["Trp"] = rel(pb.tab.Trp, poly.matrix(basisN, "pp")) if derivatives == "Spectral" else 0.0
    checks["Drz"] = rel(pb.tab.Drz, poly.derivMatrix(basisN, "pz")[1:-1]) if derivatives == "Spectral" else 0.0
    checks["dxidchi"] = rel(pb.xi_dxidchi, grid.getCompactificationDerivatives()[0])
    checks["T"] = rel(pb.T, solver.background.temperatureProfile)
    checks["v"] = rel(pb.v, solver.background.velocityProfile)
    checks["vw"] = abs(pb.vw - solver.background.velocityWall)
    oop, osrc = bo.assemble(pb)
    checks["operator"] = rel(oop, op)
    checks["source"] = rel(osrc, src)
    odF = np.linalg.solve(oop, osrc).reshape(dF.shape)
    checks["deltaF"] = rel(odF, dF)
    ores = bo.get_deltas(pb, dF.copy(), option)
    refD = np.stack([res.Deltas.Delta00.coefficients, res.Deltas.Delta02.coefficients,
                     res.Deltas.Delta20.coefficients, res.Deltas.Delta11.coefficients])
    checks["deltaF_trunc"] = rel(ores["deltaF"], res.deltaF)
    for i, nm in enumerate(("D00", "D02", "D20", "D11")):
        checks[nm] = rel(ores["Deltas"][i], refD[i])
    checks["truncErr"] = abs(ores["truncationError"] - res.truncationError) / max(res.truncationError, 1e-300)
    assert tuple(ores["truncatedTail"]) == tuple(res.truncatedTail), (ores["truncatedTail"], res.truncatedTail)
    assert tuple(ores["spectralPeaks"]) == tuple(res.spectralPeaks), (ores["spectralPeaks"], res.spectralPeaks)
    # --- moment feedback into the wall EOM: the reference's own EOM.deltaToTmunu, called unbound (it only
    #     reads self.particles), plus the two inline source terms of equationOfMotion.py:1418-1429, 1545-1555
    from WallGo.equationOfMotion import EOM
    stub = types.SimpleNamespace(particles=particles)
    fld = solver.background.fieldProfiles
    vmid = float(solver.background.velocityMid)
    tm = np.array([EOM.deltaToTmunu(stub, i, fld.getFieldPoint(i + 1), vmid, res.Deltas) for i in range(M - 1)])
    fint = fld.takeSlice(1, -1, axis=fld.overFieldPoints)
    potOut = sum(p.totalDOFs * p.msqVacuum(fint) * res.Deltas.Delta00.coefficients[i]
                 for i, p in enumerate(particles)) / 2
    dVout = np.sum([p.totalDOFs * p.msqDerivative(fint) * res.Deltas.Delta00.coefficients[i, :, None]
                    for i, p in enumerate(particles)], axis=0) / 2
    dmsqInt = np.stack([np.asarray(p.msqDerivative(fint)) for p in particles])
    msqInt = np.array([p.msqVacuum(fint) for p in particles])
    oT30, oT33, oPot, odV = bo.feedback(refD, msqInt, [p.totalDOFs for p in particles], vmid, dmsqInt)
    checks["T30"], checks["T33"] = rel(oT30, tm[:, 0]), rel(oT33, tm[:, 1])
    checks["potOut"], checks["dVout"] = rel(oPot, potOut), rel(odV, dVout)
    lin = None
    if linearization:
        lin = solver.checkLinearization(dF.copy())
        olin = bo.check_linearization(pb, dF.copy())
        checks["lin1"] = abs(olin[0] - lin[0]) / abs(lin[0])
        checks["lin2"] = abs(olin[1] - lin[1]) / abs(lin[1])
    worst = max(checks.values())
    print(f"{name:28s} n={op.shape[0]:6d} worst={worst:.2e}  " +
          " ".join(f"{k}={v:.1e}" for k, v in checks.items() if v > 1e-13))
    tol = {"deltaF": 1e-10, "deltaF_trunc": 1e-10, "D00": 1e-10, "D02": 1e-10, "D20": 1e-10,
           "D11": 1e-10, "truncErr": 1e-9, "lin1": 1e-9, "lin2": 1e-8, "source": 1e-11}
    for k, v in checks.items():
        assert v <= tol.get(k, 1e-12), f"{name}: oracle differs from the reference in {k}: {v:.3e}"
    ones = np.ones(op.shape[0])
    out = dict(
        P=P, M=M, N=N, basisM=basisM, basisN=basisN, derivatives=derivatives, option=option,
        collisionMultiplier=mult, grid3=grid3,
        chi=grid.chiValues, rz=grid.rzValues, rp=grid.rpValues,
        dxidchi=grid.getCompactificationDerivatives()[0], pz=grid.pzValues, pp=grid.ppValues,
        dpzdrz=grid.dpzdrz, dppdrp=grid.dppdrp,
        T=solver.background.temperatureProfile, v=solver.background.velocityProfile,
        vw=solver.background.velocityWall,
        msq=np.array([p.msqVacuum(solver.background.fieldProfiles) for p in particles]),
        statistics=np.array([-1.0 if p.statistics == "Fermion" else 1.0 for p in particles]),
        dofs=np.array([float(p.totalDOFs) for p in particles]),
        collision=np.array(solver.collisionArray[...]),
        ref_Tchi=pb.tab.Tchi if derivatives != "Spectral" else poly.matrix(basisM, "z"),
        ref_Dchi=pb.tab.Dchi if derivatives != "Spectral" else poly.derivMatrix(basisM, "z")[1:-1],
        ref_Trz=pb.tab.Trz if derivatives != "Spectral" else poly.matrix(basisN, "pz"),
        ref_Trp=pb.tab.Trp if derivatives != "Spectral" else poly.matrix(basisN, "pp"),
        ref_Drz=pb.tab.Drz if derivatives != "Spectral" else poly.derivMatrix(basisN, "pz")[1:-1],
        ref_source=src, ref_deltaF=dF, ref_deltaF_trunc=res.deltaF, ref_Deltas=refD,
        ref_truncationError=res.truncationError,
        ref_truncatedTail=np.array(res.truncatedTail), ref_spectralPeaks=np.array(res.spectralPeaks),
        velocityMid=vmid, dmsq_interior=dmsqInt, ref_T30=tm[:, 0], ref_T33=tm[:, 1], ref_potOut=np.asarray(potOut),
        ref_dVout=np.asarray(dVout),
        ref_op_rowsum=op @ ones, ref_op_colsum=op.T @ ones,
        ref_op_diag=np.diag(op).copy(),
        ref_op_probe_rows=np.array([0, op.shape[0] // 3, op.shape[0] - 1]),
        ref_op_probe=op[[0, op.shape[0] // 3, op.shape[0] - 1], :].copy(),
    )
    if store_operator:
        out["ref_operator"] = op
    if lin is not None:
        out["ref_linearization"] = np.array(lin)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    return checks


def known_answer_delta00():
    """Reference known-answer test tests/test_Boltzmann.py:14-85 through the oracle."""
    for (M, N, a, b, c, d, e, f) in [(25, 19, 1, 2, 3, 4, 5, 6), (5, 5, 1, 1, 2, 3, 5, 8)]:
        tab = bo.build_tables(M, N, "Cardinal", "Cardinal", "Spectral")
        pz, pp, dpzdrz, dppdrp = bo.momentum_maps(tab.rz, tab.rp, 100.0)
        fieldv = np.ones(M + 1)
        fieldv[M // 2:] = 0
        fieldv += 0.1 * np.sin(7 * 2 * np.pi * np.arange(M + 1) + 6)
        msq = 0.5 * fieldv**2
        T = 100 * np.ones(M + 1)
        rz = tab.rz[None, :, None]
        rp = tab.rp[None, None, :]
        E = np.sqrt(msq[1:-1, None, None] + pz[None, :, None] ** 2 + pp[None, None, :] ** 2)
        eps = 2e-16
        integrand = (2 * E * (1 - rz**2) * (1 - rp**2)
                     * np.sqrt((1 - rz**2) * (1 - rp) ** 2 / (1 - rp**2 + eps))
                     / (np.log(2 / (1 - rp)) + eps))
        integrand = integrand * (a + b * rz + c * rz**2) * (d + e * rp + f * rp**2)
        pb = bo.Problem(tab=tab, xi_dxidchi=np.ones(M - 1), pz=pz, pp=pp, dpzdrz=dpzdrz,
                        dppdrp=dppdrp, T=T, v=np.zeros(M + 1), msq=msq[None, :], vw=0.0,
                        statistics=np.array([-1.0]), collision=np.zeros((1, N - 1, N - 1, 1, N - 1, N - 1)))
        D = bo.moments(integrand[None, ...], pb)
        exact = (4 * a + c) * (4 * d + f) * T[1:-1] ** 3 / 64
        err = np.max(np.abs(D[0, 0] / exact - 1))
        print(f"known-answer Delta00 M={M} N={N}: rel err {err:.2e}")
        assert err < 1e-14


def interpolation_case(name, P, M, Nf, Nt, seed):
    """CollisionArray.interpolateCollisionArray of the live reference (collisionArray.py:390-478)."""
    gridF = WallGo.Grid(M, Nf, 0.05, 100.0)
    gridT = WallGo.Grid(M, Nt, 0.05, 100.0)
    parts = [WallGo.Particle(f"p{a}", a, lambda f: f.getField(0), lambda f: f.getField(0), "Fermion", 1)
             for a in range(P)]
    rng = np.random.default_rng(seed)
    C = rng.standard_normal((P, Nf - 1, Nf - 1, P, Nf - 1, Nf - 1))
    poly = Polynomial(C.copy(), gridF, ("Array", "Cardinal", "Cardinal", "Array", "Chebyshev", "Chebyshev"),
                      CollisionArray.AXIS_TYPES, endpoints=False)
    src = CollisionArray.newFromPolynomial(poly, parts)
    out = CollisionArray.interpolateCollisionArray(src, gridT)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), P=P, M=M, Nf=Nf, Nt=Nt, source=C,
                        ref_interpolated=np.array(out[...]), basis=out.getBasisType())
    print(f"{name:28s} interpolated {C.shape} -> {out[...].shape}")


if __name__ == "__main__":
    interpolation_case("interp_P1_N9_to_N5", 1, 6, 9, 5, 0)
    interpolation_case("interp_P2_N7_to_N5", 2, 6, 7, 5, 1)
    known_answer_delta00()
    one_case("tiny_P1_M6_N5", 1, 6, 5, store_operator=True, linearization=True)
    one_case("small_P2_M12_N7", 2, 12, 7, linearization=True)
    one_case("small_P2_M12_N7_rta", 2, 12, 7, r=0.05, seed=1)
    one_case("cfg1_P1_M20_N11", 1, 20, 11)
    one_case("cfg1_P1_M20_N11_mult1e-2", 1, 20, 11, mult=1e-2, seed=3)
    one_case("third_P1_M10_N7", 1, 10, 7, option="THIRD")
    one_case("none_P1_M10_N7", 1, 10, 7, option="NONE")
    one_case("cardcard_P1_M8_N7", 1, 8, 7, basisN="Cardinal", option="AUTO")
    one_case("chebcheb_P1_M8_N7", 1, 8, 7, basisM="Chebyshev", option="AUTO")
    one_case("fd_P1_M10_N7", 1, 10, 7, basisN="Cardinal", derivatives="Finite Difference")
    one_case("simplegrid_P3_M9_N5", 3, 9, 5, grid3=False)
    print("oracle pinned against the live reference; fixtures written to", GOLD)
