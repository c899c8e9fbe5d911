"""Direct parity at the BASELINE.json configs (full size) against the reference ITSELF, side by side.

For each config the synthetic system of SURVEY.md section 8(d) is built from the reference's own classes
(Grid3Scales, Particle, BoltzmannBackground, CollisionArray) and handed, the same objects, to
  * `WallGo.BoltzmannSolver` of the pure-Python reference copy under oracle/_ref (NumPy + LAPACK dgesv on the
    host), boltzmann.py:141-294, 628-820, and
  * the drop-in subclass backed by the CUDA engine (through the C ABI).
Gates: operator rows <= 2e-13 of the row maximum, source <= 1e-11 of its maximum (the device basis tables agree with
scipy.special to 2e-13), deltaF and each Delta <= 1e-10 relative (north star), truncation flags, truncated
shape and spectral peaks EQUAL.
"""
import os

import numpy as np
import pytest

from conftest import relmax

pytestmark = pytest.mark.gpu

CONFIGS = {"cfg1": (1, 20, 11), "cfg3": (3, 30, 11), "cfg4": (4, 30, 11), "cfg2a": (1, 40, 21), "cfg2b": (2, 40, 21)}


@pytest.fixture(scope="module")
def rd():
    from oracle import ref_drivers
    if ref_drivers.reference_root() is None:
        pytest.skip("reference copy not present (oracle/fetch_ref.sh fills oracle/_ref in the build container)")
    return ref_drivers


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_baseline_config_against_reference_class(cuda_device, rd, cfg):
    import psutil
    from wallgo_b200 import dropin

    P, M, N = CONFIGS[cfg]
    n = P * (M - 1) * (N - 1) ** 2
    need = 6 * 8 * n * n      # the reference materialises three rank-8 arrays plus temporaries; the test holds two
    if psutil.virtual_memory().available < need + (8 << 30):
        pytest.skip(f"{cfg}: the reference needs ~{need / 2**30:.0f} GiB of host RAM")
    ref, grid, bg, particles = rd.synthetic_solver(P, M, N)
    new = dropin.makeDropInClass()(grid, ref.basisM, ref.basisN, ref.derivatives, ref.collisionMultiplier,
                                   ref.truncationOption)
    new.updateParticleList(particles)
    new.setBackground(bg)
    new.setCollisionArray(ref.collisionArray)

    # ---- operator and source, row by row
    opRef, srcRef, _, _ = ref.buildLinearEquations()
    eng = new._getEngine()
    new._uploadBackground(eng)
    eng.assemble(float(new.background.velocityWall), float(new.collisionMultiplier))
    srcNew = eng.S.cpu().numpy()
    assert relmax(srcNew, srcRef) <= 1e-11, cfg     # exp() of the equilibrium distribution: libdevice vs libm
    worst = 0.0
    step = max(1, (1 << 27) // n)       # rows per device->host slice (1 GiB)
    for r0 in range(0, n, step):
        blk = eng.A[r0:r0 + step].cpu().numpy()
        refBlk = opRef[r0:r0 + step]
        scale = np.max(np.abs(refBlk), axis=1, keepdims=True)
        worst = max(worst, float(np.max(np.abs(blk - refBlk) / scale)))
    assert worst <= 2e-13, (cfg, worst)

    # ---- deltaF = np.linalg.solve(operator, source) on the host vs LU on the device
    dRef = np.linalg.solve(opRef, srcRef).reshape(P, M - 1, N - 1, N - 1)
    del opRef
    dNew = new.solveBoltzmannEquations()
    errF = relmax(dNew, dRef)
    assert errF <= 1e-10, (cfg, errF)

    # ---- truncation + moments from each side's own solution (the production sequence)
    rRef = ref.getDeltas(dRef.copy())
    rNew = new.getDeltas()
    assert tuple(rNew.truncatedTail) == tuple(rRef.truncatedTail)
    assert tuple(rNew.spectralPeaks) == tuple(rRef.spectralPeaks)
    assert rNew.deltaF.shape == rRef.deltaF.shape
    assert relmax(rNew.deltaF, rRef.deltaF) <= 1e-10
    DRef, DNew = rd.deltas_table(rRef), rd.deltas_table(rNew)
    errD = max(relmax(DNew[i], DRef[i]) for i in range(4))
    assert errD <= 1e-10, (cfg, errD)
    assert abs(rNew.truncationError - rRef.truncationError) <= 1e-8 * rRef.truncationError
    shpRef = ref.checkSpectralConvergence(dRef.copy())[1]
    shpNew = new.checkSpectralConvergence(dRef.copy())[1]
    assert tuple(shpNew) == tuple(shpRef)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "refparity.log"), "a") as f:
        f.write(f"{cfg} P={P} M={M} N={N} n={n}: operator rows {worst:.2e}, source {relmax(srcNew, srcRef):.2e}, "
                f"deltaF {errF:.2e}, Deltas {errD:.2e}, tail {tuple(rNew.truncatedTail)}, peaks {tuple(rNew.spectralPeaks)}\n")
