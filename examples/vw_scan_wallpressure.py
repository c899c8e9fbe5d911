"""Wall-velocity scan as BASELINE.json's metric defines it: every point is one unmodified
`EOM.wallPressure(vw_i, wallParams0, None)` of WallGo (equationOfMotion.py:786-1131), the Boltzmann solves
inside it served by the B200 (wallgo_b200.dropin), over 1..8 GPUs.

    python examples/vw_scan_wallpressure.py --wallgo /path/to/WallGo --model InertDoubletModel/inertDoubletModel.py
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/vw_scan_wallpressure.py ...

Needs an importable WallGo (pure Python) and one of its example model files; `makeManager(module)` below must return
a `WallGoManager` after `setupThermodynamicsHydrodynamics` plus the `WallSolverSettings` to use -- adapt it to your
model (the bench and the tests use oracle/ref_drivers.py for the four shipped examples).

Structure (wallgo_b200/vwscan.py): each rank builds its own manager, forks its host workers BEFORE it touches CUDA,
then runs a solve server on its GPU; workers execute the reference's pressure iteration and send every getDeltas()
to the server; one all_gather returns the table to every rank.  By default the server re-uses the LU factors across
the solves of a point (third solve on: refinement steps on the last factors instead of a new LU; --no-reuse for the
plain path, bit-identical to the sequential drop-in).
"""
import argparse
import importlib.util
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def makeManager(module):
    """Adapt to your model: the shipped examples expose a class derived from WallGoExampleBase; oracle/ref_drivers.py
    shows the few lines that configure it without the command-line front end."""
    raise SystemExit("edit makeManager() for your model (see oracle/ref_drivers.py::build_manager)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--wallgo", required=True, help="directory holding src/WallGo and Models/")
    ap.add_argument("--model", required=True, help="model file below Models/")
    ap.add_argument("--points", type=int, default=256)
    ap.add_argument("--vmax", type=float, default=0.53)
    ap.add_argument("--no-reuse", action="store_true", help="factor every system afresh")
    a = ap.parse_args()
    sys.path.insert(0, os.path.join(a.wallgo, "src"))
    sys.path.insert(0, os.path.join(a.wallgo, "Models"))

    import WallGo  # noqa: F401  pylint: disable=unused-import
    from wallgo_b200 import dropin, vwscan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dropin.install(vectoriseEOM=True)
    spec = importlib.util.spec_from_file_location("scan_model", os.path.join(a.wallgo, "Models", a.model))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    manager, settings = makeManager(module)
    vws = np.linspace(max(float(manager.hydrodynamics.vMin), 0.01), min(a.vmax, 0.995 * manager.hydrodynamics.vJ), a.points)
    workers = vwscan.HostWorkers(manager, settings, vwscan.defaultWorkers(world))    # before CUDA is initialised

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stats = {}
    table = vwscan.scanWallPressure(manager, settings, vws, stats=stats, hostWorkers=workers,
                                    reuseFactors=not a.no_reuse)
    if rank == 0:
        print("factor re-use on rank 0:", stats.get("reuse"))
        ws = manager.setupWallSolver(settings)
        nF, P, M = ws.eom.nbrFields, len(ws.boltzmannSolver.offEqParticles), ws.boltzmannSolver.grid.M
        for i in (0, len(vws) // 2, len(vws) - 1):
            row = vwscan.unpackRow(table[i], nF, P, M)
            print(f"vw = {vws[i]:.4f}: pressure {row['pressure']:.6e}, widths {row['widths']}, {int(row['solves'])} solves")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
