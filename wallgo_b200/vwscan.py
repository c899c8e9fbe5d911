"""Wall-velocity scan: `EOM.wallPressure(vw_i, wallParams0, boltzmannResultsInput=None)` for many vw
(/root/reference/src/WallGo/equationOfMotion.py:786-1131), the second half of the BASELINE metric.

A scan point is the reference's own, unmodified pressure iteration: ~4 Boltzmann solves interleaved with
host work that only the reference's control plane can do (hydrodynamic matching, the per-z temperature
root finds that call the user's effective potential, the Nelder-Mead minimisation of the action).  After
the Boltzmann solve has moved to the GPU that host work is >90 % of a point, so one Python process per GPU
cannot keep a B200 busy.  Layout per rank (one process per GPU, as everywhere in this package):

  * the rank process owns the GPU and acts as a SOLVE SERVER: a slot pool of independent systems in flight
    (BoltzmannSolver.slotPool: one engine, operator buffer, stream and captured CUDA graphs per slot);
  * `workers` forked host processes each run the unmodified `EOM.wallPressure` on their own
    EOM / Grid3Scales (the grid is re-mapped per vw, equationOfMotion.py:1457-1503) with a
    `WallGo.BoltzmannSolver` subclass whose getDeltas() ships the O(P*M) boosted background numbers to the
    server over a pipe and receives deltaF, the four moments and the truncation outputs back (127 KB);
  * points are handed out dynamically (the iteration count varies per vw); ranks take points round-robin
    and one all_gather of the per-point rows ends the scan (wallgo_b200.scan.gatherPoints).

  * FACTOR RE-USE (default, `reuseFactors`): the ~4 solves of a point share wall velocity and grid mapping, so the
    server gives every worker its own slot and, from the third solve of a point on, reaches the solution by
    refinement steps preconditioned with the LU factors of the point's last factorisation instead of factoring
    again (BoltzmannSolver.slotPool(reuseFactors=True)); workers send the position of each solve within its point.

Every point starts from the same (wallParams0, deltaF = 0), the device path is batch-invariant and the re-use
decisions depend only on the point's own sequence, so the table is identical for any number of workers and GPUs
(checked by a checksum in bench.py and by the tests); without re-use it is also bit-identical to the sequential
in-process loop, with re-use it agrees with it to ~1e-9 (deltaF to ~1e-13).
`workers=0` runs the points sequentially in the rank process through the plain drop-in class.

Needs WallGo (the reference) importable, like wallgo_b200.dropin.
"""
from __future__ import annotations

import os
import time

import numpy as np


def rowWidth(nFields: int, P: int, M: int) -> int:
    """[pressure, widths (nFields), offsets (nFields), iterations, Delta00|02|20|11 (4*P*(M-1))]"""
    return 2 + 2 * nFields + 4 * P * (M - 1)


def _row(out, nSolves):
    pressure, wallParams, res = out[0], out[1], out[2]
    d = res.Deltas
    return np.concatenate([[float(pressure)], np.ravel(wallParams.widths), np.ravel(wallParams.offsets),
                           [float(nSolves)], d.Delta00.coefficients.ravel(), d.Delta02.coefficients.ravel(),
                           d.Delta20.coefficients.ravel(), d.Delta11.coefficients.ravel()])


def unpackRow(row, nFields: int, P: int, M: int):
    """One row of the scan table -> dict(pressure, widths, offsets, solves, Deltas [4, P, M-1]); NaN rows are
    points that were not started within the budget."""
    row = np.asarray(row)
    assert row.size == rowWidth(nFields, P, M)
    o = 1 + 2 * nFields
    return dict(pressure=float(row[0]), widths=row[1:1 + nFields].copy(), offsets=row[1 + nFields:o].copy(),
                solves=row[o], Deltas=row[o + 1:].reshape(4, P, M - 1).copy())


def _initialWallParams(WallGo, ws):
    eom = ws.eom
    eom.pressAbsErrTol = 1e-8          # EOM.solveWall sets this before its first wallPressure (:474)
    return WallGo.WallParams(widths=ws.initialWallThickness * np.ones(eom.nbrFields),
                             offsets=np.zeros(eom.nbrFields))


def makeRemoteClass(conn):
    """Subclass of WallGo.BoltzmannSolver for the host workers: setBackground as the reference, getDeltas()
    by remote procedure call to the rank's solve server."""
    import WallGo  # pylint: disable=import-outside-toplevel
    import WallGo.results  # noqa: F401  pylint: disable=import-outside-toplevel
    from WallGo.boltzmann import BoltzmannSolver as _RefSolver  # pylint: disable=import-outside-toplevel

    from .boltzmann import BoltzmannSolverB200Mixin

    class _WallGoTypes:
        Polynomial = WallGo.Polynomial
        BoltzmannDeltas = WallGo.BoltzmannDeltas
        BoltzmannResults = WallGo.results.BoltzmannResults

    class RemoteBoltzmannSolver(BoltzmannSolverB200Mixin, _RefSolver):
        _types = _WallGoTypes
        nSolves = 0
        pointSeq = 0      # position of the next solve within the current scan point (reset by the worker loop)

        def __init__(self, grid, basisM="Cardinal", basisN="Chebyshev", derivatives="Spectral",
                     collisionMultiplier=1.0, truncationOption=None, device=None):
            from WallGo.boltzmann import ETruncationOption  # pylint: disable=import-outside-toplevel
            self._initB200(grid, basisM, basisN, derivatives, collisionMultiplier,
                           ETruncationOption.AUTO if truncationOption is None else truncationOption, device)

        def loadCollisions(self, directoryPath):     # the server holds the tensor
            pass

        def _getEngine(self):
            raise RuntimeError("host worker: the GPU belongs to the rank's solve server")

        def getDeltas(self, deltaF=None):
            assert deltaF is None, "host workers only forward full solves"
            bg = self.background
            dxidchi, _, _ = self.grid.getCompactificationDerivatives()
            conn.send(("solve", np.asarray(bg.temperatureProfile, dtype=np.float64),
                       np.asarray(bg.velocityProfile, dtype=np.float64), self._msq(bg.fieldProfiles),
                       np.asarray(dxidchi, dtype=np.float64), float(bg.velocityWall), int(type(self).pointSeq)))
            type(self).pointSeq += 1
            kind, payload = conn.recv()
            if kind == "error":
                raise np.linalg.LinAlgError(payload)
            dF, deltas, err, keepList, peaks = payload
            type(self).nSolves += 1
            return self._wrapResults(dF, deltas, err, keepList, peaks)

    return RemoteBoltzmannSolver


def _workerMain(conn, manager, settings):
    """Host worker: unmodified EOM.wallPressure, Boltzmann solves forwarded to the server."""
    import gc  # pylint: disable=import-outside-toplevel

    # The worker never touches CUDA.  If the rank process had initialised CUDA before the fork, garbage inherited
    # from it (cyclic references holding device tensors) must not be collected HERE, where the driver is unusable:
    # everything that exists at this point is moved out of the collector's reach.
    gc.freeze()
    import WallGo  # pylint: disable=import-outside-toplevel
    import WallGo.manager as _manager  # pylint: disable=import-outside-toplevel

    try:
        try:      # one core per host worker: the EOM's NumPy calls are small, BLAS threads would only collide
            from threadpoolctl import threadpool_limits  # pylint: disable=import-outside-toplevel
            threadpool_limits(limits=1)
        except Exception:  # pylint: disable=broad-except
            pass
        cls = makeRemoteClass(conn)
        _manager.BoltzmannSolver = cls
        ws = manager.setupWallSolver(settings)
        wp0 = _initialWallParams(WallGo, ws)
        conn.send(("ready",))
        while True:
            msg = conn.recv()
            if msg[0] == "stop":
                break
            _, idx, vw = msg
            before = cls.nSolves
            cls.pointSeq = 0            # the solves of one wallPressure form one sequence of nearby systems
            t0 = time.perf_counter()
            out = ws.eom.wallPressure(float(vw), wp0, boltzmannResultsInput=None)
            conn.send(("done", idx, _row(out, cls.nSolves - before), time.perf_counter() - t0))
    except BaseException as exc:  # pylint: disable=broad-except
        import traceback
        try:
            conn.send(("crash", traceback.format_exc()))
        except Exception:  # pylint: disable=broad-except
            pass
        raise exc
    finally:
        conn.close()


class HostWorkers:
    """The forked host processes of one rank.  Start them as early as possible -- ideally before the process
    has initialised CUDA -- and hand them to scanLocal / scanWallPressure; they idle on their pipe until then."""

    def __init__(self, manager, settings, count: int):
        import gc  # pylint: disable=import-outside-toplevel
        import multiprocessing as mp  # pylint: disable=import-outside-toplevel

        gc.collect()        # no pending garbage (possibly holding device tensors) goes into the children
        ctx = mp.get_context("fork")
        self.conns, self.procs = [], []
        for _ in range(max(0, int(count))):
            parent, child = ctx.Pipe(duplex=True)
            pr = ctx.Process(target=_workerMain, args=(child, manager, settings), daemon=True)
            pr.start()
            child.close()
            self.conns.append(parent)
            self.procs.append(pr)

    def __len__(self):
        return len(self.conns)

    def close(self):
        for c in self.conns:
            try:
                c.send(("stop",))
            except Exception:  # pylint: disable=broad-except
                pass
            try:
                c.close()
            except Exception:  # pylint: disable=broad-except
                pass
        for pr in self.procs:
            pr.join(timeout=5)
            if pr.is_alive():
                pr.terminate()
        self.conns, self.procs = [], []


def _physicalGpus() -> int:
    """GPUs installed in this node (NVML: not narrowed by CUDA_VISIBLE_DEVICES, does not initialise CUDA)."""
    try:
        import pynvml  # pylint: disable=import-outside-toplevel
        pynvml.nvmlInit()
        try:
            return max(1, int(pynvml.nvmlDeviceGetCount()))
        finally:
            pynvml.nvmlShutdown()
    except Exception:  # pylint: disable=broad-except
        try:
            import torch  # pylint: disable=import-outside-toplevel
            return max(1, torch.cuda.device_count())
        except Exception:  # pylint: disable=broad-except
            return 1


def defaultWorkers(worldSize: int = 1) -> int:
    """Host processes per rank.  The node's cores are divided evenly among the GPUs INSTALLED in it, so the host
    budget of one GPU is the same whether a run uses one rank or all of them (a scan is bounded by the host loop
    of the reference, ~2 s of Python per point, long before the GPU: the scaling a run reports is then that of the
    GPUs with their share of the host, not that of idle cores); one core stays with the rank process when the share is at least six cores; at most 12
    (the solve server of one GPU keeps up with ~25 workers: ~0.05 s of GPU per solve against ~0.5 s of host Python
    per solve per worker)."""
    gpus = _physicalGpus()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    share = cores // max(gpus, worldSize)
    # the rank process (solve server) sleeps in select() most of the time: with a small share it does not get a core
    # of its own (an 8-GPU node with 32 cores: 4 workers per rank, not 3)
    return max(1, min(12, share - 1 if share >= 6 else share))


def _serverWallSolver(manager, settings):
    """The rank's own WallSolver (EOM + grid + drop-in Boltzmann solver with its engines, slots and captured
    graphs), built once per (manager, settings) and kept on the manager."""
    cached = getattr(manager, "_b200ScanSolver", None)
    if cached is None or cached[0] is not settings:
        cached = (settings, manager.setupWallSolver(settings))
        manager._b200ScanSolver = cached  # pylint: disable=protected-access
    return cached[1]


def reuseSlots(solver, workers: int, reuseFactors) -> int:
    """Slots of the solve server when factors are re-used across the solves of a scan point: one per host worker
    (a worker's sequence of systems must meet its own previous factors), or 0 when re-use is off or the operators
    of that many slots do not fit the GPU's free memory."""
    if not reuseFactors or workers <= 0:
        return 0
    eng = solver._getEngine()
    have = sum(1 for s_ in (getattr(solver, "_sibs", None) or [solver])
               if s_._engine is not None and s_._engine.A is not None)
    free = eng.torch.cuda.mem_get_info(eng.dev)[0]
    perSlot = 8 * eng.n * eng.n + 6 * 8 * 256 * eng.n + (64 << 20)
    return int(workers) if max(0, int(workers) - have) * perSlot < 0.8 * free else 0


REUSE_FACTORS = True      # default of scanLocal / scanWallPressure / warmUp


def warmUp(manager, settings, vw, concurrency=None, workers=0, reuseFactors=None):
    """One point in-process, then one batch through every slot: engines, operator buffers and CUDA graphs of the
    whole slot pool exist before a timed scan starts (with factor re-use: one slot per host worker)."""
    rows = scanLocal(manager, settings, [vw], [0], workers=0)
    solver = _serverWallSolver(manager, settings).boltzmannSolver
    reuseFactors = REUSE_FACTORS if reuseFactors is None else reuseFactors
    pool = solver.slotPool(reuseSlots(solver, workers, reuseFactors) or concurrency)
    count = len(pool.sibs)
    pool.close()
    import copy  # pylint: disable=import-outside-toplevel
    bg = copy.deepcopy(solver.background)
    bg.velocityMid = 0.0       # already in the plasma frame: the batch's own boost becomes the identity
    solver.getDeltasBatch([bg] * count, concurrency=count)
    return rows


def scanLocal(manager, settings, vws, indices, workers=None, concurrency=None, stats=None, budgetSeconds=None,
              hostWorkers=None, reuseFactors=None):
    """The points `indices` of `vws` on this process' GPU.  Returns [len(indices), width] rows in the order
    of `indices`.  `manager` is a WallGoManager after setupThermodynamicsHydrodynamics with
    wallgo_b200.dropin installed (dropin.install()); `settings` its WallSolverSettings.
    budgetSeconds: stop handing out points after this long (points are handed out in the order of `indices`;
    rows of points that were not started stay NaN and stats["done"] counts the finished ones).
    hostWorkers: a HostWorkers started earlier (closed by this call); None: forked here.
    reuseFactors (None: REUSE_FACTORS): from the third Boltzmann solve of a point on, the server first tries to
    reach the solution by refinement steps preconditioned with the LU factors of the point's last factorisation
    (BoltzmannSolver.slotPool(reuseFactors=True)); it needs one slot per host worker and is switched off when
    those do not fit the GPU.  stats["reuse"] = dict(fresh, stale, fallbacks, steps)."""
    from multiprocessing.connection import wait as mpWait

    import WallGo  # pylint: disable=import-outside-toplevel

    indices = list(indices)
    stats = {} if stats is None else stats
    ws = _serverWallSolver(manager, settings)       # the server's solver: holds the collision tensor
    solver = ws.boltzmannSolver
    nF, P, M = ws.eom.nbrFields, len(solver.offEqParticles), solver.grid.M
    rows = np.full((len(indices), rowWidth(nF, P, M)), np.nan)
    tStart = time.perf_counter()

    def overBudget():
        return budgetSeconds is not None and time.perf_counter() - tStart > budgetSeconds

    if hostWorkers is not None:
        workers = len(hostWorkers)
    elif workers is None:
        workers = defaultWorkers()
    workers = int(workers) if hostWorkers is not None else min(int(workers), len(indices))
    stats.update(workers=workers, host_s=0.0, solves=0, done=0)
    if not indices:
        if hostWorkers is not None:
            hostWorkers.close()
        return rows
    if workers <= 0:
        wp0 = _initialWallParams(WallGo, ws)
        for k, i in enumerate(indices):
            if overBudget():
                break
            t0 = time.perf_counter()
            calls = [0]
            orig = solver.getDeltas

            def counted(deltaF=None, _o=orig, _c=calls):
                _c[0] += 1
                return _o(deltaF)

            solver.getDeltas = counted
            try:
                out = ws.eom.wallPressure(float(vws[i]), wp0, boltzmannResultsInput=None)
            finally:
                del solver.getDeltas
            rows[k] = _row(out, calls[0])
            stats["host_s"] += time.perf_counter() - t0
            stats["solves"] += calls[0]
            stats["done"] += 1
        return rows

    # ---- solve server + forked host workers
    if hostWorkers is None:
        hostWorkers = HostWorkers(manager, settings, workers)
    conns = hostWorkers.conns
    reuseFactors = REUSE_FACTORS if reuseFactors is None else reuseFactors
    nReuse = reuseSlots(solver, len(conns), reuseFactors)
    if nReuse:
        pool = solver.slotPool(nReuse, reuseFactors=True)
        # factorisations in flight: what fills the GPU (the other slots wait their turn or iterate on stale factors)
        pool.maxFresh = int(concurrency) if concurrency else solver._autoSlotCount()
    else:
        pool = solver.slotPool(concurrency)
    slotOf = {c: k for k, c in enumerate(conns)} if nReuse else {}
    pool.rawResults = True

    def submit(c, payload):
        if nReuse:
            return pool.submit(c, raw=payload[:5], slot=slotOf[c], seq=payload[5])
        return pool.submit(c, raw=payload[:5])

    todo = list(enumerate(indices))[::-1]
    waiting = []                      # solve requests that found no free slot: (conn, raw)
    done = 0
    started = 0
    tServerBusy = 0.0
    try:
        live = set(conns)
        while live:
            ready = mpWait(list(live), timeout=0.0004 if (pool.busy() or waiting) else 0.05)
            for c in ready:
                msg = c.recv()
                kind = msg[0]
                if kind == "solve":
                    if not submit(c, msg[1:]):
                        waiting.append((c, msg[1:]))
                elif kind in ("ready", "done"):
                    if kind == "done":
                        _, k, row, secs = msg
                        rows[k] = row
                        stats["host_s"] += secs
                        stats["solves"] += int(row[1 + 2 * nF])
                        done += 1
                    if todo and not overBudget():
                        k, i = todo.pop()
                        c.send(("point", k, float(vws[i])))
                        started += 1
                    else:
                        c.send(("stop",))
                        live.discard(c)
                elif kind == "crash":
                    raise RuntimeError("scan worker failed:\n" + msg[1])
            t0 = time.perf_counter()
            for c, payload in pool.poll():
                dF, deltas, err, keep, peaks = payload
                c.send(("result", (dF, deltas, err, list(keep), peaks)))
            while waiting and pool.freeSlots():
                if not submit(*waiting[0]):
                    break              # factorisation limit reached: the head of the queue keeps its place
                waiting.pop(0)
            tServerBusy += time.perf_counter() - t0
    finally:
        pool.close()
        hostWorkers.close()
    stats["server_s"] = tServerBusy
    stats["done"] = done
    stats["reuse"] = dict(pool.reuseStats, slots=len(pool.sibs), enabled=bool(nReuse))
    assert done == started
    return rows


def scanWallPressure(manager, settings, vws, workers=None, concurrency=None, stats=None, budgetSeconds=None,
                     hostWorkers=None, reuseFactors=None):
    """All points of `vws`, sharded round-robin over the ranks of the default process group (one per GPU),
    gathered with ONE all_gather.  Every rank returns the full [len(vws), width] table (NaN rows: points a
    rank did not start within budgetSeconds)."""
    import torch.distributed as dist  # pylint: disable=import-outside-toplevel

    from . import scan  # pylint: disable=import-outside-toplevel

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = scan.shardPoints(len(vws), rank, world)
    if workers is None:
        workers = defaultWorkers(world)
    rows = scanLocal(manager, settings, vws, mine, workers=workers, concurrency=concurrency, stats=stats,
                     budgetSeconds=budgetSeconds, hostWorkers=hostWorkers, reuseFactors=reuseFactors)
    return scan.gatherPoints(rows, len(vws), rows.shape[1])
