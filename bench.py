#!/usr/bin/env python
"""Benchmark of the Boltzmann hot path (BASELINE.json metric: Boltzmann solves/s, assemble+LU+moments).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2a]

One "step" = one full pass of the hot path over one synthetic system: operator assembly, dense LU
factor-and-solve, spectral truncation and the four moment integrals.  Default workload is
BASELINE.json configs[1] (single fermion, M=40, N=21, n=15600, operator 1.95 GB >> 126 MB L2, so
every step streams from HBM; no explicit L2 flush is needed).  Multi-GPU: independent systems
(scan points) are sharded one process per GPU, weak scaling, one NCCL all_gather of the moments.

Prints ONE JSON line on rank 0 (contract in the task statement).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {  # BASELINE.json configs -> (P, M, N)
    "cfg1": (1, 20, 11), "cfg2a": (1, 40, 21), "cfg2b": (2, 40, 21), "cfg3": (3, 30, 11), "cfg4": (4, 30, 11),
    # configs[4] (P=4, M=64, N=31) needs 411 GB.  The member of that family that keeps the N=31 momentum grid and
    # fits one B200: M=32, n = 111600, operator 99.6 GB; and the M=64 member at N=21 (81 GB)
    "cfg5fit": (4, 32, 31), "cfg5fit21": (4, 64, 21),
    # the literal configs[4]: assembled in row panels that are never resident together (--workload cfg5stream)
    "cfg5stream": (4, 64, 31),
}
METRIC = "Boltzmann solves/sec (assemble+LU+moments)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_traffic(stem):
    """DRAM read+write bytes of one launch of a kernel, parsed from the committed `ncu --set full` summary
    profiles/rNN_ncu_full_<stem>.csv of the newest round that has one (first launch column).  Returns
    (bytes, file name) or (None, None): the number is never a literal in this file."""
    import glob
    import re

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r[0-9][0-9]_ncu_full_{stem}.csv")))
    if not files:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total = 0.0
    found = 0
    try:
        for ln in open(files[-1]):
            m = re.match(r"dram__bytes_(read|write)\.sum,([A-Za-z]+),([0-9.eE+-]+)", ln)
            if m:
                total += float(m.group(3)) * unit[m.group(2)]
                found += 1
    except (OSError, KeyError, ValueError):
        return None, None
    return (total, os.path.basename(files[-1])) if found == 2 else (None, None)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
                power.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_inputs(P, M, N, nsys, rank):
    """Deterministic synthetic systems; every step is a different wall velocity (a scan point)."""
    from wallgo_b200 import synthetic

    grid = synthetic.makeGrid(M, N)
    particles = synthetic.makeParticles(P)
    coll = synthetic.makeCollision(grid, particles)
    bgs = [synthetic.makeBackground(grid, velocityMid=-0.3 - 0.4 * ((i + rank * nsys) % 97) / 97.0)
           for i in range(nsys)]
    return grid, particles, coll, bgs


def cpu_port_one_solve(grid, particles, coll, bg):
    """The oracle (NumPy/LAPACK port of the reference path) on the host cores: one full solve."""
    from oracle import boltzmann_oracle as bo
    from copy import deepcopy

    b = deepcopy(bg)
    b.boostToPlasmaFrame()
    tab = bo.build_tables(grid.M, grid.N)
    pb = bo.Problem(tab=tab, xi_dxidchi=grid.dxidchi, pz=grid.pzValues, pp=grid.ppValues, dpzdrz=grid.dpzdrz,
                    dppdrp=grid.dppdrp, T=b.temperatureProfile, v=b.velocityProfile,
                    msq=np.array([p.msqVacuum(b.fieldProfiles) for p in particles]), vw=float(b.velocityWall),
                    statistics=np.array([-1.0 if p.statistics == "Fermion" else 1.0 for p in particles]),
                    collision=np.asarray(coll[...]), collisionMultiplier=1.0)
    # all host cores for BLAS, even under torchrun (which exports OMP_NUM_THREADS=1)
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=os.cpu_count())
    except Exception:
        limiter = None
    try:
        t0 = time.perf_counter()
        op, src = bo.assemble(pb)
        t1 = time.perf_counter()
        dF = np.linalg.solve(op, src).reshape(len(particles), grid.M - 1, grid.N - 1, grid.N - 1)
        t2 = time.perf_counter()
        del op
        out = bo.get_deltas(pb, dF, "AUTO")
        t3 = time.perf_counter()
    finally:
        if limiter is not None:
            limiter.restore_original_limits()
    return dict(assemble_s=t1 - t0, dgesv_s=t2 - t1, moments_s=t3 - t2, total_s=t3 - t0), out


def reference_available():
    from oracle import ref_drivers as rd
    return rd.reference_root() is not None


def reference_one_solve(P, M, N, velocityMid, want_solver=False):
    """The UNMODIFIED reference class (WallGo.BoltzmannSolver from oracle/_ref, pure Python: NumPy broadcasting
    + LAPACK dgesv through np.linalg.solve) on the host cores: one full solve of the synthetic system of
    SURVEY.md section 8(d), the stages timed as the reference runs them (boltzmann.py:141-294, 628-820)."""
    from oracle import ref_drivers as rd

    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=os.cpu_count())
    except Exception:
        limiter = None
    try:
        solver, grid, bg, particles = rd.synthetic_solver(P, M, N, velocityMid=velocityMid)
        t0 = time.perf_counter()
        op, src, _, _ = solver.buildLinearEquations()
        t1 = time.perf_counter()
        dF = np.linalg.solve(op, src).reshape(P, M - 1, N - 1, N - 1)       # boltzmann.py:283
        t2 = time.perf_counter()
        del op
        res = solver.getDeltas(dF)
        t3 = time.perf_counter()
    finally:
        if limiter is not None:
            limiter.restore_original_limits()
    t = dict(assemble_s=t1 - t0, dgesv_s=t2 - t1, moments_s=t3 - t2, total_s=t3 - t0)
    if want_solver:
        return t, res, (solver, grid, bg, particles)
    return t, res


def blas_info():
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        with threadpool_limits(limits=os.cpu_count()):
            return ";".join(f"{d.get('internal_api')} {d.get('version')} threads={d.get('num_threads')}"
                            for d in threadpool_info())
    except Exception:
        return "unknown"


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P, M, N = WORKLOADS[a.workload]
    real = reference_available()
    grid, particles, coll, bgs = build_inputs(P, M, N, a.steps + a.warmup, 0)
    budget = float(a.ref_budget_s)

    def one(i):
        if real:     # same wall velocities as the b200 arm's steps (build_inputs)
            return reference_one_solve(P, M, N, -0.3 - 0.4 * (i % 97) / 97.0)
        return cpu_port_one_solve(grid, particles, coll, bgs[i])

    for i in range(min(a.warmup, 1)):
        one(i)
    t0 = time.perf_counter()
    done, parts = 0, []
    for i in range(a.steps):
        t, _ = one(a.warmup + i)
        parts.append(t)
        done += 1
        if time.perf_counter() - t0 > budget:
            break
    el = sum(p_["total_s"] for p_ in parts)     # the hot path only: building the synthetic inputs is not a step
    val = done / el
    n = P * (M - 1) * (N - 1) ** 2
    sample = (f"{done} of {a.steps} requested full solves (assemble+dgesv+moments) at P={P},M={M},N={N}, "
              f"time budget {budget:.0f}s; NumPy {np.__version__}, BLAS {blas_info()}; "
              f"mean assemble {np.mean([p['assemble_s'] for p in parts]):.2f}s dgesv "
              f"{np.mean([p['dgesv_s'] for p in parts]):.2f}s")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "solves/s", "n_gpus": a.gpus,
            "steps": done, "warmup": min(a.warmup, 1), "ms_per_step": 1e3 * el / done, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{a.workload}: P={P} M={M} N={N} n={n}, synthetic collision tensor",
                       "note": ("the unmodified reference class WallGo.BoltzmannSolver (pure Python copy under "
                                "oracle/_ref): buildLinearEquations + np.linalg.solve + getDeltas" if real else
                                "oracle/_ref absent on this machine: the oracle port (NumPy broadcasting + LAPACK "
                                "dgesv) of the same path")},
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": os.cpu_count(),
                             "kind": "reference" if real else "port", "sample": sample},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if a.scan_points > 0 and real and a.ref_scan_points > 0:
        try:
            import contextlib
            with contextlib.redirect_stdout(sys.stderr):      # the reference prints its phase tracing to stdout
                line["scan"] = reference_scan(a)
        except Exception as exc:  # the main line must survive a failure of the secondary metric
            line["scan"] = {"error": repr(exc)[:300]}
    print(json.dumps(line), flush=True)


SCAN_MODELS = {"idm": (4, 30, 11), "yukawa": (2, 20, 11), "singlet": (1, 20, 11), "sm": (3, 30, 11)}


def scan_velocities(manager, npts):
    """256 wall velocities uniformly in [vMin, 0.53] for the InertDoublet driver (vJ = 0.5333): the valid
    deflagration range of equationOfMotion.py:186-187."""
    hyd = manager.hydrodynamics
    hi = min(0.53, 0.995 * hyd.vJ)
    lo = max(float(hyd.vMin), 0.01)
    return np.linspace(lo, hi, npts)


def reference_scan(a):
    """CPU anchor of the scan metric: the unmodified reference (its own BoltzmannSolver on the host) on a
    bounded sample of the same scan points."""
    from oracle import ref_drivers as rd

    WallGo = rd.load_reference()
    P, M, N = SCAN_MODELS[a.scan_model]
    rd.patch_load_collisions(WallGo.boltzmann.BoltzmannSolver)
    t0 = time.perf_counter()
    manager, settings, _ = rd.build_manager(a.scan_model, M, N)
    setup = time.perf_counter() - t0
    vws = scan_velocities(manager, a.scan_points)
    ws = manager.setupWallSolver(settings)
    eom = ws.eom
    eom.pressAbsErrTol = 1e-8
    wp0 = WallGo.WallParams(widths=ws.initialWallThickness * np.ones(eom.nbrFields), offsets=np.zeros(eom.nbrFields))
    idx = list(np.linspace(0, len(vws) - 1, a.ref_scan_points).astype(int))
    t0 = time.perf_counter()
    press = []
    for i in idx:
        out = eom.wallPressure(float(vws[i]), wp0, boltzmannResultsInput=None)
        press.append(float(out[0]))
    el = time.perf_counter() - t0
    return {"metric": "vw scan pts/sec", "value": len(idx) / el, "unit": "points/s", "points": len(idx),
            "seconds": el, "setup_seconds": setup, "point_indices": [int(i) for i in idx], "pressures": press,
            "workload": f"{a.scan_model}: P={P} M={M} N={N}, EOM.wallPressure(vw_i, wallParams0, None), "
                        f"{len(idx)} of the {a.scan_points} scan points, host cores {os.cpu_count()}"}


def prepare_scan(a, world):
    """Host side of the scan, done BEFORE this process initialises CUDA: the reference's WallGoManager for the
    scan model (phase tracing, hydrodynamics: tens of seconds of host work) and the forked host workers, which
    then idle on their pipes until the scan starts."""
    import contextlib

    from oracle import ref_drivers as rd
    if rd.reference_root() is None:
        return None
    with contextlib.redirect_stdout(sys.stderr):
        rd.load_reference()
        from wallgo_b200 import dropin, vwscan

        P, M, N = SCAN_MODELS[a.scan_model]
        cls = dropin.install(vectoriseEOM=True)
        rd.patch_load_collisions(cls)      # shipped .hdf5 are Git-LFS pointers: relaxation-time tensor, both arms
        t0 = time.perf_counter()
        manager, settings, _ = rd.build_manager(a.scan_model, M, N)
        setup = time.perf_counter() - t0
        vws = scan_velocities(manager, a.scan_points)
        nworkers = vwscan.defaultWorkers(world) if a.scan_workers < 0 else a.scan_workers
        hostWorkers = vwscan.HostWorkers(manager, settings, nworkers) if nworkers > 0 else None
    return {"manager": manager, "settings": settings, "vws": vws, "setup_s": setup, "hostWorkers": hostWorkers}


def run_scan(a, world, rank, local, prep=None):
    """Second half of the BASELINE metric: wall-velocity scan points/s, a point being what SURVEY.md section
    8(d)(2) defines: one `EOM.wallPressure(vw_i, wallParams0, boltzmannResultsInput=None)` of the UNMODIFIED
    reference wall solver (equationOfMotion.py:786-1131; ~4 Boltzmann solves + host loop) on the InertDoublet
    driver (P=4, M=30, N=11, n=11600), 256 vw in [vMin, 0.53], strong scaling: ranks take points round-robin,
    each rank with its own WallGoManager / EOM / Grid3Scales, Boltzmann solves on the rank's GPU through
    wallgo_b200.dropin (vectoriseEOM=True), pressures + wall parameters + moments gathered by ONE all_gather.
    Needs the reference copy under oracle/_ref; without it the Boltzmann-only sweep below is reported."""
    import contextlib

    import torch
    import torch.distributed as dist

    from oracle import ref_drivers as rd
    if rd.reference_root() is None:
        return run_scan_boltzmann_only(a, world, rank, local)
    if prep is None and a.scan_points > 0:
        raise RuntimeError("scan preparation failed (see stderr)")
    prep = prep if prep is not None else prepare_scan(a, world)
    P, M, N = SCAN_MODELS[a.scan_model]
    with contextlib.redirect_stdout(sys.stderr):
        from wallgo_b200 import _lib, vwscan

        manager, settings, vws, setup = prep["manager"], prep["settings"], prep["vws"], prep["setup_s"]
        workers = len(prep["hostWorkers"]) if prep["hostWorkers"] is not None else 0
        # warm-up: engines, buffers and CUDA graphs of every slot (in-process, one point + one batch)
        reuse = bool(a.scan_reuse)
        vwscan.warmUp(manager, settings, float(vws[rank % len(vws)]), workers=workers, reuseFactors=reuse)
        lib = _lib.load()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = lib.wg_launch_count()
        stats = {}
        t0 = time.perf_counter()
        table = vwscan.scanWallPressure(manager, settings, vws, workers=workers, stats=stats,
                                        budgetSeconds=a.scan_budget_s, hostWorkers=prep["hostWorkers"],
                                        reuseFactors=reuse)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        launches = lib.wg_launch_count() - l0
    mine = len(range(rank, len(vws), world))
    agg = torch.tensor([el, stats.get("host_s", 0.0), float(stats.get("solves", 0)), float(launches)],
                       dtype=torch.float64, device=torch.device("cuda", local))
    if world > 1:
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        el = float(mx[0].item())
    hostTotal, solves, launches = float(agg[1].item()), int(agg[2].item()), int(agg[3].item())
    nF = (table.shape[1] - 2 - 4 * P * (M - 1)) // 2
    finished = np.all(np.isfinite(table), axis=1)
    ndone = int(np.sum(finished))
    assert ndone > 0
    n = P * (M - 1) * (N - 1) ** 2
    head = min(64, len(vws))
    return {"metric": "vw scan pts/sec", "value": ndone / el, "unit": "points/s", "points": ndone,
            "points_requested": len(vws), "budget_seconds": a.scan_budget_s,
            "seconds": el, "scaling": "strong", "setup_seconds_per_rank": setup,
            "workload": f"{a.scan_model}: P={P} M={M} N={N} n={n}, vw in [{vws[0]:.3f}, {vws[-1]:.3f}], "
                        "relaxation-time collision tensor",
            "point": "EOM.wallPressure(vw_i, wallParams0, None) of the unmodified reference wall solver, Boltzmann "
                     "solves on the GPU via wallgo_b200.dropin (vectoriseEOM), LU factors re-used across the solves of a "
                     "point unless --scan-reuse 0 (factor_reuse below); one all_gather of the result rows",
            "host_workers_per_rank": int(workers), "points_per_rank": mine,
            "host_cores": os.cpu_count(), "gpus_in_node": vwscan._physicalGpus(),
            "host_budget": "workers per rank = min(12, share - 1 if share >= 6 else share), share = cores // GPUs installed in the node: the same host share per "
                           "GPU at every rank count, so the scan's strong scaling is that of the GPUs with their share of "
                           "the host.  A point costs ~0.7 s of the reference's host Python per worker and ~0.15 s of GPU "
                           "(3.7 factorisations; ~0.09 s with factor re-use: 2 factorisations + 1.7 stale-factor solves): "
                           "5-8 workers saturate one GPU's solve server, fewer leave the scan host-bound",
            "boltzmann_solves": solves, "solves_per_point": solves / max(1, ndone),
            "host_seconds_per_point": hostTotal / max(1, ndone),
            "host_note": "wall time a point spends in the reference's host loop + waiting for its solves (queueing "
                         "included), summed over the host workers; the GPU time of one solve at this size is ~0.04 s",
            "gpu_launches": launches,
            "factor_reuse": dict(stats.get("reuse", {"enabled": False}),
                                 note="rank 0's server: from the third Boltzmann solve of a point on, the solution is "
                                      "reached by refinement steps (double-double residual of the regenerated operator) "
                                      "preconditioned with the LU factors of the point's last factorisation; `fresh` "
                                      "factorisations, `stale` solves that converged that way in `steps` refinement "
                                      "steps, `fallbacks` to a fresh factorisation (counted in `fresh` too)"),
            "checksum_first64": (float(np.sum(table[:head, 0])) if bool(np.all(finished[:head])) else None),
            "checksum_moments_first64": (float(np.sum(table[:head, 2 + 2 * nF:])) if bool(np.all(finished[:head])) else None),
            "checksum_note": "sum of the pressures / of all moments of scan points 0..63: equal for every GPU and "
                             "worker count (each point starts from the same guess; the device path is batch-invariant)",
            "pressure_first": float(table[0, 0]),
            "pressure_last": (float(table[-1, 0]) if bool(finished[-1]) else None),
            "pressure_note": "compare with `pressures` of the reference arm's scan block (points 0 and 255): the "
                             "1e-6 gate of the north star applies"}


def run_scan_boltzmann_only(a, world, rank, local):
    """Fallback when the reference copy is absent: Boltzmann-only sweep.  256 independent vw points of the
    cfg4-sized system (P=4, M=30, N=11, n=11600) sharded round-robin over the ranks (strong scaling), every
    point one full assemble+LU+moments through the public class with host buffers, one NCCL all_gather of
    the moment tables at the end.  A point here is ONE Boltzmann solve at that vw: the EOM pressure
    iteration around it is the reference's own unmodified host code (out of scope)."""
    import torch
    import torch.distributed as dist

    import wallgo_b200 as wb
    from wallgo_b200 import scan, synthetic

    P, M, N = WORKLOADS[a.scan_workload]
    npts = a.scan_points
    grid = synthetic.makeGrid(M, N)
    particles = synthetic.makeParticles(P)
    coll = synthetic.makeCollision(grid, particles)
    bgs = [synthetic.makeBackground(grid, velocityMid=-(0.05 + 0.48 * i / max(1, npts - 1))) for i in range(npts)]
    solver = wb.BoltzmannSolver(grid, device=local)
    solver.updateParticleList(particles)
    solver.setCollisionArray(coll)
    scan.scanDeltas(lambda: solver, bgs[:4 * world])      # warm-up: graphs captured, slots allocated
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    table = scan.scanDeltas(lambda: solver, bgs)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([el], dtype=torch.float64, device=torch.device("cuda", local))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())
    assert table.shape == (npts, 4 * P * (M - 1)) and np.all(np.isfinite(table))
    n = P * (M - 1) * (N - 1) ** 2
    out = {"metric": "vw scan pts/sec", "value": npts / el, "unit": "points/s", "points": npts, "seconds": el,
           "scaling": "strong", "workload": f"{a.scan_workload}: P={P} M={M} N={N} n={n}, vw in [0.05, 0.53]",
           "point": "one Boltzmann solve (assemble+LU+truncation+moments) per vw through getDeltasBatch, host buffers "
                    "in/out; NCCL all_gather of the moment tables",
           "concurrency_per_gpu": wb.BoltzmannSolver.autoConcurrency(n),
           "checksum": float(np.sum(table[:, :P * (M - 1)]))}
    del solver
    torch.cuda.empty_cache()
    return out


def run_stream(a, rank, world, local):
    """BASELINE configs[4] as written (P=4, M=64, N=31: n = 226800, a 411 GB operator) cannot be resident on one
    B200.  This mode generates it anyway, in row panels of `--panel-groups` (species, position) groups
    (q = 900 rows x n columns each) written round-robin into two panel buffers -- the access pattern of an
    out-of-core consumer -- and reports the assembly rate against the HBM write roofline.  One step = the whole
    operator once.  Multi-GPU: every rank streams its own operator (weak scaling, no collective)."""
    import torch
    import torch.distributed as dist

    import wallgo_b200 as wb
    from wallgo_b200 import _lib, synthetic
    from wallgo_b200._lib import check

    lib = _lib.load()          # (device and process group were set up by run_b200)
    P, M, N = WORKLOADS[a.workload]
    K, W = a.steps, a.warmup
    grid = synthetic.makeGrid(M, N)
    particles = synthetic.makeParticles(P)
    solver = wb.BoltzmannSolver(grid, device=local)
    solver.updateParticleList(particles)
    solver.setCollisionArray(synthetic.makeCollision(grid, particles))
    solver.setBackground(synthetic.makeBackground(grid, velocityMid=-0.45))
    eng = solver._getEngine()
    solver._uploadBackground(eng)
    n, q, dev = eng.n, eng.q, eng.dev
    G = a.panel_groups
    groups = P * (M - 1)
    panels = [torch.empty((G * q, n), dtype=torch.float64, device=dev) for _ in range(2)]
    sums = torch.zeros(1, dtype=torch.float64, device=dev)
    vw, mult = float(solver.background.velocityWall), 1.0
    st = eng._stream()

    def one_pass(check_sum=False):
        for i, g0 in enumerate(range(0, groups, G)):
            ng = min(G, groups - g0)
            pnl = panels[i % 2]
            check(lib.wg_assemble_rows(eng.h, eng.prof.data_ptr(), vw, mult, 3, g0, ng, pnl.data_ptr(), n,
                                       eng.S.data_ptr(), st), "wg_assemble_rows")
            if check_sum:
                sums.add_(pnl[:ng * q].sum())

    for _ in range(W):
        one_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.wg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    launches = lib.wg_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # end to end: the same pass with every panel reduced to a checksum that is read back (a consumer touching
    # every byte); host buffers in: the background profiles
    t0 = time.perf_counter()
    solver._uploadBackground(eng)
    one_pass(check_sum=True)
    total = float(sums.item())
    e2e_s = time.perf_counter() - t0
    if rank == 0:
        hbm_peak, hbm_src = measured_peaks()
        bytes_pass = 8.0 * n * n
        gbs = K * bytes_pass / (ms * 1e-3) / 1e9
        print(json.dumps({
            "metric": "operator assembly GB/s (streamed row panels)", "value": world * gbs, "unit": "GB/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg5stream: P={P} M={M} N={N} n={n}: the literal BASELINE configs[4] operator "
                                   f"({bytes_pass / 1e9:.0f} GB) generated in {wg_cdiv(groups, G)} row panels of {G * q} rows "
                                   f"({8.0 * G * q * n / 1e9:.1f} GB each, two buffers, never resident together)",
                       "l2": "every panel is far larger than the 126 MB L2"},
            "roofline": {"kernel": "wg_assemble_kernel (row-panel mode)", "bound": "hbm", "achieved": gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None, "peak_source": hbm_src,
                         "algorithmic": "8 n^2 bytes written per pass / CUDA-event time of the pass"},
            "e2e": {"value": bytes_pass / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(eng.nprof * 8),
                    "d2h_bytes_per_step": 8, "note": "one pass with every panel summed on the device and the checksum "
                                                     "read back", "checksum": total},
            "gpu_launches": int(launches), "clocks": clocks,
            "cpu_baseline": {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "reference",
                             "sample": "skipped: buildLinearEquations materialises the 411 GB operator in host RAM"}}),
            flush=True)
    if world > 1:
        dist.destroy_process_group()


def wg_cdiv(a, b):
    return (a + b - 1) // b


def run_small(a, rank, world, local, lib, fp64_peak_fn):
    """Small systems (n <= BoltzmannSolver.STRIDED_MAX_N; BASELINE configs[0], n = 1900): the strided-batch path.
    One step = one lock-step batch of `--batch` independent systems (assemble + batched LU + solves + refinement
    + truncation + moments, every kernel launched once for the whole batch), two batches in flight on two streams.
    value = systems/s with all inputs resident in HBM; e2e = BoltzmannSolver.getDeltasBatch with host buffers."""
    import torch
    import torch.distributed as dist

    import wallgo_b200 as wb
    from wallgo_b200.engine import BatchEngine

    P, M, N = WORKLOADS[a.workload]
    K, W, B = a.steps, a.warmup, a.batch
    nsys = B * (K + W)
    grid, particles, coll, bgs = build_inputs(P, M, N, nsys, rank)
    solver = wb.BoltzmannSolver(grid, device=local)
    solver.updateParticleList(particles)
    solver.setCollisionArray(coll)
    solver.setBackground(bgs[0])
    eng = solver._getEngine()
    n, nprof, dev = eng.n, eng.nprof, eng.dev
    fp64_peak = fp64_peak_fn(dev)
    F = max(1, int(a.inflight))
    bes = [BatchEngine(eng, B) for _ in range(F)]
    inp = np.empty((K + W, B, nprof + 1))
    L = M + 1
    for i, bg in enumerate(bgs):
        solver.setBackground(bg)
        b = solver.background
        row = inp[i // B, i % B]
        row[:L] = b.temperatureProfile
        row[L:2 * L] = b.velocityProfile
        row[2 * L:(2 + P) * L] = solver._msq(b.fieldProfiles).reshape(-1)
        row[(2 + P) * L:nprof] = grid.dxidchi
        row[nprof] = float(b.velocityWall)
    inp_d = torch.from_numpy(inp).to(dev)
    keep = torch.tensor([[M - 1, N - 1, N - 1]] * B, dtype=torch.int32, device=dev)
    from wallgo_b200._lib import check

    def step(be, i, rec=None):
        with torch.cuda.stream(be.stream):
            be.prof.copy_(inp_d[i, :, :nprof])
            be.vw.copy_(inp_d[i, :, nprof])
            sp, sA = be._sp, be.strideA
            if rec:
                rec[0].record(be.stream)
            check(lib.wg_batch_assemble(be.h, be.prof.data_ptr(), be.vw.data_ptr(), 1.0, 3, be.A.data_ptr(), n, sA,
                                        be.S.data_ptr(), sp), "wg_batch_assemble")
            if rec:
                rec[1].record(be.stream)
            check(lib.wg_batch_factor(be.h, be.A.data_ptr(), n, sA, sp), "wg_batch_factor")
            if rec:
                rec[2].record(be.stream)
            check(lib.wg_batch_solve(be.h, be.A.data_ptr(), n, sA, be.S.data_ptr(), sp), "wg_batch_solve")
            if rec:
                rec[3].record(be.stream)
            if solver.refinementSteps:
                check(lib.wg_batch_refine(be.h, be.prof.data_ptr(), be.vw.data_ptr(), 1.0, be.A.data_ptr(), n, sA,
                                          be.S.data_ptr(), be.refstats.data_ptr(), sp), "wg_batch_refine")
            if rec:
                rec[4].record(be.stream)
            check(lib.wg_batch_spectra(be.h, be.S.data_ptr(), be.spectra.data_ptr(), sp), "wg_batch_spectra")
            check(lib.wg_batch_moments(be.h, be.S.data_ptr(), be.prof.data_ptr(), keep.data_ptr(), 1, be.dF.data_ptr(),
                                       be.deltas.data_ptr(), be.err.data_ptr(), sp), "wg_batch_moments")
            if rec:
                rec[5].record(be.stream)

    def barrier():
        if world > 1:
            dist.barrier()

    for i in range(max(W, F)):
        step(bes[i % F], i % W)
    torch.cuda.synchronize()
    rec = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    step(bes[0], 0, rec)
    torch.cuda.synchronize()
    stage = [rec[j].elapsed_time(rec[j + 1]) for j in range(5)]
    sampler = ClockSampler(local)
    barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    l0 = lib.wg_launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    start.record(cur)
    for be in bes:
        be.stream.wait_event(start)
    for i in range(K):
        step(bes[i % F], W + i)
    for be in bes:
        done = torch.cuda.Event()
        done.record(be.stream)
        cur.wait_event(done)
    end.record(cur)
    torch.cuda.synchronize()
    launches = lib.wg_launch_count() - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = start.elapsed_time(end)
    for be in bes:
        assert not any(be.infos()), "singular operator in the benchmark"
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * K * B / (ms / 1e3)
    for be in bes:
        be.close()
    # ---- end to end through the public class, host buffers
    solver.getDeltasBatch(bgs[:2 * B])
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    allres = solver.getDeltasBatch(bgs[W * B:])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = world * K * B / e2e_s
    t0 = time.perf_counter()
    for i in range(5):
        solver.setBackground(bgs[i])
        solver.getDeltas()
    seq_ms = (time.perf_counter() - t0) / 5 * 1e3
    if rank != 0:
        return
    lu_flops = 2.0 / 3.0 * float(n) ** 3
    lu_tf = K * B * lu_flops / (ms * 1e-3) / 1e12
    hbm_peak, hbm_src = measured_peaks()
    line = {
        "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{a.workload}: P={P} M={M} N={N} n={n}; one step = one lock-step strided batch of {B} "
                               "independent systems (wg_batch_*: every kernel launched once per batch), two batches in flight",
                   "l2": f"batch operators {8.0 * n * n * B / 1e9:.2f} GB per batch >> 126 MB L2",
                   "batch": B, "refinement_steps": int(solver.refinementSteps),
                   "parallelism": f"{world} x independent batches, no collective inside a solve"},
        "stages_ms_one_batch": {"assemble": stage[0], "lu_factor": stage[1], "tri_solves": stage[2], "refinement_step": stage[3],
                                "spectra_moments": stage[4], "note": f"one batch of {B} systems alone on the GPU"},
        "latency_ms_sequential_getDeltas": seq_ms,
        "roofline": {"kernel": "whole batched LU (lock-step: leaf panels, TRSM, DMMA trailing updates of all systems)",
                     "bound": "tensor", "achieved": lu_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": lu_tf / fp64_peak,
                     "traffic": None,
                     "algorithmic": "steps * batch * 2/3 n^3 LU flops / CUDA-event duration of the whole timed region "
                                    "(assembly, solves, refinement and moments count as overhead)",
                     "lu_alone": {"achieved": B * lu_flops / (stage[1] * 1e-3) / 1e12, "unit": "TFLOP/s",
                                  "note": "wg_batch_factor of one batch alone"},
                     "peak_source": "measured in this run: torch.matmul (cuBLAS DGEMM) fp64 8192^3 best of 5; "
                                    "MEASURED_PEAKS.json has no FP64 entry"},
        "roofline_assembly": {"kernel": "wg_assemble_kernel (batched)", "bound": "hbm",
                              "achieved": B * (8.0 * n * n + 8.0 * n) / (stage[0] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                              "frac": B * (8.0 * n * n + 8.0 * n) / (stage[0] * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                              "peak_source": hbm_src},
        "e2e": {"value": e2e, "unit": "solves/s", "h2d_bytes_per_step": int(B * (nprof + 1) * 8 + B * 12),
                "d2h_bytes_per_step": int(B * (eng.nspec + eng.nout) * 8 + B * 4)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    res = allres[-1]
    assert np.all(np.isfinite(res.Deltas.Delta00.coefficients))
    if world == 1 and not a.no_cpu_baseline and reference_available():
        import contextlib
        from oracle import ref_drivers as rd
        from wallgo_b200 import dropin
        ts = []
        with contextlib.redirect_stdout(sys.stderr):
            for i in range(12):
                vmid = -0.3 - 0.4 * (i % 97) / 97.0
                t, rres, (rsolver, rgrid, rbg, rparts) = reference_one_solve(P, M, N, vmid, want_solver=True)
                ts.append(t["total_s"])
            new = dropin.makeDropInClass()(rgrid, rsolver.basisM, rsolver.basisN, rsolver.derivatives,
                                           rsolver.collisionMultiplier, rsolver.truncationOption, device=local)
            new.updateParticleList(rparts)
            new.setBackground(rbg)
            new.setCollisionArray(rsolver.collisionArray)
            gres = new.getDeltas()
            Dr, Dg = rd.deltas_table(rres), rd.deltas_table(gres)
        line["cpu_baseline"] = {"value": 1.0 / float(np.mean(ts[2:])), "unit": "solves/s", "cores": os.cpu_count(),
                                "kind": "reference",
                                "sample": f"10 full solves of the same workload by the unmodified reference class "
                                          f"(oracle/_ref), mean {np.mean(ts[2:]):.3f}s each, NumPy {np.__version__}, "
                                          f"BLAS {blas_info()}"}
        line["parity"] = {
            "against": "the same system solved by the reference class on the host and by the CUDA path",
            "max_rel_deltaF": float(np.max(np.abs(gres.deltaF - rres.deltaF)) / np.max(np.abs(rres.deltaF))),
            "max_rel_Deltas": float(max(np.max(np.abs(Dg[i] - Dr[i])) / np.max(np.abs(Dr[i])) for i in range(4))),
            "flags_equal": bool(tuple(gres.truncatedTail) == tuple(rres.truncatedTail)
                                and tuple(gres.spectralPeaks) == tuple(rres.spectralPeaks)), "tolerance": 1e-10}
    print(json.dumps(line), flush=True)


def measure_fp64_peak(dev):
    """cuBLAS DGEMM through torch.matmul, 8192^3, best of 5: the denominator of the LU rooflines."""
    import torch

    x = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    y = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    torch.matmul(x, y)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(x, y)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del x, y
    torch.cuda.empty_cache()
    return 2.0 * 8192.0 ** 3 / (best * 1e-3) / 1e12


def run_b200(a):
    import torch
    import torch.distributed as dist

    import wallgo_b200 as wb
    from wallgo_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    scan_prep = None
    if a.scan_points > 0:
        try:
            scan_prep = prepare_scan(a, world)      # forks the scan's host workers before CUDA is initialised
        except Exception:
            import traceback
            traceback.print_exc()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    lib.wg_set_tuning(a.nb, a.graph)
    if a.bulk_hi >= 0:
        lib.wg_set_bulk_priority_rest(a.bulk_hi)
    P, M, N = WORKLOADS[a.workload]
    if a.workload == "cfg5stream":
        return run_stream(a, rank, world, local)
    if P * (M - 1) * (N - 1) ** 2 <= wb.BoltzmannSolver.STRIDED_MAX_N and a.batch > 0:
        run_small(a, rank, world, local, lib, measure_fp64_peak)
        if world > 1:
            dist.destroy_process_group()
        return
    K, W = a.steps, a.warmup
    nsys = K + W
    grid, particles, coll, bgs = build_inputs(P, M, N, nsys, rank)
    solver = wb.BoltzmannSolver(grid, device=local)
    solver.updateParticleList(particles)
    solver.setCollisionArray(coll)
    solver.setBackground(bgs[0])
    eng = solver._getEngine()
    n, nprof = eng.n, eng.nprof
    dev = eng.dev

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- measured FP64 GEMM peak (cuBLAS through torch) = denominator of the LU roofline
    x = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    y = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    torch.matmul(x, y)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(x, y)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fp64_peak = 2 * 8192.0**3 / best / 1e9  # TFLOP/s
    del x, y
    torch.cuda.empty_cache()

    # ---- device-resident inputs: every step's background already in HBM
    profs_h = np.empty((nsys, nprof))
    vws = []
    for i, bg in enumerate(bgs):
        solver.setBackground(bg)
        b = solver.background
        L = M + 1
        profs_h[i, :L] = b.temperatureProfile
        profs_h[i, L:2 * L] = b.velocityProfile
        profs_h[i, 2 * L:(2 + P) * L] = solver._msq(b.fieldProfiles).reshape(-1)
        profs_h[i, (2 + P) * L:] = grid.dxidchi
        vws.append(float(b.velocityWall))
    profs = torch.from_numpy(profs_h).to(dev)
    keep = (M - 1, N - 1, N - 1)
    if a.concurrency <= 0:
        a.concurrency = wb.BoltzmannSolver.autoConcurrency(n, torch.cuda.mem_get_info(dev)[0] + 8 * n * n)
    C_ = max(1, min(a.concurrency, K))
    sibs = solver._siblings(C_)
    engines = []
    for s_ in sibs:
        s_.setBackground(bgs[0])
        engines.append(s_._getEngine())
    streams = [torch.cuda.Stream(device=dev) for _ in range(C_)]
    own_prof = [e.prof for e in engines]

    # the library's own path (BoltzmannSolver._solveOnDevice): wg_factor + wg_solve
    fused_now = [False]

    def step(e, i, rec=None):
        e.prof = profs[i]
        if rec:
            rec[0].record()
        e.assemble(vws[i], 1.0)
        if rec:
            rec[1].record()
        if rec and len(rec) > 5:      # stage breakdown: factorisation and both triangular solves as separate calls
            e.factor()
            rec[2].record()
            e.solve()
            rec[5].record()
        elif fused_now[0]:             # alternative entry: forward solve fused into the factorisation
            e.factor_solve()
            if rec:
                rec[2].record()
        else:
            e.factor()
            e.solve()
            if rec:
                rec[2].record()
        if solver.refinementSteps:     # part of the product path (BoltzmannSolver._solveOnDevice)
            e.refine(vws[i], 1.0)
        if rec:
            rec[3].record()
        e.spectra_async(e.S)
        e.moments_async(e.S, keep, True)
        if rec:
            rec[4].record()

    # ---- warm-up (every engine), then a short solo pass for the per-stage breakdown
    for e in engines:
        for i in range(W):
            step(e, i)
    torch.cuda.synchronize()
    big = 8.0 * n * n > 40e9          # tens of seconds per step: one sample of each diagnostic pass is enough
    nsolo = 1 if big else min(3, K)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(nsolo)]
    import ctypes as C
    gms, gfl, gcnt = C.c_double(0), C.c_double(0), C.c_int(0)
    for i in range(nsolo):
        step(engines[0], W + i, evs[i])
    torch.cuda.synchronize()
    # one more solo factorisation on plain streams (events cannot sit inside the CUDA graph) with every
    # trailing-update GEMM launch bracketed by CUDA events on its launching stream
    lib.wg_set_tuning(a.nb, 0)
    lib.wg_set_bulk_split(0)     # one trailing-update launch per step, so a launch is timed on its own stream
    lib.wg_profile_gemm(engines[0].h, 1, None, None, None)
    step(engines[0], W)
    torch.cuda.synchronize()
    lib.wg_profile_gemm(engines[0].h, 0, C.byref(gms), C.byref(gfl), C.byref(gcnt))
    lib.wg_set_tuning(a.nb, a.graph)
    lib.wg_set_bulk_split(1)
    gemm_tf = gfl.value / (gms.value * 1e-3) / 1e12 if gms.value > 0 else float("nan")
    stage = np.array([[e[j].elapsed_time(e[j + 1]) for j in range(4)] for e in evs]).mean(axis=0)
    # the same with factorisation and triangular solves as separate calls (comparable with LAPACK getrf / getrs)
    evu = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    if not big:
        step(engines[0], W, evu)
    step(engines[0], W, evu)
    torch.cuda.synchronize()
    unfused = [evu[j].elapsed_time(evu[j + 1]) for j in range(4)]
    unfused[2] = evu[2].elapsed_time(evu[5])           # both triangular solves
    refine_ms = evu[5].elapsed_time(evu[3])            # residual (double-double, operator regenerated) + solve + update

    fused_now[0] = bool(a.fused)
    pm = 4 * P * (M - 1)
    gathered = torch.empty(world * pm, dtype=torch.float64, device=dev) if world > 1 else None
    sampler = ClockSampler(local)
    stagger = (2.0 / 3.0) * float(n) ** 3 / 20e12 / C_ if a.stagger < 0 else a.stagger
    barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    l0 = lib.wg_launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    start.record(cur)
    for st_ in streams:
        st_.wait_event(start)
    for i in range(K):
        slot = i % C_
        if i < C_ and i > 0:
            time.sleep(stagger)   # de-phase the slots: one LU's latency-bound tail under the other's GEMMs
        with torch.cuda.stream(streams[slot]):
            step(engines[slot], W + i)
    for st_ in streams:
        done = torch.cuda.Event()
        done.record(st_)
        cur.wait_event(done)
    if world > 1:  # the only collective on the path: gather the moments of the last scan point
        dist.all_gather_into_tensor(gathered, engines[(K - 1) % C_].out[n:n + pm].contiguous())
    end.record(cur)
    torch.cuda.synchronize()
    launches = lib.wg_launch_count() - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = start.elapsed_time(end)
    for e in engines:
        info = e.info()
        assert info == 0, f"singular operator in the benchmark (info={info})"
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * K / (ms / 1e3)

    # ---- end to end through the public class, host buffers in and out
    for e, pr in zip(engines, own_prof):
        e.prof = pr
    solver.getDeltasBatch(bgs[:(1 if big else min(W, 2 * C_))], concurrency=C_)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    allres = solver.getDeltasBatch(bgs[W:W + K], concurrency=C_)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    res = allres[-1]
    # sequential getDeltas() latency (one system at a time), for reference
    t0 = time.perf_counter()
    nseq = 1 if big else min(3, K)
    for i in range(nseq):
        solver.setBackground(bgs[W + i])
        solver.getDeltas()
    torch.cuda.synchronize()
    seq_ms = (time.perf_counter() - t0) / nseq * 1e3
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = world * K / e2e_s
    assert np.all(np.isfinite(res.Deltas.Delta00.coefficients))
    eng = engines[0]

    scan_line = None
    if a.scan_points > 0:
        try:
            scan_line = run_scan(a, world, rank, local, scan_prep)
        except Exception as exc:   # the secondary metric must not take the headline line down with it
            import traceback
            traceback.print_exc()
            scan_line = {"metric": "vw scan pts/sec", "error": repr(exc)[:300]}
    if rank == 0:
        hbm_peak, hbm_src = measured_peaks()
        lu_flops = 2.0 / 3.0 * float(n) ** 3
        lu_tf_solo = lu_flops / (unfused[1] * 1e-3) / 1e12
        lu_tf = K * lu_flops / (ms * 1e-3) / 1e12   # whole-GPU FP64 rate over the timed region (LU flops only)
        asm_gbs = (8.0 * n * n + 8.0 * n) / (stage[0] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{a.workload}: P={P} M={M} N={N} n={n} (BASELINE configs[1] when cfg2a), "
                                   "one independent system (scan point) per step per GPU",
                       "l2": "operator 8n^2 bytes >> 126 MB L2: inputs larger than L2, no flush needed"
                             if 8.0 * n * n > 4 * 126e6 else "small operator: L2-resident between stages",
                       "lu": {"nb": a.nb, "cuda_graph": bool(a.graph)},
                       "concurrency": f"{C_} independent systems in flight per GPU (staggered streams)",
                       "parallelism": f"{world} x independent systems, no collective inside a solve"},
            "stages_ms_solo": {"assemble": stage[0], "lu_factor_and_solves": stage[1],
                               "spectra_moments": stage[3], "lu_factor": unfused[1],
                               "tri_solve": unfused[2], "refinement_step": refine_ms,
                               "lu_tflops_solo": lu_tf_solo,
                               "note": "one system alone on the GPU, CUDA events around each stage"},
            "latency_ms_sequential_getDeltas": seq_ms,
            "roofline": {"kernel": "dgemm_sub_kernel<128,64,16,2,2,3> (trailing update of the LU, DMMA.8x8x4)",
                         "bound": "tensor", "achieved": gemm_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": gemm_tf / fp64_peak,
                         "traffic": profile_traffic("dgemm_128x64_2cta")[0] if (P, M, N) == (1, 40, 21) else None,
                         "traffic_source": profile_traffic("dgemm_128x64_2cta")[1],
                         "launches_timed": gcnt.value, "ms_timed": gms.value,
                         "algorithmic": "2*M*N*K flops of every trailing-update launch / its CUDA-event duration on the "
                                        "launching stream, summed over one solo factorisation on plain streams with the "
                                        "bulk columns unsplit (the look-ahead chain runs beside it); "
                                        "traffic = dram read+write of the first-step launch (15344^2 x 256) parsed from "
                                        "the ncu --set full summary under profiles/ (traffic_source) vs 3.83e9 algorithmic",
                         "peak_source": "measured in this run: torch.matmul (cuBLAS DGEMM) fp64 8192^3 best of 5; "
                                        "MEASURED_PEAKS.json has no FP64 entry"},
            "roofline_lu": {"kernel": "whole blocked LU (all kernels)", "bound": "tensor",
                         "achieved": lu_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": lu_tf / fp64_peak,
                         "traffic": None,
                         "peak_source": "measured in this run: torch.matmul (cuBLAS DGEMM) fp64 8192^3 best of 5; "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "algorithmic": "steps * 2/3 n^3 LU flops / CUDA-event duration of the whole timed region "
                                        "(assembly, solves and moments count as overhead)"},
            "roofline_assembly": {"kernel": "wg_assemble_kernel", "bound": "hbm", "achieved": asm_gbs, "peak": hbm_peak,
                                  "unit": "GB/s", "frac": asm_gbs / hbm_peak,
                                  "traffic": profile_traffic("assemble")[0] if (P, M, N) == (1, 40, 21) else None,
                                  "traffic_source": profile_traffic("assemble")[1], "peak_source": hbm_src,
                                  "algorithmic": "8n^2+8n bytes written per solve / mean CUDA-event time of wg_assemble"},
            "roofline_trisolve": {"kernel": "lu_trisolve_wave_kernel (forward + backward launch)", "bound": "hbm",
                                  "achieved": 8.0 * n * n / (unfused[2] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": 8.0 * n * n / (unfused[2] * 1e-3) / 1e9 / hbm_peak,
                                  "traffic": (2.0 * profile_traffic("trisolve_wave")[0]
                                              if (P, M, N) == (1, 40, 21) and profile_traffic("trisolve_wave")[0] else None),
                                  "traffic_source": profile_traffic("trisolve_wave")[1], "peak_source": hbm_src,
                                  "algorithmic": "8 n^2 bytes (every factor entry read once over the two launches) / "
                                                 "CUDA-event time of wg_solve, one system alone; traffic = 2 x the "
                                                 "per-launch dram bytes of the ncu summary"},
            "e2e": {"value": e2e, "unit": "solves/s", "h2d_bytes_per_step": int(nprof * 8),
                    "d2h_bytes_per_step": int((eng.nspec + eng.nout) * 8 + 4)},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if scan_line is not None:
            line["scan"] = scan_line
        if world == 1 and not a.no_cpu_baseline and 8.0 * n * n > 24e9:
            line["cpu_baseline"] = {"value": None, "unit": "solves/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "skipped: the NumPy path needs several copies of the operator in host RAM"}
        elif world == 1 and not a.no_cpu_baseline and reference_available():
            # the reference itself on the host cores, and the SAME system through the CUDA path: parity block
            try:
                import contextlib
                vmid = -0.3 - 0.4 * (W % 97) / 97.0
                with contextlib.redirect_stdout(sys.stderr):
                    t, rres, (rsolver, rgrid, rbg, rparts) = reference_one_solve(P, M, N, vmid, want_solver=True)
                    from oracle import ref_drivers as rd
                    from wallgo_b200 import dropin
                    new = dropin.makeDropInClass()(rgrid, rsolver.basisM, rsolver.basisN, rsolver.derivatives,
                                                   rsolver.collisionMultiplier, rsolver.truncationOption, device=local)
                    new.updateParticleList(rparts)
                    new.setBackground(rbg)
                    new.setCollisionArray(rsolver.collisionArray)
                    gres = new.getDeltas()
                    Dr, Dg = rd.deltas_table(rres), rd.deltas_table(gres)
                line["cpu_baseline"] = {
                    "value": 1.0 / t["total_s"], "unit": "solves/s", "cores": os.cpu_count(), "kind": "reference",
                    "sample": f"1 full solve of the same workload by the unmodified reference class (oracle/_ref): "
                              f"buildLinearEquations {t['assemble_s']:.2f}s + np.linalg.solve (LAPACK dgesv) "
                              f"{t['dgesv_s']:.2f}s + getDeltas(deltaF) {t['moments_s']:.2f}s, NumPy {np.__version__}, "
                              f"BLAS {blas_info()}"}
                line["parity"] = {
                    "against": "the same system (reference grid / background / collision array objects) solved by the "
                               "reference class on the host and by the CUDA path, full size",
                    "max_rel_deltaF": float(np.max(np.abs(gres.deltaF - rres.deltaF)) / np.max(np.abs(rres.deltaF))),
                    "max_rel_Deltas": float(max(np.max(np.abs(Dg[i] - Dr[i])) / np.max(np.abs(Dr[i])) for i in range(4))),
                    "flags_equal": bool(tuple(gres.truncatedTail) == tuple(rres.truncatedTail)
                                        and tuple(gres.spectralPeaks) == tuple(rres.spectralPeaks)),
                    "truncatedTail": [bool(x) for x in gres.truncatedTail],
                    "spectralPeaks": [int(x) for x in gres.spectralPeaks], "tolerance": 1e-10}
                del new, rsolver
            except MemoryError:
                line["cpu_baseline"] = {"value": None, "unit": "solves/s", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": "skipped: operator does not fit host RAM"}
        elif world == 1 and not a.no_cpu_baseline:
            try:
                t, ores = cpu_port_one_solve(grid, particles, coll, bgs[W])
                solver.setBackground(bgs[W])
                gres = solver.getDeltas()
                Dg = np.stack([gres.Deltas.Delta00.coefficients, gres.Deltas.Delta02.coefficients,
                               gres.Deltas.Delta20.coefficients, gres.Deltas.Delta11.coefficients])
                line["cpu_baseline"] = {
                    "value": 1.0 / t["total_s"], "unit": "solves/s", "cores": os.cpu_count(), "kind": "port",
                    "sample": f"1 full solve of the same workload (assemble {t['assemble_s']:.2f}s + LAPACK dgesv "
                              f"{t['dgesv_s']:.2f}s + moments {t['moments_s']:.2f}s), NumPy {np.__version__}, "
                              f"BLAS {blas_info()}"}
                line["parity"] = {
                    "against": "the oracle port on the same system (oracle/_ref absent on this machine)",
                    "max_rel_deltaF": float(np.max(np.abs(gres.deltaF - ores["deltaF"])) / np.max(np.abs(ores["deltaF"]))),
                    "max_rel_Deltas": float(max(np.max(np.abs(Dg[i] - ores["Deltas"][i])) / np.max(np.abs(ores["Deltas"][i]))
                                                for i in range(4))),
                    "flags_equal": bool(tuple(gres.truncatedTail) == tuple(ores["truncatedTail"])), "tolerance": 1e-10}
            except MemoryError:
                line["cpu_baseline"] = {"value": None, "unit": "solves/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": "skipped: operator does not fit host RAM"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2a", choices=sorted(WORKLOADS))
    ap.add_argument("--nb", type=int, default=256)
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--concurrency", type=int, default=0, help="independent systems in flight per GPU (0: auto)")
    ap.add_argument("--bulk-hi", type=int, default=-1, help="developer knob: wg_set_bulk_priority_rest")
    ap.add_argument("--stagger", type=float, default=-1.0, help="developer knob: seconds between slot starts")
    ap.add_argument("--fused", type=int, default=0, help="developer knob: 1 = wg_factor_solve, 0 = wg_factor + wg_solve (shipped path)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=64,
                    help="small systems (cfg1): systems per lock-step strided batch (0: the slot-pool path instead)")
    ap.add_argument("--inflight", type=int, default=2,
                    help="developer knob (small systems): strided batches in flight for the device-resident value")
    ap.add_argument("--panel-groups", type=int, default=6,
                    help="cfg5stream: (species, position) row groups per panel (6 x 900 rows x 226800 columns = 9.8 GB)")
    ap.add_argument("--scan-points", type=int, default=None,
                    help="vw-scan points of the secondary metric (default: 256 with the headline workload, 0 = skip otherwise)")
    ap.add_argument("--scan-workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--scan-model", default="idm", choices=sorted(SCAN_MODELS))
    ap.add_argument("--scan-workers", type=int, default=-1, help="host processes per rank for the scan (-1: auto, 0: in-process)")
    ap.add_argument("--scan-reuse", type=int, default=1,
                    help="scan server re-uses LU factors across the solves of a point (stale-factor refinement); 0: off")
    ap.add_argument("--scan-budget-s", type=float, default=200.0,
                    help="stop handing out scan points after this many seconds (the rate is reported over the points done)")
    ap.add_argument("--ref-scan-points", type=int, default=2, help="scan points the reference arm times (0: skip)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0)
    a = ap.parse_args()
    if a.scan_points is None:
        a.scan_points = 256 if a.workload == "cfg2a" else 0
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    if a.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host core, and BLAS reads the
        # variable when it is loaded, so restart this process once with the limit lifted.
        if os.environ.get("OMP_NUM_THREADS") and not os.environ.get("WG_REF_REEXEC"):
            env = dict(os.environ, WG_REF_REEXEC="1", OMP_NUM_THREADS=str(os.cpu_count()),
                       OPENBLAS_NUM_THREADS=str(os.cpu_count()))
            os.execve(sys.executable, [sys.executable] + sys.argv, env)
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
