"""Developer probe: throughput of independent systems sharing one GPU as a function of the scheduling knobs
(slots in flight, panel width, 32-column leaves), through the public class with host buffers (getDeltasBatch),
i.e. what bench.py reports as `e2e`.  The slot table of BoltzmannSolver.autoConcurrency comes from sweeps like
this one (VERDICT round 1, weak point 11: tuned at four sizes only).
    python tools/knob_sweep.py P M N  slots=2,3,4  nb=256,128  leaf32=0,1  [split=1,0] [per=4]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import wallgo_b200 as wb
from wallgo_b200 import _lib, synthetic

P, M, N = (int(x) for x in sys.argv[1:4])
opts = dict(kv.split("=") for kv in sys.argv[4:])
slots = [int(x) for x in opts.get("slots", "0").split(",")]
nbs = [int(x) for x in opts.get("nb", "256").split(",")]
leafs = [int(x) for x in opts.get("leaf32", "0").split(",")]
per = int(opts.get("per", "4"))
splits = [int(x) for x in opts.get("split", "1").split(",")]     # wg_set_bulk_split: 0 off, 1 default (50 %), else percent
lib = _lib.load()
grid = synthetic.makeGrid(M, N); parts = synthetic.makeParticles(P)
s = wb.BoltzmannSolver(grid); s.updateParticleList(parts); s.setCollisionArray(synthetic.makeCollision(grid, parts))
n = P * (M - 1) * (N - 1) ** 2
small = n <= wb.BoltzmannSolver.STRIDED_MAX_N
ref = None
import itertools
for nb, l32, sp in itertools.product(nbs, leafs, splits):
    for c in slots:
        lib.wg_set_tuning(nb, 1); lib.wg_debug_leaf32(l32); lib.wg_set_bulk_split(sp)
        cc = None if (small or c <= 0) else c
        nsys = 256 if small else max(2, (c if c > 0 else 4) * per)
        bgs = [synthetic.makeBackground(grid, velocityMid=-0.3 - 0.001 * i) for i in range(nsys)]
        s.getDeltasBatch(bgs[:max(2, (128 if small else (c if c > 0 else 4)))], concurrency=cc)   # graphs captured
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = s.getDeltasBatch(bgs, concurrency=cc)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        d = np.concatenate([r.Deltas.Delta00.coefficients.ravel() for r in res])
        same = "" if ref is None else f"  Delta00 == first run: {bool(np.array_equal(d[:ref.size], ref[:d.size]))}"
        if ref is None: ref = d
        print(f"P={P} M={M} N={N} n={n} nb={nb} leaf32={l32} split={sp} slots={'auto' if cc is None else c}: "
              f"{nsys / dt:.2f} solves/s ({1e3 * dt / nsys:.2f} ms/solve, "
              f"{nsys * 2 / 3 * n ** 3 / dt / 1e12:.2f} TFLOP/s){same}", flush=True)
lib.wg_debug_leaf32(0); lib.wg_set_tuning(256, 1); lib.wg_set_bulk_split(1)
