"""Developer probe: throughput of getDeltasStrided (lock-step strided batches, two in flight) vs batch size.
    python tools/strided_probe.py [P M N] [batches] [nItems]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import wallgo_b200 as wb
from wallgo_b200 import synthetic
P, M, N = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 20, 11)
batches = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [16, 24, 32, 37, 48, 64]
nItems = int(sys.argv[5]) if len(sys.argv) > 5 else 512
grid = synthetic.makeGrid(M, N); parts = synthetic.makeParticles(P)
s = wb.BoltzmannSolver(grid); s.updateParticleList(parts); s.setCollisionArray(synthetic.makeCollision(grid, parts))
bgs = [synthetic.makeBackground(grid, velocityMid=-0.3 - 0.4 * (i % 97) / 97.0) for i in range(nItems)]
for ref in (1, 0):
    s.refinementSteps = ref
    for B in batches:
        s.getDeltasStrided(bgs[:2 * B], batch=B); torch.cuda.synchronize()
        t0 = time.perf_counter(); out = s.getDeltasStrided(bgs, batch=B); torch.cuda.synchronize(); el = time.perf_counter() - t0
        print(f"P={P} M={M} N={N} refine={ref} batch={B}: {nItems / el:.0f} solves/s through getDeltasStrided (host buffers)", flush=True)
s.refinementSteps = 1
t0 = time.perf_counter(); out = s.getDeltasBatch(bgs, concurrency=32); torch.cuda.synchronize(); el = time.perf_counter() - t0
print(f"slot pool (32 streams x graphs): {nItems / el:.0f} solves/s", flush=True)
