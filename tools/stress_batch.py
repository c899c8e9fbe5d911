"""Developer stress: batched (several systems in flight) vs sequential results must be bit-identical, repeatedly."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import wallgo_b200 as wb
from wallgo_b200 import synthetic
P, M, N, reps, npts = (int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (1, 40, 21, 3, 9)))
grid = synthetic.makeGrid(M, N); parts = synthetic.makeParticles(P)
s = wb.BoltzmannSolver(grid); s.updateParticleList(parts); s.setCollisionArray(synthetic.makeCollision(grid, parts))
bgs = [synthetic.makeBackground(grid, velocityMid=-0.3 - 0.02 * i) for i in range(npts)]
seq = []
for bg in bgs:
    s.setBackground(bg); seq.append(s.getDeltas())
bad = 0
for r in range(reps):
    out = s.getDeltasBatch(bgs)
    for i, (a, b) in enumerate(zip(seq, out)):
        if not (np.array_equal(a.deltaF, b.deltaF) and np.array_equal(a.Deltas.Delta02.coefficients, b.Deltas.Delta02.coefficients)):
            bad += 1; print("MISMATCH rep", r, "item", i, float(np.max(np.abs(a.deltaF - b.deltaF))))
    # and sequential again (graph replay determinism)
    s.setBackground(bgs[r % npts]); again = s.getDeltas()
    if not np.array_equal(again.deltaF, seq[r % npts].deltaF):
        bad += 1; print("MISMATCH sequential replay rep", r)
print(f"P={P} M={M} N={N}: {reps} reps x {npts} points, mismatches: {bad}")
sys.exit(1 if bad else 0)
