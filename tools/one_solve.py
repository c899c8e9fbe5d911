"""One assemble+factor+solve+moments at a given size (profiling target)."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wallgo_b200 as wb
from wallgo_b200 import _lib, synthetic
ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=1); ap.add_argument("--M", type=int, default=40)
ap.add_argument("--N", type=int, default=21); ap.add_argument("--nb", type=int, default=256)
ap.add_argument("--graph", type=int, default=0); ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
_lib.load().wg_set_tuning(a.nb, a.graph)
grid = synthetic.makeGrid(a.M, a.N); parts = synthetic.makeParticles(a.P)
s = wb.BoltzmannSolver(grid); s.updateParticleList(parts)
s.setBackground(synthetic.makeBackground(grid)); s.setCollisionArray(synthetic.makeCollision(grid, parts))
for _ in range(a.reps):
    r = s.getDeltas()
torch.cuda.synchronize()
print("ok", float(abs(r.Deltas.Delta00.coefficients).max()))
