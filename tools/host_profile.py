"""Developer probe: where the host time of getDeltasBatch goes at a small system size."""
import cProfile, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import wallgo_b200 as wb
from wallgo_b200 import synthetic
P, M, N = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (1, 20, 11)))
grid = synthetic.makeGrid(M, N); parts = synthetic.makeParticles(P)
s = wb.BoltzmannSolver(grid); s.updateParticleList(parts); s.setCollisionArray(synthetic.makeCollision(grid, parts))
bgs = [synthetic.makeBackground(grid, velocityMid=-0.3 - 0.001 * i) for i in range(512)]
s.getDeltasBatch(bgs[:64])
torch.cuda.synchronize(); t0 = time.perf_counter()
s.getDeltasBatch(bgs)
torch.cuda.synchronize(); print(f"{len(bgs) / (time.perf_counter() - t0):.0f} solves/s through getDeltasBatch")
pr = cProfile.Profile(); pr.enable(); s.getDeltasBatch(bgs); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
