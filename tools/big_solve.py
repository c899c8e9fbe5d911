"""Developer probe: stage timings of one very large system (default: P=4, M=64, N=21, n=100800, 81 GB)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import wallgo_b200 as wb
from wallgo_b200 import synthetic
P, M, N = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (4, 64, 21)))
grid = synthetic.makeGrid(M, N); parts = synthetic.makeParticles(P)
s = wb.BoltzmannSolver(grid); s.updateParticleList(parts); s.setCollisionArray(synthetic.makeCollision(grid, parts))
s.setBackground(synthetic.makeBackground(grid))
eng = s._getEngine(); s._uploadBackground(eng)
vw = float(s.background.velocityWall); n = eng.n
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
eng.assemble(vw, 1.0); torch.cuda.synchronize()
ev[0].record(); eng.assemble(vw, 1.0); ev[1].record(); eng.factor(); ev[2].record(); eng.solve(); ev[3].record(); torch.cuda.synchronize()
ta, tf, ts = (ev[i].elapsed_time(ev[i + 1]) for i in range(3))
print(f"n={n} operator {8.0*n*n/1e9:.1f} GB: assemble {ta:.2f} ms ({8.0*n*n/ta/1e6:.0f} GB/s), factor {tf/1e3:.2f} s ({2/3*n**3/tf/1e9:.1f} TFLOP/s, first run incl. graph capture), solves {ts:.1f} ms, info {eng.info()}")
eng.assemble(vw, 1.0); torch.cuda.synchronize()
ev[1].record(); eng.factor(); ev[2].record(); torch.cuda.synchronize()
tf = ev[1].elapsed_time(ev[2]); print(f"factor again (graph replay): {tf/1e3:.2f} s ({2/3*n**3/tf/1e9:.1f} TFLOP/s)")
