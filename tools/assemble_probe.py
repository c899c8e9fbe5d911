"""Runs the assembly kernel alone at the benchmark size (profiling target)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import wallgo_b200 as wb
from wallgo_b200 import synthetic
P, M, N = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (1, 40, 21)))
grid = synthetic.makeGrid(M, N); parts = synthetic.makeParticles(P)
s = wb.BoltzmannSolver(grid); s.updateParticleList(parts)
s.setBackground(synthetic.makeBackground(grid)); s.setCollisionArray(synthetic.makeCollision(grid, parts))
eng = s._getEngine(); s._uploadBackground(eng)
vw = float(s.background.velocityWall)
from wallgo_b200 import _lib
if os.environ.get("WG_ASM_WIDE"): _lib.load().wg_debug_asm_wide(int(os.environ["WG_ASM_WIDE"]))
for _ in range(3): eng.assemble(vw, 1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): eng.assemble(vw, 1.0)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 5
print(f"assemble n={eng.n}: {t:.4f} ms {8.0*eng.n*eng.n/t/1e6:.0f} GB/s")
