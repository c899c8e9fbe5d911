"""Developer probe: unphysical (NaN-producing) inputs must not fault."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import wallgo_b200 as wb
from wallgo_b200 import synthetic, _lib
lib = _lib.load(); lib.wg_set_tuning(256, int(os.environ.get("WG_GRAPH", "0")))
grid = synthetic.makeGrid(20, 11); parts = synthetic.makeParticles(1)
s = wb.BoltzmannSolver(grid); s.updateParticleList(parts); s.setCollisionArray(synthetic.makeCollision(grid, parts))
vm = float(sys.argv[1]) if len(sys.argv) > 1 else -1.02
with np.errstate(all="ignore"):
    s.setBackground(synthetic.makeBackground(grid, velocityMid=vm))
eng = s._getEngine(); s._uploadBackground(eng)
vw = float(s.background.velocityWall)
def stage(name, fn):
    fn(); torch.cuda.synchronize(); print(name, "ok", flush=True)
stage("assemble", lambda: eng.assemble(vw, 1.0))
print("nan in A:", bool(torch.isnan(eng.A).any()), "nan in S:", bool(torch.isnan(eng.S).any()), flush=True)
stage("factor", lambda: eng.factor())
print("info", eng.info(), flush=True)
stage("solve", lambda: eng.solve())
stage("spectra", lambda: eng.spectra_async(eng.S))
