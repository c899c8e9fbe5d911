"""Developer probe: create/use/destroy solvers and LU plans repeatedly; device memory must come back."""
import ctypes as C, gc, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import wallgo_b200 as wb
from wallgo_b200 import _lib, synthetic
lib = _lib.load()
torch.cuda.init(); torch.zeros(1, device="cuda")
def free(): torch.cuda.synchronize(); torch.cuda.empty_cache(); return torch.cuda.mem_get_info()[0]
f0 = free()
for it in range(60):
    grid = synthetic.makeGrid(12, 7); parts = synthetic.makeParticles(2)
    s = wb.BoltzmannSolver(grid); s.updateParticleList(parts); s.setCollisionArray(synthetic.makeCollision(grid, parts))
    s.setBackground(synthetic.makeBackground(grid)); s.getDeltas(); s.getDeltasBatch([synthetic.makeBackground(grid, velocityMid=-0.3 - 0.01 * i) for i in range(6)])
    for x in (getattr(s, "_sibs", None) or [s]):
        if x._engine is not None: x._engine.close()
    del s; gc.collect()
    if it in (0, 9, 59): print(f"solver iteration {it}: free {free() / 2**20:.0f} MiB (start {f0 / 2**20:.0f})", flush=True)
    if it == 9: f9 = free()      # after the one-time costs (module load, graph memory pools)
n = 3000
A = torch.randn(n, n, dtype=torch.float64, device="cuda"); ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
f1 = free()
for it in range(40):
    plan = C.c_void_p(); _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)))
    W = A.clone()
    _lib.check(lib.wg_lu_factor(plan, W.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), st)); torch.cuda.synchronize()
    lib.wg_lu_destroy(plan); del W
    if it in (0, 39): print(f"plan iteration {it}: free {free() / 2**20:.0f} MiB (start {f1 / 2**20:.0f})", flush=True)
del A
leak = (f9 - free()) / 2**20
print(f"change since solver iteration 9: {leak:.0f} MiB (one-time costs before it: {(f0 - f9) / 2**20:.0f} MiB)")
sys.exit(1 if leak > 50 else 0)
