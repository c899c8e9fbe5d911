"""Developer probe: event timeline of the blocked LU (main stream vs look-ahead panel stream)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import wallgo_b200 as wb
from wallgo_b200 import _lib, synthetic
lib = _lib.load()
lib.wg_debug_lu_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
lib.wg_set_tuning(256, int(os.environ.get("WG_GRAPH", "0")))
P, M, N = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (1, 40, 21)))
grid = synthetic.makeGrid(M, N); parts = synthetic.makeParticles(P)
s = wb.BoltzmannSolver(grid); s.updateParticleList(parts)
s.setBackground(synthetic.makeBackground(grid)); s.setCollisionArray(synthetic.makeCollision(grid, parts))
eng = s._getEngine(); s._uploadBackground(eng)
vw = float(s.background.velocityWall)
n = eng.n
A = eng.assemble(vw, 1.0).clone()
ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
plan = C.c_void_p(); _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)))
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for rep in range(3):
    W = A.clone()
    torch.cuda.synchronize()
    lib.wg_debug_lu_trace(plan, 1 if rep == 2 else 0, None, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.wg_lu_factor(plan, W.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), st))
    e1.record(); torch.cuda.synchronize()
    print(f"rep {rep}: factor {e0.elapsed_time(e1):.2f} ms")
out = np.zeros(5 * 256)
ns = lib.wg_debug_lu_trace(plan, 0, out.ctypes.data, 256)
t = out[:5 * ns].reshape(ns, 5)
print("step rest  start   bulk-perm  wideTRSM  wideGEMM  chain(perm+narrow+panel k+1)  wait_for_chain")
tot = np.zeros(5)
for i in range(ns - 1):
    s0, fk, tr, ge, pn = t[i]
    nxt = t[i + 1][0]
    if fk < 0:
        continue
    row = (fk - s0, tr - fk, ge - tr, pn - s0, max(0.0, pn - ge))
    tot += np.array(row)
    if i % 4 == 0 or i > ns - 6:
        print(f"{i:3d} {n - 256 * (i + 1):6d} {s0:8.2f} {row[0]:8.3f} {row[1]:8.3f} {row[2]:8.3f} {row[3]:8.3f} {row[4]:8.3f}")
print("totals: bulk-perm %.2f wideTRSM %.2f wideGEMM %.2f chain %.2f exposed-chain-wait %.2f ms" % tuple(tot))
