"""Developer probe: where the time of the wavefront triangular solves goes (ablation knobs of the kernel).
    python tools/trisolve_probe.py [n]"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from wallgo_b200 import _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15600
lib.wg_set_tuning(256, 1)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
A = torch.randn(n, n, dtype=torch.float64, device="cuda")
b0 = torch.randn(n, dtype=torch.float64, device="cuda")
ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
plan = C.c_void_p(); _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)))
_lib.check(lib.wg_lu_factor(plan, A.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), st)); torch.cuda.synchronize()
x = b0.clone()
lib.wg_set_tuning(256, 0)      # plain launches: the knobs below are read at launch time
names = {0: "full", 1: "no diagonal step", 2: "no waiting", 3: "no diagonal step, no waiting",
         8: "no streaming products", 11: "nothing but tickets"}
for abl in (0, 1, 2, 3, 8, 11, 0):
    lib.wg_debug_wave_solve(1 | (abl << 4))
    ts = []
    for it in range(4):
        x.copy_(b0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.check(lib.wg_lu_solve(plan, A.data_ptr(), n, x.data_ptr(), st)); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"n={n} ablate={abl:2d} ({names[abl]}): {min(ts):.3f} ms", flush=True)
lib.wg_debug_wave_solve(0)
x.copy_(b0); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); _lib.check(lib.wg_lu_solve(plan, A.data_ptr(), n, x.data_ptr(), st)); e1.record(); torch.cuda.synchronize()
print(f"n={n} one launch per block (old form): {e0.elapsed_time(e1):.3f} ms")
lib.wg_debug_wave_solve(1); lib.wg_set_tuning(256, 1)
lib.wg_lu_destroy(plan)
