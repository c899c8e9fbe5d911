"""Developer probe: solo LU factorisation time with 16- and 32-column leaf panels (wg_debug_leaf32).
    python tools/leaf32_probe.py [n]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from wallgo_b200 import _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15600
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
A0 = torch.randn(n, n, dtype=torch.float64, device="cuda")
A = A0.clone()
ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
ref = None
for leaf32 in (0, 1, 0, 1):
    lib.wg_debug_leaf32(leaf32)
    plan = C.c_void_p(); _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)))
    ts = []
    for it in range(3):
        A.copy_(A0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.check(lib.wg_lu_factor(plan, A.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), st)); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    piv = ipiv.clone()
    same = "" if ref is None else f" pivots equal to the first run: {bool(torch.equal(piv, ref))}"
    if ref is None: ref = piv
    print(f"n={n} leaf32={leaf32}: factor {min(ts[1:]):.2f} ms ({2 / 3 * n ** 3 / min(ts[1:]) / 1e9:.2f} TFLOP/s){same}", flush=True)
    lib.wg_lu_destroy(plan)
lib.wg_debug_leaf32(0)
