import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from wallgo_b200 import _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15600
torch.manual_seed(0)
A = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, dtype=torch.float64, device="cuda")
ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
plan = C.c_void_p(); _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)))
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
_lib.check(lib.wg_lu_factor(plan, A.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    e0.record()
    _lib.check(lib.wg_lu_solve(plan, A.data_ptr(), n, b.data_ptr(), st))
    e1.record(); torch.cuda.synchronize()
    print("solve ms", e0.elapsed_time(e1))
