import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from wallgo_b200 import _lib
lib = _lib.load()
lib.wg_debug_leaf.argtypes = [C.c_void_p, C.c_int]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15600
lib.wg_set_tuning(256, 0)
torch.manual_seed(0)
A0 = torch.randn(n, n, dtype=torch.float64, device="cuda")
ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
plan = C.c_void_p(); _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)))
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for c0 in (0, 16, 4096 + 16, n - 8192 + 32, n - 6000, n - 4096 + 16, n - 3000, n - 2048 + 16, n - 1024 + 16, n - 256 + 16):
    c0 -= c0 % 16
    dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
    lib.wg_debug_leaf(dbg.data_ptr(), c0)
    A = A0.clone()
    _lib.check(lib.wg_lu_factor(plan, A.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), st))
    torch.cuda.synchronize()
    d = dbg.cpu().numpy()
    cols = np.diff(np.concatenate([[d[22]], d[2:18]]))
    print(f"c0={c0:6d} m={n-c0:6d} load={d[22]-d[0]:6d} col0={cols[0]:6d} cols(mean 1..15)={cols[1:].mean():8.0f} min={cols[1:].min()} max={cols[1:].max()} "
          f"col8: argmax={d[25]-d[24]} slot={d[26]-d[25]} sync={d[27]-d[26]} push={d[28]-d[27]} wait={d[29]-d[28]} sel+div={d[30]-d[29]} ext={d[31]-d[30]} upd={d[10]-d[31]} | "
          f"loop_total={d[1]-d[0]} writeback={d[20]-d[1]} moves={d[21]-d[20]} total={d[21]-d[0]} cycles")
lib.wg_debug_leaf(None, -1)
