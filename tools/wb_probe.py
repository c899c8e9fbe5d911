"""Developer probe: the LU with a forced leaf width class against LAPACK at small sizes."""
import ctypes as C, os, sys
import numpy as np, scipy.linalg, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from wallgo_b200 import _lib
lib = _lib.load()
wb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lib.wg_debug_force_wb(wb)
for n in (4, 8, 30, 100, 256, 300, 700, 1500, 4200):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    Ad = torch.from_numpy(A).cuda(); ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    plan = C.c_void_p(); _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.wg_lu_factor(plan, Ad.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), st)); torch.cuda.synchronize()
    lib.wg_lu_destroy(plan)
    LU = Ad.cpu().numpy(); p = ipiv.cpu().numpy()
    LUr, pr = scipy.linalg.lu_factor(A)
    print(f"wb={wb} n={n}: same pivots {np.array_equal(p, pr)} first diff {int(np.argmax(p != pr)) if not np.array_equal(p, pr) else -1} max|LU-ref| {np.max(np.abs(LU - LUr)):.2e}", flush=True)
