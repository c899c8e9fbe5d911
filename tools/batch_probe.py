"""Developer probe: lock-step strided-batch LU (wg_lu_*_batched) timing vs batch size.
    python tools/batch_probe.py [n] [B1,B2,...]"""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from wallgo_b200 import _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1900
Bs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 8, 16, 32, 37, 64, 74, 128]
graph = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lib.wg_set_tuning(256, graph)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for B in Bs:
    torch.manual_seed(0)
    A0 = torch.randn(B, n, n, dtype=torch.float64, device="cuda")
    A = A0.clone()
    b = torch.randn(B, n, dtype=torch.float64, device="cuda")
    ipiv = torch.zeros(B * n, dtype=torch.int32, device="cuda"); info = torch.zeros(B, dtype=torch.int32, device="cuda")
    plan = C.c_void_p(); _lib.check(lib.wg_lu_create_batched(n, B, 0, C.byref(plan)))
    def run():
        _lib.check(lib.wg_lu_factor_batched(plan, A.data_ptr(), n, n * n, ipiv.data_ptr(), info.data_ptr(), st))
    A.copy_(A0); run(); torch.cuda.synchronize()
    ts = []
    for it in range(int(os.environ.get('BATCH_PROBE_ITERS', '4'))):
        A.copy_(A0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    x = b.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.check(lib.wg_lu_solve_batched(plan, A.data_ptr(), n, n * n, x.data_ptr(), n, st)); torch.cuda.synchronize()
    x = b.clone(); e0.record()
    _lib.check(lib.wg_lu_solve_batched(plan, A.data_ptr(), n, n * n, x.data_ptr(), n, st)); e1.record(); torch.cuda.synchronize()
    tsolve = e0.elapsed_time(e1)
    res = (torch.bmm(A0, x.unsqueeze(2)).squeeze(2) - b).abs().max().item()
    t = min(ts) if ts else float('nan')
    print(f"n={n} B={B} factor {t:.3f} ms  -> {B / t * 1e3:.0f} systems/s, {B * 2 / 3 * n ** 3 / t / 1e9:.2f} TFLOP/s; solve {tsolve:.3f} ms; "
          f"residual {res:.2e} info {int(info.abs().max())}", flush=True)
    lib.wg_lu_destroy(plan)
    del A, A0, b, x
