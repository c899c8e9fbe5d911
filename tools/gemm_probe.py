"""Runs the DMMA GEMM alone at the trailing-update shape of the first LU step (profiling target)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from wallgo_b200 import _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15600
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m = n - K
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib.wg_debug_gemm_variant(variant)
torch.manual_seed(0)
A = torch.randn(n, n, dtype=torch.float64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
Cp = A.data_ptr() + (K * n + K) * 8
Lp = A.data_ptr() + (K * n) * 8
Up = A.data_ptr() + K * 8
def run():
    _lib.check(lib.wg_dgemm_sub(Cp, Lp, Up, m, m, K, n, n, n, st))
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / reps
print(f"variant {variant} dgemm_sub M=N={m} K={K}: {t:.3f} ms  {2.0*m*m*K/t/1e9:.2f} TFLOP/s")
x = torch.randn(m, K, dtype=torch.float64, device="cuda"); y = torch.randn(K, m, dtype=torch.float64, device="cuda")
c = torch.randn(m, m, dtype=torch.float64, device="cuda")
torch.addmm(c, x, y, alpha=-1.0, out=c); torch.cuda.synchronize()
e0.record()
for _ in range(reps): torch.addmm(c, x, y, alpha=-1.0, out=c)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / reps
print(f"cuBLAS addmm same shape: {t:.3f} ms  {2.0*m*m*K/t/1e9:.2f} TFLOP/s")
