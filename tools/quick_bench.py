"""Developer timing probe (not the contract bench): per-stage CUDA-event timings."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wallgo_b200 as wb  # noqa: E402
from wallgo_b200 import _lib, synthetic  # noqa: E402


def ev():
    return torch.cuda.Event(enable_timing=True)


def time_stage(fn, reps):
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=1)
    ap.add_argument("--M", type=int, default=40)
    ap.add_argument("--N", type=int, default=21)
    ap.add_argument("--nb", type=int, default=256)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--dgemm", type=int, default=1)
    ap.add_argument("--look", type=int, default=1)
    a = ap.parse_args()
    lib = _lib.load()
    lib.wg_set_tuning(a.nb, a.graph)
    lib.wg_set_lookahead(a.look)
    if a.dgemm:
        x = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        y = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        torch.matmul(x, y)
        t = min(time_stage(lambda: torch.matmul(x, y), 1) for _ in range(5))
        print(f"cuBLAS DGEMM 8192^3: {t:.2f} ms  {2 * 8192**3 / t / 1e9:.1f} TFLOP/s")
        del x, y
    grid = synthetic.makeGrid(a.M, a.N)
    parts = synthetic.makeParticles(a.P)
    s = wb.BoltzmannSolver(grid)
    s.updateParticleList(parts)
    s.setBackground(synthetic.makeBackground(grid))
    t0 = time.perf_counter()
    s.setCollisionArray(synthetic.makeCollision(grid, parts))
    eng = s._getEngine()
    s._uploadBackground(eng)
    n = eng.n
    print(f"P={a.P} M={a.M} N={a.N} n={n} nb={a.nb} graph={a.graph} look={a.look} setup {time.perf_counter() - t0:.2f}s")
    vw = float(s.background.velocityWall)
    eng.assemble(vw, 1.0)
    t_as = time_stage(lambda: eng.assemble(vw, 1.0), a.reps)
    print(f"assemble: {t_as:.3f} ms  -> {8.0 * n * n / t_as / 1e6:.0f} GB/s")
    A0 = eng.A.clone() if n <= 20000 else None
    S0 = eng.S.clone()

    def fac():
        eng.assemble(vw, 1.0)
        eng.factor()
    fac()
    t_f = time_stage(fac, a.reps) - t_as
    print(f"factor: {t_f:.2f} ms  -> {2.0 / 3.0 * n**3 / t_f / 1e9:.2f} TFLOP/s")
    eng.S.copy_(S0)
    t_s = time_stage(lambda: eng.solve(), a.reps)
    print(f"solve: {t_s:.3f} ms")
    eng.assemble(vw, 1.0)
    eng.factor()
    eng.solve()
    print("info", eng.info())
    if A0 is not None:
        r = torch.mv(A0, eng.S) - S0
        print("residual", float(torch.linalg.norm(r) / torch.linalg.norm(S0)))
    keep = (a.M - 1, a.N - 1, a.N - 1)

    def mom():
        eng.spectra_async(eng.S)
        eng.moments_async(eng.S, keep, True)
    mom()
    t_m = time_stage(mom, a.reps)
    print(f"spectra+moments: {t_m:.3f} ms")
    tot = t_as + t_f + t_s + t_m
    print(f"total {tot:.2f} ms -> {1000.0 / tot:.2f} solves/s; launches so far {lib.wg_launch_count()}")
    t0 = time.perf_counter()
    for _ in range(a.reps):
        s.getDeltas()
    torch.cuda.synchronize()
    print(f"e2e getDeltas: {(time.perf_counter() - t0) / a.reps * 1e3:.2f} ms")


if __name__ == "__main__":
    main()
