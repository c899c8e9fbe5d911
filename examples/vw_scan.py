"""Wall-velocity scan of independent Boltzmann systems over 1..8 GPUs.

    python examples/vw_scan.py                                      # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/vw_scan.py

Each rank builds its own solver (one process per GPU), takes the scan points i = rank (mod world), keeps several
systems in flight on its GPU (BoltzmannSolver.getDeltasBatch) and one NCCL all_gather returns the table of
moments [nPoints, 4*P*(M-1)] to every rank.  Inputs are the synthetic ones of bench.py;
with WallGo installed the same call works on a `wallgo_b200.dropin` solver and real BoltzmannBackgrounds.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=4)
    ap.add_argument("--M", type=int, default=30)
    ap.add_argument("--N", type=int, default=11)
    ap.add_argument("--points", type=int, default=256)
    a = ap.parse_args()

    import torch
    import torch.distributed as dist

    import wallgo_b200 as wb
    from wallgo_b200 import scan, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    grid = synthetic.makeGrid(a.M, a.N)
    particles = synthetic.makeParticles(a.P)
    solver = wb.BoltzmannSolver(grid, device=local)
    solver.updateParticleList(particles)
    solver.setCollisionArray(synthetic.makeCollision(grid, particles))
    vws = np.linspace(0.05, 0.53, a.points)
    backgrounds = [synthetic.makeBackground(grid, velocityMid=-vw) for vw in vws]

    scan.scanDeltas(lambda: solver, backgrounds[:4 * world])        # warm-up: buffers, CUDA graphs
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    table = scan.scanDeltas(lambda: solver, backgrounds)            # [points, 4*P*(M-1)] on every rank
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        d00 = table[:, :a.P * (a.M - 1)].reshape(a.points, a.P, a.M - 1)
        print(f"{a.points} points, n = {a.P * (a.M - 1) * (a.N - 1) ** 2}, {world} GPU(s): {dt:.2f} s "
              f"({a.points / dt:.1f} points/s)")
        for i in (0, a.points // 2, a.points - 1):
            print(f"  vw = {vws[i]:.3f}: max |Delta00| per species = {np.abs(d00[i]).max(axis=1)}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
