"""Sharding of independent Boltzmann systems (wall-velocity scan points, parameter points) over
one process per GPU.  A single system never leaves its GPU; the only collective is one gather of
the per-point results.  Works with NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

import numpy as np


def shardPoints(nPoints: int, rank: int, worldSize: int) -> list[int]:
    """Round-robin: rank r takes points i = r (mod G); every point starts from the same initial
    guess so 1- and 8-GPU results are identical."""
    assert 0 <= rank < worldSize
    return list(range(rank, nPoints, worldSize))


def gatherPoints(localValues: np.ndarray, nPoints: int, width: int, device=None) -> np.ndarray:
    """All ranks receive the full [nPoints, width] table (one all_gather)."""
    import torch
    import torch.distributed as dist

    localValues = np.asarray(localValues, dtype=np.float64).reshape(-1, width)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert localValues.shape[0] == nPoints
        return localValues
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (nPoints + world - 1) // world
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    buf = torch.zeros((per, width), dtype=torch.float64, device=dev)
    mine = shardPoints(nPoints, rank, world)
    assert localValues.shape[0] == len(mine)
    if len(mine):
        buf[:len(mine)] = torch.from_numpy(localValues).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    full = np.empty((nPoints, width))
    for r in range(world):
        idx = shardPoints(nPoints, r, world)
        full[idx] = out[r][:len(idx)].cpu().numpy()
    return full


def scanDeltas(makeSolver, backgrounds, device=None, concurrency=None) -> np.ndarray:
    """Solve one Boltzmann system per background on this rank's share (several systems in flight on the
    rank's GPU, see BoltzmannSolver.getDeltasBatch) and gather the [nPoints, 4*P*(M-1)] moment tables.
    `makeSolver()` builds this rank's own solver."""
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    solver = makeSolver()
    mine = shardPoints(len(backgrounds), rank, world)
    if hasattr(solver, "getDeltasBatch"):
        results = solver.getDeltasBatch([backgrounds[i] for i in mine], concurrency=concurrency)
    else:
        results = []
        for i in mine:
            solver.setBackground(backgrounds[i])
            results.append(solver.getDeltas())
    rows = []
    for res in results:
        d = res.Deltas
        rows.append(np.concatenate([d.Delta00.coefficients.ravel(), d.Delta02.coefficients.ravel(),
                                    d.Delta20.coefficients.ravel(), d.Delta11.coefficients.ravel()]))
    width = rows[0].size if rows else 0
    if world > 1:
        import torch
        w = torch.tensor([width], dtype=torch.int64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        width = int(w.item())
    local = np.array(rows).reshape(-1, width) if rows else np.empty((0, width))
    return gatherPoints(local, len(backgrounds), width, device)
