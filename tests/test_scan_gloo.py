"""Multi-process path (one process per GPU in production, NCCL): sharding of independent systems
and the single result gather, exercised here with world_size=2 on the gloo backend."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from wallgo_b200 import scan

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nPoints, width = 7, 5
        mine = scan.shardPoints(nPoints, rank, world)
        local = np.array([[100.0 * i + c for c in range(width)] for i in mine])
        full = scan.gatherPoints(local, nPoints, width)
        q.put((rank, mine, full))
    finally:
        dist.destroy_process_group()


def test_shard_points_cover_everything_once():
    from wallgo_b200 import scan

    for n in (1, 7, 256):
        for w in (1, 2, 4, 8):
            allp = sorted(sum((scan.shardPoints(n, r, w) for r in range(w)), []))
            assert allp == list(range(n))
    assert scan.shardPoints(256, 3, 8)[:3] == [3, 11, 19]


def test_gather_world_size_2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.array([[100.0 * i + c for c in range(5)] for i in range(7)])
    for rank, mine, full in out:
        assert mine == list(range(rank, 7, 2))
        assert np.array_equal(full, expect)


def test_gather_single_process_is_identity():
    from wallgo_b200 import scan

    x = np.arange(12.0).reshape(4, 3)
    assert np.array_equal(scan.gatherPoints(x, 4, 3), x)


class _FakeBatchSolver:
    """Stands in for BoltzmannSolver in the sharding logic: four 'moments' that identify the scan point."""

    class _Poly:
        def __init__(self, arr):
            self.coefficients = arr

    class _Res:
        pass

    def getDeltasBatch(self, backgrounds, concurrency=None):
        out = []
        for bg in backgrounds:
            r = self._Res()
            r.Deltas = self._Res()
            base = np.full((2, 3), float(bg))
            r.Deltas.Delta00 = self._Poly(base + 0.0)
            r.Deltas.Delta02 = self._Poly(base + 0.25)
            r.Deltas.Delta20 = self._Poly(base + 0.5)
            r.Deltas.Delta11 = self._Poly(base + 0.75)
            out.append(r)
        return out


def _scan_worker(rank, world, port, q):
    import torch.distributed as dist
    from wallgo_b200 import scan

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        table = scan.scanDeltas(lambda: _FakeBatchSolver(), list(range(9)), device="cpu")
        q.put((rank, table))
    finally:
        dist.destroy_process_group()


def test_scanDeltas_world_size_2_gloo():
    """scanDeltas end to end on two ranks: each rank batches its own share, every rank receives the full table
    in input order."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scan_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.array([np.concatenate([np.full(6, i + o) for o in (0.0, 0.25, 0.5, 0.75)]) for i in range(9)])
    for _, table in out:
        assert table.shape == (9, 24)
        assert np.array_equal(table, expect)
    # single process: same table without a process group
    from wallgo_b200 import scan
    assert np.array_equal(scan.scanDeltas(lambda: _FakeBatchSolver(), list(range(9))), expect)
