"""Developer probe: how independent systems share one GPU (single-thread vs thread-per-slot enqueue, graphs)."""
import argparse, os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import wallgo_b200 as wb
from wallgo_b200 import _lib, synthetic
ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=1); ap.add_argument("--M", type=int, default=40); ap.add_argument("--N", type=int, default=21)
ap.add_argument("--slots", type=int, default=2); ap.add_argument("--per", type=int, default=6)
ap.add_argument("--graph", type=int, default=0); ap.add_argument("--threads", type=int, default=1)
ap.add_argument("--stagger", type=float, default=0.05)
a = ap.parse_args()
lib = _lib.load(); lib.wg_set_tuning(256, a.graph)
if os.environ.get("WG_SPLIT"): lib.wg_set_bulk_split(int(os.environ["WG_SPLIT"]))
grid = synthetic.makeGrid(a.M, a.N); parts = synthetic.makeParticles(a.P)
s = wb.BoltzmannSolver(grid); s.updateParticleList(parts)
s.setBackground(synthetic.makeBackground(grid)); s.setCollisionArray(synthetic.makeCollision(grid, parts))
sibs = s._siblings(a.slots)
engs = []
for x in sibs:
    x.setBackground(synthetic.makeBackground(grid)); e = x._getEngine(); x._uploadBackground(e); engs.append(e)
vw = float(s.background.velocityWall)
keep = (a.M - 1, a.N - 1, a.N - 1)
streams = [torch.cuda.Stream() for _ in engs]
def step(e):
    e.assemble(vw, 1.0); e.factor(); e.solve(); e.spectra_async(e.S); e.moments_async(e.S, keep, True)
for e, st in zip(engs, streams):
    with torch.cuda.stream(st):
        step(e); step(e)
torch.cuda.synchronize()
# (a) host enqueue time, empty queue
t0 = time.perf_counter()
with torch.cuda.stream(streams[0]):
    step(engs[0])
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"solo: host enqueue {1e3*(t1-t0):.1f} ms, until done {1e3*(t2-t0):.1f} ms (graph={a.graph})")
def worker(i, n):
    with torch.cuda.stream(streams[i]):
        for _ in range(n):
            step(engs[i])
torch.cuda.synchronize(); t0 = time.perf_counter()
if a.threads:
    th = []
    for i in range(a.slots):
        t = threading.Thread(target=worker, args=(i, a.per)); t.start(); th.append(t)
        if i + 1 < a.slots: time.sleep(a.stagger)
    for t in th: t.join()
else:
    for k in range(a.per):
        for i in range(a.slots):
            with torch.cuda.stream(streams[i]):
                step(engs[i])
            if k == 0 and i + 1 < a.slots: time.sleep(a.stagger)
te = time.perf_counter()
torch.cuda.synchronize(); t1 = time.perf_counter()
tot = a.slots * a.per
print(f"slots={a.slots} per={a.per} threads={a.threads} graph={a.graph}: {1e3*(t1-t0)/tot:.1f} ms/solve -> {tot/(t1-t0):.2f} solves/s (host enqueue finished after {1e3*(te-t0):.0f} of {1e3*(t1-t0):.0f} ms)")
for e in engs: assert e.info() == 0
