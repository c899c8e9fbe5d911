# developer aid: compute-sanitizer over the round-2 kernels at small sizes (memcheck + racecheck)
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import wallgo_b200 as wb
from wallgo_b200 import synthetic
for (P, M, N) in ((2, 9, 7), (1, 14, 9)):
    grid = synthetic.makeGrid(M, N); parts = synthetic.makeParticles(P)
    s = wb.BoltzmannSolver(grid); s.updateParticleList(parts); s.setCollisionArray(synthetic.makeCollision(grid, parts))
    bgs = [synthetic.makeBackground(grid, velocityMid=-0.3 - 0.05 * i) for i in range(5)]
    s.setBackground(bgs[0]); r = s.getDeltas()
    out = s.getDeltasStrided(bgs, batch=3)
    out2 = s.getDeltasBatch(bgs, concurrency=2)
    assert np.array_equal(out[0].deltaF, r.deltaF) and np.array_equal(out2[0].deltaF, r.deltaF)
    s.checkLinearization()
torch.cuda.synchronize(); print("ok")
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_small.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san_small.py > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/san_racecheck.log
