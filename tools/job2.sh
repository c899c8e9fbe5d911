mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -x -q --timeout 300 -k "trisolve or batched or lu_random or schedules" > gpurun_out/t_tri.log 2>&1; tail -15 gpurun_out/t_tri.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "refinement or Strided or getDeltasBatch" -s > gpurun_out/t_refine.log 2>&1; grep -E "cond~|passed|failed|Error|assert" gpurun_out/t_refine.log | head -30
timeout 300 python tools/batch_probe.py 1900 1,32,64 > gpurun_out/batch_probe2.log 2>&1; cat gpurun_out/batch_probe2.log
timeout 300 python tools/batch_probe.py 15600 1 > gpurun_out/batch_probe3.log 2>&1; cat gpurun_out/batch_probe3.log
timeout 600 python tools/strided_probe.py > gpurun_out/strided_probe.log 2>&1; cat gpurun_out/strided_probe.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/t_gpu_all.log 2>&1; tail -3 gpurun_out/t_gpu_all.log
timeout 500 python bench.py --scan-points 0 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2d.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['stages_ms_solo'], d.get('parity'))"
