mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -x -q --timeout 300 -k "trisolve or batched or lu_random or fused" > gpurun_out/t_tri.log 2>&1; tail -4 gpurun_out/t_tri.log
timeout 200 python tools/batch_probe.py 15600 1 > gpurun_out/batch_probe3.log 2>&1; cat gpurun_out/batch_probe3.log
timeout 200 python tools/batch_probe.py 1900 64 > gpurun_out/batch_probe2.log 2>&1; cat gpurun_out/batch_probe2.log
timeout 500 python bench.py --scan-points 0 --no-cpu-baseline > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2f.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['stages_ms_solo'])"
