mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_e2e.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 > gpurun_out/t_scan.log 2>&1; tail -3 gpurun_out/t_scan.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value']); print(json.dumps(d.get('scan'))[:1800])"
