mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -x -q --timeout 300 -k "batched" > gpurun_out/t_batched.log 2>&1; tail -15 gpurun_out/t_batched.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "Strided" > gpurun_out/t_strided.log 2>&1; tail -15 gpurun_out/t_strided.log
timeout 300 python tools/batch_probe.py 1900 > gpurun_out/batch_probe.log 2>&1; cat gpurun_out/batch_probe.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/t_gpu_all.log 2>&1; tail -3 gpurun_out/t_gpu_all.log
timeout 500 python bench.py > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2c.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']); print(json.dumps(d.get('scan'))[:3000])"
