mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -x -q --timeout 600 > gpurun_out/t_dense.log 2>&1; tail -4 gpurun_out/t_dense.log
timeout 300 python tools/lu_fuzz.py 150 3 > gpurun_out/fuzz_lu.log 2>&1; tail -1 gpurun_out/fuzz_lu.log
timeout 200 python tools/batch_probe.py 1900 64 > gpurun_out/batch_probe2.log 2>&1; cat gpurun_out/batch_probe2.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refparity.py -m gpu -x -q --timeout 300 > gpurun_out/t_par.log 2>&1; tail -3 gpurun_out/t_par.log
timeout 500 python bench.py --scan-points 0 --no-cpu-baseline > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2h.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['stages_ms_solo'], d['roofline']['frac'], d['roofline_lu']['frac'], d['roofline_trisolve']['frac'])"
