mkdir -p gpurun_out
timeout 1000 python bench.py --workload cfg5fit --steps 2 > gpurun_out/bench_cfg5fit.json 2> gpurun_out/bench_cfg5fit.err; tail -c 400 gpurun_out/bench_cfg5fit.json; tail -3 gpurun_out/bench_cfg5fit.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_cfg5fit.json').read().strip().splitlines()[-1]); print(d['value'], d['stages_ms_solo'], d['roofline']['frac'], d['roofline_lu']['frac'], d['roofline_trisolve']['frac'])"
