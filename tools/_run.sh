timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -4
for c in 1 2 3; do timeout 300 python bench.py --steps 10 --warmup 3 --concurrency $c --no-cpu-baseline > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_c$c.json').read().strip().splitlines()[-1]);print($c, d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms_solo'], d['roofline']['frac'], d['roofline_lu']['frac'], d['gpu_launches'])"; tail -2 gpurun_out/bench_c$c.err; done
