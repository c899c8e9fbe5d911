mkdir -p gpurun_out
timeout 400 python tools/lu_fuzz.py 44 5 big > gpurun_out/fuzz_lu_big.log 2>&1; tail -1 gpurun_out/fuzz_lu_big.log
timeout 300 python tools/stress_batch.py > gpurun_out/stress_batch.log 2>&1; tail -2 gpurun_out/stress_batch.log
for c in 2 4; do timeout 300 python bench.py --scan-points 0 --no-cpu-baseline --concurrency $c > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_c$c.json').read().strip().splitlines()[-1]); print('concurrency $c', d['value'], d['e2e']['value'])"; done
