timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -3 gpurun_out/bench_1gpu.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1]);print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline_lu']['frac'], d['gpu_launches'], d.get('scan'), d.get('cpu_baseline'))"
