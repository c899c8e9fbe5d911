# single-GPU run of the final tree (factor re-use in the scan): GPU suite, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/t_gpu_all.log 2>&1; tail -3 gpurun_out/t_gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo rc=$?; tail -c 600 gpurun_out/bench_1gpu.json
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_1gpu.json").read().strip().splitlines()[-1]); s = d["scan"]
print(d["value"], d["e2e"], d["roofline"]["frac"], d["roofline_lu"]["frac"], d.get("parity"))
print({k: s.get(k) for k in ("value", "seconds", "points", "boltzmann_solves", "host_seconds_per_point", "checksum_first64", "checksum_moments_first64", "pressure_first", "pressure_last", "host_workers_per_rank")}, s.get("factor_reuse"))
P
