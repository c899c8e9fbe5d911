# single-GPU run of the final tree (factor re-use in the scan): GPU suite, smoke, the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/t_gpu_all.log 2>&1; tail -3 gpurun_out/t_gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
if [ "$1" = bench ]; then
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo rc=$?; tail -c 300 gpurun_out/bench_1gpu.json
fi
grep "re-use" gpurun_out/e2e_gate.log
