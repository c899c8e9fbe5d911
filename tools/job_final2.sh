# final single-GPU run of the final tree: GPU suite, smoke, every bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/t_gpu_all.log 2>&1; tail -3 gpurun_out/t_gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 300 gpurun_out/bench_1gpu.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_1gpu.json 2> gpurun_out/bench_ref_1gpu.err; tail -c 200 gpurun_out/bench_ref_1gpu.json
timeout 500 python bench.py --workload cfg1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; tail -c 200 gpurun_out/bench_cfg1.json
timeout 500 python bench.py --workload cfg3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 200 gpurun_out/bench_cfg3.json
timeout 500 python bench.py --workload cfg4 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; tail -c 200 gpurun_out/bench_cfg4.json
timeout 500 python bench.py --workload cfg2b --steps 6 --no-cpu-baseline > gpurun_out/bench_cfg2b.json 2> gpurun_out/bench_cfg2b.err; tail -c 200 gpurun_out/bench_cfg2b.json
