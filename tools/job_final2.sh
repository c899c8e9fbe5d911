# final single-GPU run of the final tree: GPU suite, smoke, bench lines, ncu evidence of the kernels that changed last
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/t_gpu_all.log 2>&1; tail -3 gpurun_out/t_gpu_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 300 gpurun_out/bench_1gpu.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_1gpu.json 2> gpurun_out/bench_ref_1gpu.err; tail -c 200 gpurun_out/bench_ref_1gpu.json
timeout 500 python bench.py --workload cfg1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; tail -c 200 gpurun_out/bench_cfg1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r2_cfg2a.csv python tools/one_solve.py > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trisolve_wave -c 2 -o gpurun_out/r02_trisolve_wave2 -f python tools/one_solve.py > gpurun_out/ncu_tri.log 2>&1; tail -1 gpurun_out/ncu_tri.log
