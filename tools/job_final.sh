# final single-GPU validation of a tree: full GPU suite (slow end-to-end cases included), smoke, bench lines, ncu evidence
mkdir -p gpurun_out
WG_SLOW=1 timeout 2400 python -m pytest tests -m gpu -x -q --timeout 1500 > gpurun_out/t_gpu_all.log 2>&1; tail -3 gpurun_out/t_gpu_all.log
cat gpurun_out/e2e_gate.log | grep -E "^[a-z]" 
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 600 gpurun_out/bench_1gpu.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_1gpu.json 2> gpurun_out/bench_ref_1gpu.err; tail -c 400 gpurun_out/bench_ref_1gpu.json
timeout 500 python bench.py --workload cfg1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; tail -c 300 gpurun_out/bench_cfg1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r2_cfg2a.csv python tools/one_solve.py > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trisolve_wave -c 2 -o gpurun_out/r02_trisolve_wave2 python tools/one_solve.py > gpurun_out/ncu_tri.log 2>&1; tail -1 gpurun_out/ncu_tri.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wg_residual -c 1 -o gpurun_out/r02_residual python tools/one_solve.py > gpurun_out/ncu_res.log 2>&1; tail -1 gpurun_out/ncu_res.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:diag_invert -c 1 -o gpurun_out/r02_diag_invert python tools/one_solve.py > gpurun_out/ncu_inv.log 2>&1; tail -1 gpurun_out/ncu_inv.log
