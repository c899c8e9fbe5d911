mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -x -q --timeout 300 -k "trisolve or batched" > gpurun_out/t_tri.log 2>&1; tail -4 gpurun_out/t_tri.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "refinement or Strided or assemble_rows" -s > gpurun_out/t_refine.log 2>&1; grep -E "cond~|passed|failed|Error|assert" gpurun_out/t_refine.log | head -30
timeout 300 python tools/batch_probe.py 1900 32,64,128 > gpurun_out/batch_probe2.log 2>&1; cat gpurun_out/batch_probe2.log
timeout 500 python bench.py --workload cfg1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; tail -c 3000 gpurun_out/bench_cfg1.json; tail -5 gpurun_out/bench_cfg1.err
timeout 300 python bench.py --workload cfg5stream --steps 3 > gpurun_out/bench_cfg5stream.json 2> gpurun_out/bench_cfg5stream.err; cat gpurun_out/bench_cfg5stream.json; tail -5 gpurun_out/bench_cfg5stream.err
BATCH_PROBE_ITERS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/ncu_batch64.csv python tools/batch_probe.py 1900 64 0 > gpurun_out/ncu_batch64.log 2>&1; tail -2 gpurun_out/ncu_batch64.log
timeout 500 python tools/parity_fuzz.py 150 1 > gpurun_out/pfuzz_r2.log 2>&1; tail -3 gpurun_out/pfuzz_r2.log
timeout 1000 python bench.py --workload cfg5fit --steps 2 > gpurun_out/bench_cfg5fit.json 2> gpurun_out/bench_cfg5fit.err; tail -c 2500 gpurun_out/bench_cfg5fit.json; tail -5 gpurun_out/bench_cfg5fit.err
