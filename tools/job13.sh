mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -x -q --timeout 300 -k "trisolve or batched or lu_random or fused" > gpurun_out/t_tri.log 2>&1; tail -3 gpurun_out/t_tri.log
timeout 300 python tools/trisolve_probe.py 15600 > gpurun_out/trisolve_probe.log 2>&1; cat gpurun_out/trisolve_probe.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_refparity.py -m gpu -x -q --timeout 300 > gpurun_out/t_par.log 2>&1; tail -2 gpurun_out/t_par.log
