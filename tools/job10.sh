mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "Strided or strided or refinement" > gpurun_out/t_strided.log 2>&1; tail -3 gpurun_out/t_strided.log
timeout 300 python tools/leak_probe.py > gpurun_out/leak.log 2>&1; tail -3 gpurun_out/leak.log
