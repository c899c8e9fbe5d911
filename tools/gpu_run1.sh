#!/usr/bin/env bash
# first GPU session: correctness of each piece in order of risk, then timings
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
ls /root/reference > gpurun_out/ref_present.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dense.py -q -m gpu --timeout 300 -k "dgemm" > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?"
tail -5 gpurun_out/t_gemm.log
timeout 1200 python -m pytest tests/test_gpu_dense.py -q -m gpu --timeout 300 -k "lu" > gpurun_out/t_lu.log 2>&1; echo "lu rc=$?"
tail -15 gpurun_out/t_lu.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
tail -25 gpurun_out/t_parity.log
timeout 900 python -m pytest tests/test_gpu_dense.py -q -m gpu --timeout 600 -k "full_size" > gpurun_out/t_full.log 2>&1; echo "full rc=$?"
tail -8 gpurun_out/t_full.log
timeout 600 python tools/quick_bench.py --P 1 --M 20 --N 11 > gpurun_out/qb_cfg1.log 2>&1; tail -12 gpurun_out/qb_cfg1.log
timeout 900 python tools/quick_bench.py --P 1 --M 40 --N 21 --dgemm 0 > gpurun_out/qb_cfg2a.log 2>&1; tail -12 gpurun_out/qb_cfg2a.log
timeout 900 python tools/quick_bench.py --P 1 --M 40 --N 21 --dgemm 0 --graph 1 > gpurun_out/qb_cfg2a_graph.log 2>&1; tail -8 gpurun_out/qb_cfg2a_graph.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
