# knob sweeps (slots / panel width / 32-column leaves) and fresh --set full captures of the trailing-update GEMM,
# the assembly and one leaf panel on the round-2 tree
mkdir -p gpurun_out
S=gpurun_out/knob_sweep.log; : > $S
timeout 240 python tools/knob_sweep.py 1 40 21 slots=2,3,4 leaf32=0,1 per=4 >> $S 2>&1
timeout 160 python tools/knob_sweep.py 1 40 21 slots=3 nb=192,128 per=4 >> $S 2>&1
timeout 120 python tools/knob_sweep.py 1 40 21 slots=3 split=0,35,65 per=4 >> $S 2>&1
timeout 240 python tools/knob_sweep.py 4 30 11 slots=3,4,5,6,8 leaf32=0,1 per=4 >> $S 2>&1
timeout 120 python tools/knob_sweep.py 4 30 11 slots=4 nb=192,128 per=4 >> $S 2>&1
timeout 240 python tools/knob_sweep.py 3 30 11 slots=4,6,8,12 leaf32=0,1 per=4 >> $S 2>&1
timeout 240 python tools/knob_sweep.py 1 20 11 slots=0 nb=256,128 leaf32=0,1 >> $S 2>&1
grep -v Warning $S | tail -60
for n in 15600 11600; do timeout 120 python tools/leaf32_probe.py $n >> gpurun_out/leaf32_probe.log 2>&1; done; tail -8 gpurun_out/leaf32_probe.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dgemm_sub -c 1 -o gpurun_out/r02_gemm -f python tools/gemm_probe.py 15600 256 3 1 > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wg_assemble_kernel -c 1 -o gpurun_out/r02_assemble -f python tools/assemble_probe.py > gpurun_out/ncu_asm.log 2>&1; tail -1 gpurun_out/ncu_asm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lu_panel_leaf -s 700 -c 1 -o gpurun_out/r02_leaf -f python tools/one_solve.py > gpurun_out/ncu_leaf.log 2>&1; tail -1 gpurun_out/ncu_leaf.log
ls -la gpurun_out/*.ncu-rep
