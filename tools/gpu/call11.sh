set -x
mkdir -p gpurun_out/c11
export NXDI_B200_DECODE_STEP=0
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv2_kernel -s 12 -c 4 -o gpurun_out/c11/ncu_gemv2_tp8 python tools/ncu_target_tp8.py > gpurun_out/c11/ncu_tp8.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv2_kernel -s 9 -c 3 -o gpurun_out/c11/ncu_gemv2_tp1 python tools/ncu_target.py > gpurun_out/c11/ncu_tp1.log 2>&1
timeout 200 python tools/bench_gemv_fixed.py > gpurun_out/c11/gemv_fixed.txt 2>&1
ls -la gpurun_out/c11; tail -n 3 gpurun_out/c11/ncu_tp8.log; cat gpurun_out/c11/gemv_fixed.txt | head -12
