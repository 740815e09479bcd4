set -x
mkdir -p gpurun_out/c3
timeout 120 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "test_gemv and 6144-4096-8" > gpurun_out/c3/t8_pl4.log 2>&1; echo "rc=$?" >> gpurun_out/c3/t8_pl4.log
NXDI_B200_GEMV_PL=1 timeout 120 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "test_gemv and 6144-4096-8" > gpurun_out/c3/t8_pl1.log 2>&1; echo "rc=$?" >> gpurun_out/c3/t8_pl1.log
NXDI_B200_GEMV_SMEM_KB=224 timeout 120 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "test_gemv and 6144-4096-8" > gpurun_out/c3/t8_224.log 2>&1; echo "rc=$?" >> gpurun_out/c3/t8_224.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "test_gemv and 6144-4096-8" > gpurun_out/c3/t8_memcheck.log 2>&1
timeout 200 python tools/prof_decode.py --layers 4 --shard-shapes 8 > gpurun_out/c3/prof_tp8shapes.txt 2>&1
NXDI_B200_GEMV_SMEM_KB=224 timeout 200 python tools/prof_decode.py --layers 4 --shard-shapes 8 > gpurun_out/c3/prof_tp8shapes_224.txt 2>&1
NXDI_B200_GEMV_SMEM_KB=224 timeout 200 python tools/prof_decode.py --layers 4 > gpurun_out/c3/prof_tp1_224.txt 2>&1
tail -2 gpurun_out/c3/t8_*.log
