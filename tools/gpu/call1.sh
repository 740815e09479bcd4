set -x
mkdir -p gpurun_out/c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1/smi.txt 2>&1
timeout 420 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/c1/pytest_kernels.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1/pytest_kernels.log
timeout 300 python bench.py --steps 64 --warmup 8 --skip-ci > gpurun_out/c1/bench_tp1_108.json 2> gpurun_out/c1/bench_tp1_108.err
NXDI_B200_GEMV_SMEM_KB=224 timeout 300 python bench.py --steps 64 --warmup 8 --skip-ci > gpurun_out/c1/bench_tp1_224.json 2> gpurun_out/c1/bench_tp1_224.err
timeout 300 python bench.py --steps 64 --warmup 8 --shard-shapes 8 > gpurun_out/c1/bench_tp8shapes_108.json 2> gpurun_out/c1/bench_tp8shapes_108.err
NXDI_B200_GEMV_SMEM_KB=224 timeout 300 python bench.py --steps 64 --warmup 8 --shard-shapes 8 > gpurun_out/c1/bench_tp8shapes_224.json 2> gpurun_out/c1/bench_tp8shapes_224.err
timeout 200 python tools/trace_decode.py --layers 4 --out gpurun_out/c1/trace_tp1.json > gpurun_out/c1/trace_tp1.txt 2>&1
timeout 200 python tools/trace_decode.py --layers 4 --shard-shapes 8 --out gpurun_out/c1/trace_tp8shapes.json > gpurun_out/c1/trace_tp8shapes.txt 2>&1
timeout 200 python tools/bench_gemv_fixed.py > gpurun_out/c1/gemv_fixed.txt 2>&1
tail -3 gpurun_out/c1/pytest_kernels.log; cat gpurun_out/c1/bench_*.json | cut -c1-400
