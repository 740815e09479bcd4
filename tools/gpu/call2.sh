set -x
mkdir -p gpurun_out/c2
timeout 420 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/c2/pytest_kernels.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2/pytest_kernels.log
for kb in 108 160 224; do
NXDI_B200_GEMV_SMEM_KB=$kb timeout 300 python bench.py --steps 64 --warmup 8 --skip-ci > gpurun_out/c2/bench_tp1_$kb.json 2> gpurun_out/c2/bench_tp1_$kb.err
NXDI_B200_GEMV_SMEM_KB=$kb timeout 300 python bench.py --steps 64 --warmup 8 --shard-shapes 8 > gpurun_out/c2/bench_tp8shapes_$kb.json 2> gpurun_out/c2/bench_tp8shapes_$kb.err
done
timeout 200 python tools/trace_decode.py --layers 4 --out gpurun_out/c2/trace_tp1.json > gpurun_out/c2/trace_tp1.txt 2>&1
timeout 200 python tools/trace_decode.py --layers 4 --shard-shapes 8 --out gpurun_out/c2/trace_tp8shapes.json > gpurun_out/c2/trace_tp8shapes.txt 2>&1
timeout 200 python tools/bench_gemv_fixed.py > gpurun_out/c2/gemv_fixed.txt 2>&1
tail -3 gpurun_out/c2/pytest_kernels.log; for f in gpurun_out/c2/bench_*.json; do echo $f; cut -c1-330 $f; done
