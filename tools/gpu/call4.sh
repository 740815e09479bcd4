set -x
mkdir -p gpurun_out/c4
timeout 300 tools/bin/bench_stream > gpurun_out/c4/stream.txt 2>&1
timeout 420 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/c4/pytest_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/c4/pytest_kernels.log
timeout 200 python tools/prof_decode.py --layers 4 --shard-shapes 8 > gpurun_out/c4/prof_tp8shapes.txt 2>&1
timeout 300 python bench.py --steps 64 --warmup 8 --skip-ci > gpurun_out/c4/bench_tp1_108.json 2> gpurun_out/c4/bench_tp1_108.err
timeout 300 python bench.py --steps 64 --warmup 8 --shard-shapes 8 > gpurun_out/c4/bench_tp8shapes_108.json 2> gpurun_out/c4/bench_tp8shapes_108.err
cat gpurun_out/c4/stream.txt; tail -n 3 gpurun_out/c4/pytest_kernels.log; cut -c1-200 gpurun_out/c4/bench_*.json
