set -x
mkdir -p gpurun_out/c8
timeout 300 python -m pytest tests/test_features_gpu.py -x -q -m gpu -k "persistent_decode_step" > gpurun_out/c8/pytest_dstep.log 2>&1; echo "rc=$?" >> gpurun_out/c8/pytest_dstep.log
tail -n 25 gpurun_out/c8/pytest_dstep.log
timeout 300 python bench.py --steps 64 --warmup 8 --skip-ci > gpurun_out/c8/bench_tp1.json 2> gpurun_out/c8/bench_tp1.err
timeout 300 python bench.py --steps 64 --warmup 8 --shard-shapes 8 > gpurun_out/c8/bench_tp8shapes.json 2> gpurun_out/c8/bench_tp8shapes.err
NXDI_B200_DECODE_STEP=0 timeout 300 python bench.py --steps 64 --warmup 8 --skip-ci > gpurun_out/c8/bench_tp1_off.json 2> gpurun_out/c8/bench_tp1_off.err
cut -c1-220 gpurun_out/c8/bench_*.json; tail -n 3 gpurun_out/c8/bench_tp1.err
