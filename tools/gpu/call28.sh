set -x
mkdir -p gpurun_out/c28
timeout 900 python -m pytest tests/test_features_gpu.py -x -q -m gpu > gpurun_out/c28/pytest_features.log 2>&1; echo "rc=$?" >> gpurun_out/c28/pytest_features.log
tail -n 25 gpurun_out/c28/pytest_features.log
timeout 400 python bench.py --config spec --steps 48 --warmup 4 > gpurun_out/c28/bench_spec.json 2> gpurun_out/c28/bench_spec.err; echo "rc=$?"
tail -c 1800 gpurun_out/c28/bench_spec.json; tail -n 5 gpurun_out/c28/bench_spec.err
