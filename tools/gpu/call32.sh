set -x
mkdir -p gpurun_out/c32
timeout 900 python -m pytest tests/test_features_gpu.py -q -m gpu > gpurun_out/c32/pytest_features.log 2>&1; echo "rc=$?" >> gpurun_out/c32/pytest_features.log
tail -n 12 gpurun_out/c32/pytest_features.log
