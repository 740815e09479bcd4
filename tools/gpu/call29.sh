set -x
mkdir -p gpurun_out/c29
timeout 900 python -m pytest tests/test_features_gpu.py -x -q -m gpu > gpurun_out/c29/pytest_features.log 2>&1; echo "rc=$?" >> gpurun_out/c29/pytest_features.log
tail -n 30 gpurun_out/c29/pytest_features.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "moe" > gpurun_out/c29/pytest_moe.log 2>&1; echo "rc=$?" >> gpurun_out/c29/pytest_moe.log
tail -n 8 gpurun_out/c29/pytest_moe.log
