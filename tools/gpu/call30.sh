set -x
mkdir -p gpurun_out/c30
timeout 900 python -m pytest tests/test_features_gpu.py -x -q -m gpu > gpurun_out/c30/pytest_features.log 2>&1; echo "rc=$?" >> gpurun_out/c30/pytest_features.log
tail -n 30 gpurun_out/c30/pytest_features.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "prefill_split or rope" > gpurun_out/c30/pytest_rope.log 2>&1; echo "rc=$?" >> gpurun_out/c30/pytest_rope.log
tail -n 8 gpurun_out/c30/pytest_rope.log
timeout 300 python tools/trace_prefill.py > gpurun_out/c30/trace_prefill.txt 2>&1
tail -n 24 gpurun_out/c30/trace_prefill.txt
