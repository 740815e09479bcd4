set -x
mkdir -p gpurun_out/c31
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "prefill_split or rope or test_gemv" > gpurun_out/c31/pytest_rope.log 2>&1; echo "rc=$?" >> gpurun_out/c31/pytest_rope.log
tail -n 40 gpurun_out/c31/pytest_rope.log
timeout 300 python tools/debug_pad.py 2>&1 | grep -v Warning | tail -8
