set -x
mkdir -p gpurun_out/c21
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention_prefill" > gpurun_out/c21/pytest_attn.log 2>&1; echo "rc=$?" >> gpurun_out/c21/pytest_attn.log
tail -n 4 gpurun_out/c21/pytest_attn.log
timeout 200 python tools/bench_kernels.py --json gpurun_out/c21/kernels.json > gpurun_out/c21/kernels.txt 2>&1; grep "prefill" gpurun_out/c21/kernels.txt
NXDI_B200_ATTN_TC=0 timeout 200 python tools/bench_kernels.py > gpurun_out/c21/kernels_mma.txt 2>&1; grep "prefill" gpurun_out/c21/kernels_mma.txt
