set -x
mkdir -p gpurun_out/c15
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm" > gpurun_out/c15/pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/c15/pytest_gemm.log
tail -n 5 gpurun_out/c15/pytest_gemm.log
timeout 300 python tools/bench_kernels.py --json gpurun_out/c15/kernels.json > gpurun_out/c15/kernels.txt 2>&1
grep "^gemm\|^gemv" gpurun_out/c15/kernels.txt
NXDI_B200_DECODE_STEP=0 timeout 300 python bench.py --steps 32 --warmup 8 --skip-ci > gpurun_out/c15/bench_tp1.json 2> gpurun_out/c15/bench_tp1.err
python -c "import json; d=json.load(open('gpurun_out/c15/bench_tp1.json')); print('ms/step', d['ms_per_step'], 'ttft', d['ttft_p50_ms'])"
