set -x
mkdir -p gpurun_out/c18
timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm" > gpurun_out/c18/pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/c18/pytest_gemm.log
tail -n 3 gpurun_out/c18/pytest_gemm.log
timeout 200 python tools/bench_kernels.py --json gpurun_out/c18/kernels.json > gpurun_out/c18/kernels.txt 2>&1; grep "^gemm" gpurun_out/c18/kernels.txt
NXDI_B200_GEMM_TM=2 timeout 200 python tools/bench_kernels.py > gpurun_out/c18/kernels_tm2.txt 2>&1; grep "^gemm" gpurun_out/c18/kernels_tm2.txt
timeout 300 python bench.py --steps 32 --warmup 8 --skip-ci > gpurun_out/c18/bench_tp1.json 2> gpurun_out/c18/bench_tp1.err
python -c "import json; d=json.load(open('gpurun_out/c18/bench_tp1.json')); print('ms/step', d['ms_per_step'], 'ttft', d['ttft_p50_ms'])"
