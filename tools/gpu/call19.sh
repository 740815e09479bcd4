set -x
mkdir -p gpurun_out/c19
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention_prefill or gemm" > gpurun_out/c19/pytest_attn.log 2>&1; echo "rc=$?" >> gpurun_out/c19/pytest_attn.log
tail -n 25 gpurun_out/c19/pytest_attn.log
timeout 200 python tools/bench_kernels.py > gpurun_out/c19/kernels.txt 2>&1; grep "^gemm\|prefill" gpurun_out/c19/kernels.txt
timeout 300 python bench.py --steps 32 --warmup 8 --skip-ci > gpurun_out/c19/bench_tp1.json 2> gpurun_out/c19/bench_tp1.err
python -c "import json; d=json.load(open('gpurun_out/c19/bench_tp1.json')); print('ms/step', d['ms_per_step'], 'ttft', d['ttft_p50_ms'])"
