set -x
mkdir -p gpurun_out/c17
export NCCL_DEBUG=WARN
timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm" > gpurun_out/c17/pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/c17/pytest_gemm.log
tail -n 3 gpurun_out/c17/pytest_gemm.log
timeout 200 python tools/bench_kernels.py > gpurun_out/c17/kernels.txt 2>&1; grep "^gemm" gpurun_out/c17/kernels.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29731 tests/mp/nvls_worker.py > gpurun_out/c17/nvls_tp2.log 2>&1; echo "rc=$?" >> gpurun_out/c17/nvls_tp2.log
grep -v "^\*\|OMP\|^$" gpurun_out/c17/nvls_tp2.log | tail -n 12
timeout 400 python -m pytest tests/test_tp_gpu.py -x -q -m gpu -k "2" > gpurun_out/c17/pytest_tp2.log 2>&1; echo "rc=$?" >> gpurun_out/c17/pytest_tp2.log
tail -n 6 gpurun_out/c17/pytest_tp2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29732 bench.py --gpus 2 --steps 32 --warmup 8 --skip-ci > gpurun_out/c17/bench_tp2.json 2> gpurun_out/c17/bench_tp2.err
python -c "import json; d=json.load(open('gpurun_out/c17/bench_tp2.json')); print('tp2 ms/step', d['ms_per_step'], 'ttft', d['ttft_p50_ms'])"; tail -n 4 gpurun_out/c17/bench_tp2.err
