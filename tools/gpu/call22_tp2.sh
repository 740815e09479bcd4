set -x
mkdir -p gpurun_out/c23
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29741 tests/mp/nvls_worker.py > gpurun_out/c23/nvls_tp2.log 2>&1; echo "rc=$?" >> gpurun_out/c23/nvls_tp2.log
grep -v "^\*\|OMP\|^$" gpurun_out/c23/nvls_tp2.log | tail -n 16
SEQUENCE_PARALLEL=1 timeout 400 python -m pytest tests/test_tp_gpu.py -x -q -m gpu -k "2" > gpurun_out/c23/pytest_tp2_sp.log 2>&1; echo "rc=$?" >> gpurun_out/c23/pytest_tp2_sp.log
tail -n 5 gpurun_out/c23/pytest_tp2_sp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 32 --warmup 8 --skip-ci > gpurun_out/c23/bench_tp2.json 2> gpurun_out/c23/bench_tp2.err
python -c "import json; d=json.load(open('gpurun_out/c23/bench_tp2.json')); print('tp2 ms/step', d['ms_per_step'], 'ttft', d['ttft_p50_ms'])"; grep -i "warn\|heap" gpurun_out/c23/bench_tp2.err | head -5
