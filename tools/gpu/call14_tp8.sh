set -x
mkdir -p gpurun_out/c14
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29721 tests/mp/symm_worker.py > gpurun_out/c14/symm_tp8.log 2>&1; echo "rc=$?" >> gpurun_out/c14/symm_tp8.log
timeout 400 python -m pytest tests/test_tp_gpu.py -x -q -m gpu -k "8" > gpurun_out/c14/pytest_tp8.log 2>&1; echo "rc=$?" >> gpurun_out/c14/pytest_tp8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 8 --steps 64 --warmup 8 --skip-ci > gpurun_out/c14/bench_tp8.json 2> gpurun_out/c14/bench_tp8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29723 bench.py --gpus 4 --steps 64 --warmup 8 --skip-ci > gpurun_out/c14/bench_tp4.json 2> gpurun_out/c14/bench_tp4.err
grep -v "^\*\|OMP" gpurun_out/c14/symm_tp8.log | tail -n 4; tail -n 3 gpurun_out/c14/pytest_tp8.log; cut -c1-260 gpurun_out/c14/bench_tp8.json; cut -c1-260 gpurun_out/c14/bench_tp4.json; tail -n 3 gpurun_out/c14/bench_tp8.err
