set -x
mkdir -p gpurun_out/c6
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29711 tests/mp/symm_worker.py > gpurun_out/c6/symm_tp2.log 2>&1; echo "rc=$?" >> gpurun_out/c6/symm_tp2.log
timeout 400 python -m pytest tests/test_tp_gpu.py -x -q -m gpu -k "2" > gpurun_out/c6/pytest_tp2.log 2>&1; echo "rc=$?" >> gpurun_out/c6/pytest_tp2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 64 --warmup 8 --skip-ci > gpurun_out/c6/bench_tp2.json 2> gpurun_out/c6/bench_tp2.err
tail -n 4 gpurun_out/c6/symm_tp2.log; tail -n 3 gpurun_out/c6/pytest_tp2.log; cut -c1-300 gpurun_out/c6/bench_tp2.json; tail -n 5 gpurun_out/c6/bench_tp2.err
