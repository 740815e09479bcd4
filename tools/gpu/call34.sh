set -x
mkdir -p gpurun_out/c34
nvidia-smi -L | head -3
timeout 400 python -m pytest tests/test_tp_gpu.py -q -m gpu > gpurun_out/c34/pytest_tp2.log 2>&1; echo "rc=$?" >> gpurun_out/c34/pytest_tp2.log
tail -n 15 gpurun_out/c34/pytest_tp2.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 48 --warmup 4 --skip-ci > gpurun_out/c34/bench_tp2.json 2> gpurun_out/c34/bench_tp2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/c34/bench_tp2.json; tail -n 3 gpurun_out/c34/bench_tp2.err
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 32 --warmup 4 --config quant > gpurun_out/c34/bench_quant_tp2.json 2> gpurun_out/c34/bench_quant_tp2.err; echo "quant rc=$?"
tail -c 600 gpurun_out/c34/bench_quant_tp2.json; tail -n 3 gpurun_out/c34/bench_quant_tp2.err
