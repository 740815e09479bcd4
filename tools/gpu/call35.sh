set -x
mkdir -p gpurun_out/c35
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29721 tests/mp/symm_worker.py > gpurun_out/c35/symm_worker_tp2_r2.log 2>&1; echo "rc=$?" >> gpurun_out/c35/symm_worker_tp2_r2.log
grep -E "quantised|rc=|fused|NCCL|us" gpurun_out/c35/symm_worker_tp2_r2.log | grep "r0\]\|rc=" | tail -20
