set -x
mkdir -p gpurun_out/c12
timeout 420 python -m pytest tests/test_kernels_gpu.py tests/test_features_gpu.py -x -q -m gpu > gpurun_out/c12/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c12/pytest.log
tail -n 4 gpurun_out/c12/pytest.log
for w in 0 6; do
NXDI_B200_DECODE_STEP=0 NXDI_B200_GEMV_INFLIGHT=$w timeout 300 python bench.py --steps 64 --warmup 8 --skip-ci > gpurun_out/c12/bench_tp1_kern_w$w.json 2> gpurun_out/c12/bench_tp1_kern_w$w.err
NXDI_B200_DECODE_STEP=0 NXDI_B200_GEMV_INFLIGHT=$w timeout 300 python bench.py --steps 64 --warmup 8 --shard-shapes 8 > gpurun_out/c12/bench_tp8s_kern_w$w.json 2> gpurun_out/c12/bench_tp8s_kern_w$w.err
done
timeout 300 python bench.py --steps 64 --warmup 8 --skip-ci > gpurun_out/c12/bench_tp1_mega.json 2> gpurun_out/c12/bench_tp1_mega.err
timeout 300 python bench.py --steps 64 --warmup 8 --shard-shapes 8 > gpurun_out/c12/bench_tp8s_mega.json 2> gpurun_out/c12/bench_tp8s_mega.err
NXDI_B200_DECODE_STEP=0 timeout 200 python tools/prof_decode.py --layers 4 --shard-shapes 8 > gpurun_out/c12/prof_tp8shapes_kern.txt 2>&1
timeout 200 python tools/prof_decode.py --layers 4 --shard-shapes 8 > gpurun_out/c12/prof_tp8shapes_mega.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 2 -o gpurun_out/c12/ncu_attn python tools/ncu_target.py > gpurun_out/c12/ncu_attn.log 2>&1
for f in gpurun_out/c12/bench_*.json; do echo -n "$f "; python -c "import json,sys; d=json.load(open('$f')); print(round(d['ms_per_step'],3))"; done
