set -x
mkdir -p gpurun_out/c10
for w in 0 2 3 4 6; do
NXDI_B200_GEMV_INFLIGHT=$w timeout 300 python bench.py --steps 64 --warmup 8 --skip-ci > gpurun_out/c10/bench_tp1_mega_w$w.json 2> gpurun_out/c10/bench_tp1_mega_w$w.err
NXDI_B200_GEMV_INFLIGHT=$w timeout 300 python bench.py --steps 64 --warmup 8 --shard-shapes 8 > gpurun_out/c10/bench_tp8s_mega_w$w.json 2> gpurun_out/c10/bench_tp8s_mega_w$w.err
NXDI_B200_DECODE_STEP=0 NXDI_B200_GEMV_INFLIGHT=$w timeout 300 python bench.py --steps 64 --warmup 8 --skip-ci > gpurun_out/c10/bench_tp1_kern_w$w.json 2> gpurun_out/c10/bench_tp1_kern_w$w.err
NXDI_B200_DECODE_STEP=0 NXDI_B200_GEMV_INFLIGHT=$w timeout 300 python bench.py --steps 64 --warmup 8 --shard-shapes 8 > gpurun_out/c10/bench_tp8s_kern_w$w.json 2> gpurun_out/c10/bench_tp8s_kern_w$w.err
done
NXDI_B200_GEMV_INFLIGHT=3 timeout 200 python tools/prof_decode.py --layers 4 --shard-shapes 8 > gpurun_out/c10/prof_tp8shapes_mega_w3.txt 2>&1
NXDI_B200_GEMV_INFLIGHT=3 timeout 200 python tools/prof_decode.py --layers 4 > gpurun_out/c10/prof_tp1_mega_w3.txt 2>&1
for f in gpurun_out/c10/bench_*.json; do echo -n "$f "; python -c "import json,sys; d=json.load(open('$f')); print(round(d['ms_per_step'],3))"; done
