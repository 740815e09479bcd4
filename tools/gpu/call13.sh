set -x
mkdir -p gpurun_out/c13
export NXDI_B200_DECODE_STEP=0
for kb in 108 224; do for w in 0 4 6 8; do
NXDI_B200_GEMV_SMEM_KB=$kb NXDI_B200_GEMV_INFLIGHT=$w timeout 300 python bench.py --steps 64 --warmup 8 --skip-ci > gpurun_out/c13/bench_tp1_kb${kb}_w$w.json 2> gpurun_out/c13/bench_tp1_kb${kb}_w$w.err
NXDI_B200_GEMV_SMEM_KB=$kb NXDI_B200_GEMV_INFLIGHT=$w timeout 300 python bench.py --steps 64 --warmup 8 --shard-shapes 8 > gpurun_out/c13/bench_tp8s_kb${kb}_w$w.json 2> gpurun_out/c13/bench_tp8s_kb${kb}_w$w.err
done; done
for f in gpurun_out/c13/bench_*.json; do echo -n "$f "; python -c "import json,sys; d=json.load(open('$f')); print(round(d['ms_per_step'],3), round(d['ttft_p50_ms'],2))"; done
