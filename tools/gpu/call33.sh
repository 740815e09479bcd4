set -x
mkdir -p gpurun_out/c33
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/c33/pytest_gpu_full.log 2>&1; echo "rc=$?" >> gpurun_out/c33/pytest_gpu_full.log
tail -n 25 gpurun_out/c33/pytest_gpu_full.log
timeout 500 python bench.py > gpurun_out/c33/bench_default.json 2> gpurun_out/c33/bench_default.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/c33/bench_default.json; tail -n 3 gpurun_out/c33/bench_default.err
timeout 200 python bench.py --config dbrx --steps 48 --warmup 4 > gpurun_out/c33/bench_dbrx.json 2> gpurun_out/c33/bench_dbrx.err; echo "dbrx rc=$?"
tail -c 700 gpurun_out/c33/bench_dbrx.json
NEW="moe_grouped or fp8_w8a8 or prefill_split or attention_prefill_tcgen05 or quantized_gemv or weight_only"
timeout 170 bash tools/sanitize.sh memcheck "$NEW" > gpurun_out/c33/sanitizer_memcheck_r2.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/c33/sanitizer_memcheck_r2.log
tail -n 4 gpurun_out/c33/sanitizer_memcheck_r2.log
timeout 170 bash tools/sanitize.sh racecheck "moe_grouped or prefill_split or (quantized_gemv and 4096-4096) or (attention_prefill_tcgen05 and 128-2)" > gpurun_out/c33/sanitizer_racecheck_r2.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/c33/sanitizer_racecheck_r2.log
tail -n 4 gpurun_out/c33/sanitizer_racecheck_r2.log
timeout 120 bash tools/sanitize.sh synccheck "moe_grouped or prefill_split or fp8_w8a8 or (attention_prefill_tcgen05 and 128-2)" > gpurun_out/c33/sanitizer_synccheck_r2.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/c33/sanitizer_synccheck_r2.log
tail -n 4 gpurun_out/c33/sanitizer_synccheck_r2.log
