set -x
mkdir -p gpurun_out/c27
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "quantized_gemv or weight_only_prefill or gemv" > gpurun_out/c27/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c27/pytest.log
tail -n 25 gpurun_out/c27/pytest.log
for c in quant spec; do
  timeout 400 python bench.py --config $c --steps 48 --warmup 4 > gpurun_out/c27/bench_$c.json 2> gpurun_out/c27/bench_$c.err; echo "rc=$?"
  tail -c 1600 gpurun_out/c27/bench_$c.json; tail -n 5 gpurun_out/c27/bench_$c.err
done
