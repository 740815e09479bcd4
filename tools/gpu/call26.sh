set -x
mkdir -p gpurun_out/c26
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "moe" > gpurun_out/c26/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c26/pytest.log
tail -n 12 gpurun_out/c26/pytest.log
for c in dbrx quant spec; do
  timeout 400 python bench.py --config $c --steps 48 --warmup 4 > gpurun_out/c26/bench_$c.json 2> gpurun_out/c26/bench_$c.err; echo "rc=$?"
  tail -c 1500 gpurun_out/c26/bench_$c.json; tail -n 5 gpurun_out/c26/bench_$c.err
done
timeout 400 python bench.py --steps 64 --warmup 4 --skip-ci > gpurun_out/c26/bench_llama8b.json 2> gpurun_out/c26/bench_llama8b.err; echo "rc=$?"
tail -c 1800 gpurun_out/c26/bench_llama8b.json; tail -n 5 gpurun_out/c26/bench_llama8b.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/c26/ncu_r2_all python tools/ncu_target_all.py > gpurun_out/c26/ncu.log 2>&1; echo "ncu rc=$?"
tail -n 5 gpurun_out/c26/ncu.log
ls -la gpurun_out/c26/
timeout 200 ncu -i gpurun_out/c26/ncu_r2_all.ncu-rep --page raw --csv > gpurun_out/c26/ncu_r2_all_raw.csv 2>/dev/null
