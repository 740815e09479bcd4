set -x
mkdir -p gpurun_out/c25
timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "moe or fp8_w8a8 or gemm" > gpurun_out/c25/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c25/pytest.log
tail -n 25 gpurun_out/c25/pytest.log
timeout 200 python tools/bench_moe.py > gpurun_out/c25/moe_bench.txt 2>&1
cat gpurun_out/c25/moe_bench.txt
timeout 120 python - > gpurun_out/c25/fp8_bench.txt 2>&1 <<'PY'
import math, sys, torch
sys.path.insert(0, ".")
from neuronx_distributed_inference_b200 import ops
C = ops._C()
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (M, N, K) in [(2048, 28672, 4096), (2048, 4096, 14336), (8192, 8192, 8192), (256, 28672, 4096)]:
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K))
    wq = w.to(torch.float8_e4m3fn); ws = torch.ones(N, device="cuda")
    wb = w.to(torch.bfloat16)
    xq, a_s = C.rmsnorm_quant(x, None, 1e-5, 0.0, float("inf"))
    us8 = t(lambda: C.gemm_fp8(xq, a_s, wq, ws, None, 0, None))
    us16 = t(lambda: C.gemm(x, wb, None, 0, None, None))
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K}: fp8 gemm {us8:8.1f} us ({fl/us8/1e6:7.1f} TFLOP/s)  bf16 gemm {us16:8.1f} us ({fl/us16/1e6:7.1f} TFLOP/s)", flush=True)
PY
cat gpurun_out/c25/fp8_bench.txt
