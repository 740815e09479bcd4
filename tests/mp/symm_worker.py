"""torchrun worker: fused GEMV->all-reduce kernel vs GEMM + NCCL all-reduce (numerics, graph capture, timing)."""
import faulthandler
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
faulthandler.dump_traceback_later(int(os.environ.get("DUMP_AFTER", "60")), exit=True)


def log(*a):
    print(f"[r{os.environ.get('RANK')}]", *a, flush=True)


def _hard_exit(code=0):
    """Tearing down NCCL communicators that were captured into CUDA graphs can block forever in
    destroy_process_group(); results are already printed, so flush and leave."""
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(code)


def main():
    from neuronx_distributed_inference_b200.parallel import state as pstate
    from neuronx_distributed_inference_b200.parallel.symm import SymmetricWorkspace
    from neuronx_distributed_inference_b200 import ops
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    pstate.init_distributed("nccl")
    pstate.initialize_model_parallel(tensor_model_parallel_size=world)
    g = pstate.get_tensor_model_parallel_group()
    dev = torch.device("cuda", torch.cuda.current_device())
    log("groups ok")
    g.symm = SymmetricWorkspace.create(g, dev, max_width=8192)
    log("workspace ok", [hex(p) for p in g.symm.recv_ptrs])
    torch.manual_seed(rank)
    worst = 0.0
    for (T, N, K) in [(2, 4096, 512), (2, 4096, 1792), (1, 1024, 256), (8, 4096, 2048), (3, 40, 512), (2, 4096, 7168),
                      (2, 4096, 2048), (2, 8192, 14336)]:
        x = torch.randn(T, K, device=dev, dtype=torch.bfloat16)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        res = torch.randn(T, N, device=dev, dtype=torch.bfloat16)
        dist.broadcast(res, 0)
        for it in range(3):
            g.symm.begin_step()
            y = ops.linear_allreduce(x, w, None, g, residual=res)
            torch.cuda.synchronize()
        ref = (x.float() @ w.float().t())
        dist.all_reduce(ref)
        ref = ref + res.float()
        err = ((y.float() - ref).norm() / ref.norm()).item()
        worst = max(worst, err)
        # every rank must hold bitwise the same result
        ys = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(ys, y)
        same = all(torch.equal(ys[0], t) for t in ys)
        log(f"T={T} N={N} K={K} rel_err={err:.2e} identical_across_ranks={same}")
        assert err < 1e-2 and same
    # weight-only 8-bit row-parallel layers through the same fused kernel (bytes streamed by TMA, expanded in registers)
    from neuronx_distributed_inference_b200.ops import reference as refops
    for wdt in (torch.float8_e4m3fn, torch.int8):
        for (T, N, K) in [(2, 4096, 1792), (4, 4096, 1408), (1, 2048, 512)]:
            x = torch.randn(T, K, device=dev, dtype=torch.bfloat16)
            q, sc = refops.quantize_per_channel(torch.randn(N, K, device=dev) / K ** 0.5, wdt)
            res = torch.randn(T, N, device=dev, dtype=torch.bfloat16)
            dist.broadcast(res, 0)
            n0 = ops.stats["gemv_allreduce"]
            g.symm.begin_step()
            y = ops.linear_allreduce(x, q, None, g, residual=res, scale=sc)
            torch.cuda.synchronize()
            assert ops.stats["gemv_allreduce"] == n0 + 1
            ref = x.float() @ (q.float() * sc.float()[:, None]).t()
            dist.all_reduce(ref)
            ref = ref + res.float()
            err = ((y.float() - ref).norm() / ref.norm()).item()
            ys = [torch.empty_like(y) for _ in range(world)]
            dist.all_gather(ys, y)
            same = all(torch.equal(ys[0], t) for t in ys)
            log(f"quantised {wdt} T={T} N={N} K={K} rel_err={err:.2e} identical_across_ranks={same}")
            assert err < 1e-2 and same
    # graph capture + replay, timing vs NCCL
    T, N, K = 2, 4096, 512
    x = torch.randn(T, K, device=dev, dtype=torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    res = torch.randn(T, N, device=dev, dtype=torch.bfloat16)
    g.symm.ensure_even()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        g.symm.begin_step()      # captured: every replay bumps the device-side step counter -> fresh tags
        y = res
        for _ in range(64):
            y = ops.linear_allreduce(x, w, None, g, residual=y)
        g.symm.ensure_even()
    torch.cuda.synchronize(); dist.barrier()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    fused_us = e0.elapsed_time(e1) / 640 * 1e3
    # baseline: GEMV kernel + NCCL all-reduce + add, also in a graph
    graph2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph2, stream=s):
        y2 = res
        for _ in range(64):
            t = ops.linear(x, w)
            dist.all_reduce(t)
            y2 = t + y2
    torch.cuda.synchronize(); dist.barrier()
    for _ in range(3):
        graph2.replay()
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(10):
        graph2.replay()
    e1.record(); torch.cuda.synchronize()
    nccl_us = e0.elapsed_time(e1) / 640 * 1e3
    t = torch.tensor([fused_us, nccl_us], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f'{{"world": {world}, "fused_gemv_allreduce_us": {t[0].item():.2f}, "gemv_plus_nccl_allreduce_us": {t[1].item():.2f}, "worst_rel_err": {worst:.3e}, "ok": true}}', flush=True)
    dist.barrier()
    _hard_exit()


if __name__ == "__main__":
    main()
