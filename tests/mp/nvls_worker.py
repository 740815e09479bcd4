"""torchrun worker: symmetric heap (VMM + NVLS multicast) and the in-switch collectives vs NCCL (numerics, graph replay, timing)."""
import faulthandler
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
faulthandler.dump_traceback_later(int(os.environ.get("DUMP_AFTER", "120")), exit=True)


def log(*a):
    print(f"[r{os.environ.get('RANK')}]", *a, flush=True)


def main():
    from neuronx_distributed_inference_b200.parallel import mappings, state as pstate
    from neuronx_distributed_inference_b200.parallel.symm import SymmetricWorkspace
    from neuronx_distributed_inference_b200.parallel.symm_heap import SymmetricHeap
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    pstate.init_distributed("nccl")
    pstate.initialize_model_parallel(tensor_model_parallel_size=world)
    g = pstate.get_tensor_model_parallel_group()
    dev = torch.device("cuda", torch.cuda.current_device())
    g.symm = SymmetricWorkspace.create(g, dev, max_width=4096)
    heap = SymmetricHeap(g, dev, 256 << 20)
    log("heap ok: size", heap.size, "multicast", heap.has_multicast)
    assert heap.has_multicast, "NVLS multicast expected on an NVSwitch box"
    torch.manual_seed(100 + rank)
    worst = 0.0

    def rel(a, b):
        return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-6)).item()
    for shape, dim in [((256, 4096), 0), ((2, 128, 4096), 1), ((1, 64 * world, 1024), 1), ((2048, 4096), 0)]:
        g.symm.begin_step()
        x = torch.randn(*shape, device=dev, dtype=torch.bfloat16)
        res = torch.randn(*shape, device=dev, dtype=torch.bfloat16)
        dist.broadcast(res, 0)
        ref = x.float().clone()
        dist.all_reduce(ref)
        # all-reduce (+ residual)
        y = heap.all_reduce(x, res)
        e = rel(y, ref + res.float()); worst = max(worst, e)
        ys = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(ys, y)
        same = all(torch.equal(ys[0], t) for t in ys)
        # reduce-scatter along dim (+ residual shard)
        rshard = res.chunk(world, dim)[rank].contiguous()
        z = heap.reduce_scatter(x, dim, rshard)
        e2 = rel(z, (ref + res.float()).chunk(world, dim)[rank]); worst = max(worst, e2)
        # all-gather along dim
        xs = x.chunk(world, dim)[rank].contiguous()
        full = heap.all_gather(xs, dim).clone()
        parts = [torch.empty_like(xs) for _ in range(world)]
        dist.all_gather(parts, xs)
        ok_ag = torch.equal(full, torch.cat(parts, dim))
        log(f"shape {shape} dim {dim}: all_reduce rel {e:.2e} identical {same}; reduce_scatter rel {e2:.2e}; all_gather exact {ok_ag}")
        assert e < 1e-2 and e2 < 1e-2 and same and ok_ag
    # through the mappings dispatch (what the layers call)
    g.heap = heap
    g.symm.begin_step()
    x = torch.randn(2, 64 * world, 512, device=dev, dtype=torch.bfloat16)
    ref = x.float().clone(); dist.all_reduce(ref)
    assert rel(mappings.all_reduce(x, g), ref) < 1e-2
    assert rel(mappings.reduce_scatter(x, 1, g), ref.chunk(world, 1)[rank]) < 1e-2
    # fused GEMM -> reduce-scatter (one kernel) vs GEMM + NCCL reduce-scatter
    from neuronx_distributed_inference_b200 import ops
    for (Bb, T, K, N) in [(1, 128 * world, 512, 1024), (2, 256 * world, 1792, 4096), (1, 2048, 2048, 4096)]:
        if T % (128 * world):
            continue
        g.symm.begin_step()
        x = torch.randn(Bb, T, K, device=dev, dtype=torch.bfloat16) * 0.5
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        res = torch.randn(Bb, T // world, N, device=dev, dtype=torch.bfloat16)
        assert heap.gemm_rs_usable(x, w, T)
        for it in range(3):      # repeated calls alternate the staging buffers and reuse the flag array
            y = heap.gemm_reduce_scatter(x.reshape(-1, K), w, None, T, res).view(Bb, T // world, N)
        ref = x.float() @ w.float().t()
        dist.all_reduce(ref)
        ref = ref.view(Bb, world, T // world, N)[:, rank] + res.float()
        e = rel(y, ref); worst = max(worst, e)
        log(f"fused gemm+reduce_scatter B={Bb} T={T} K={K} N={N}: rel {e:.2e}")
        assert e < 1e-2
    # fused GEMM -> all-reduce (reduce owned tiles in the switch, multicast the result): one kernel
    for (M, K, N) in [(128 * world, 512, 1024), (2048, 1792, 4096)]:
        g.symm.begin_step()
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16) * 0.5
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        res = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        dist.broadcast(res, 0)
        assert heap.gemm_ar_usable(x, w)
        for it in range(3):
            y = heap.gemm_all_reduce(x, w, None, res)
        ref = x.float() @ w.float().t()
        dist.all_reduce(ref)
        e = rel(y, ref + res.float()); worst = max(worst, e)
        ys = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(ys, y)
        same = all(torch.equal(ys[0], t) for t in ys)
        log(f"fused gemm+all_reduce M={M} K={K} N={N}: rel {e:.2e} identical {same}")
        assert e < 1e-2 and same
    # timing: fused vs GEMM + NVLS reduce-scatter kernel vs GEMM + NCCL
    out = {}
    Bb, T, K, N = 1, 2048, 1792, 4096
    x = torch.randn(Bb, T, K, device=dev, dtype=torch.bfloat16) * 0.5
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    res = torch.randn(Bb, T // world, N, device=dev, dtype=torch.bfloat16)
    def t_fused():
        return heap.gemm_reduce_scatter(x.reshape(-1, K), w, None, T, res)
    def t_nvls():
        y = ops.linear(x, w, None, out=ops.staging_for(g, x, w))
        return heap.reduce_scatter(y, 1, res)
    def t_nccl():
        y = ops.linear(x, w, None)
        o = torch.empty(Bb * T // world, N, device=dev, dtype=torch.bfloat16)
        dist.reduce_scatter_tensor(o, y.view(-1, N))
        return o + res.view(-1, N)
    xf = x.reshape(-1, K)
    resf = torch.randn(Bb * T, N, device=dev, dtype=torch.bfloat16)
    def t_fused_ar():
        return heap.gemm_all_reduce(xf, w, None, resf)
    def t_nccl_ar():
        y = ops.linear(xf, w, None)
        dist.all_reduce(y)
        return y + resf
    for name, fn in (("fused", t_fused), ("gemm_plus_nvls", t_nvls), ("gemm_plus_nccl", t_nccl), ("fused_allreduce", t_fused_ar),
                     ("gemm_plus_nccl_allreduce", t_nccl_ar)):
        g.symm.begin_step()
        for _ in range(3): fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.symm.begin_step()
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) / 20 * 1e3], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        out[f"rowparallel_2048x1792x4096_{name}_us"] = round(float(tt.item()), 2)
    # graph capture + replay, timing vs NCCL
    for (M, N) in [(256, 4096), (2048, 4096)]:
        x = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            g.symm.begin_step(); heap.all_reduce(x); torch.cuda.synchronize()
            with torch.cuda.graph(gr, stream=s):
                g.symm.begin_step()
                for _ in range(16):
                    y = heap.all_reduce(x)
        gr2 = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            with torch.cuda.graph(gr2, stream=s):
                for _ in range(16):
                    t = x.clone(); dist.all_reduce(t)
        def bench(gg):
            torch.cuda.synchronize(); dist.barrier()
            for _ in range(3): gg.replay()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): gg.replay()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 160 * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        out[f"allreduce_{M}x{N}_nvls_us"] = round(bench(gr), 2)
        out[f"allreduce_{M}x{N}_nccl_us"] = round(bench(gr2), 2)
    if rank == 0:
        print(json.dumps(dict(world=world, worst_rel_err=worst, ok=True, **out)), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
