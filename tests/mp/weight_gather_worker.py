"""gloo worker: weight-gathered matmuls (tiled over K, looped over N) == the dense product on every rank."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    from neuronx_distributed_inference_b200.modules.eagle.utils import WeightGatheredColumnParallel, looped_einsum, tiled_all_gather_matmul
    from neuronx_distributed_inference_b200.parallel import state as pstate
    pstate.init_distributed("gloo")
    world = int(os.environ["WORLD_SIZE"])
    pstate.initialize_model_parallel(tensor_model_parallel_size=world)
    g = pstate.get_tensor_model_parallel_group()
    torch.manual_seed(0)
    N, K = 12 * world, 40
    W, b, x = torch.randn(N, K), torch.randn(N), torch.randn(2, 5, K)
    shard = W[g.rank * 12:(g.rank + 1) * 12].contiguous()
    ref = x @ W.t()
    ok = torch.allclose(tiled_all_gather_matmul(x, shard, g, tile=16), ref, atol=1e-4)
    ok = ok and torch.allclose(looped_einsum(x, shard, g, loops=3), ref, atol=1e-4)
    layer = WeightGatheredColumnParallel(K, N, bias=True, gather_output=False, dtype=torch.float32)
    layer.weight.copy_(shard)
    layer.bias.copy_(b[g.rank * 12:(g.rank + 1) * 12])
    ok = ok and torch.allclose(layer.forward_wg(x, tile=16), ref + b, atol=1e-4)
    ok = ok and torch.allclose(layer(x), (ref + b)[..., g.rank * 12:(g.rank + 1) * 12], atol=1e-4)
    flags = [torch.tensor([1.0 if ok else 0.0]) for _ in range(world)]
    import torch.distributed as dist
    dist.all_gather(flags, torch.tensor([1.0 if ok else 0.0]))
    if g.rank == 0:
        print('{"ok": %s}' % ("true" if all(f.item() == 1.0 for f in flags) else "false"), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
