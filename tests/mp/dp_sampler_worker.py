"""gloo worker: DataParallelSampler (all-to-all of vocab shards, local full-vocab sampling) == Sampler (distributed reduce)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    from neuronx_distributed_inference_b200.config import NeuronConfig, OnDeviceSamplingConfig
    from neuronx_distributed_inference_b200.modules.sampling import DataParallelSampler, Sampler, prepare_sampling_params
    from neuronx_distributed_inference_b200.parallel import state as pstate
    pstate.init_distributed("gloo")
    world = int(os.environ["WORLD_SIZE"])
    pstate.initialize_model_parallel(tensor_model_parallel_size=world)
    g = pstate.get_tensor_model_parallel_group()
    torch.manual_seed(0)
    full = torch.randn(4, 64 * world)
    shard = full[:, g.rank * 64:(g.rank + 1) * 64].contiguous()
    ok = True
    for cfg in (OnDeviceSamplingConfig(top_k=1), OnDeviceSamplingConfig(do_sample=True, dynamic=True, deterministic=True, global_topk=16)):
        nc = NeuronConfig(tp_degree=world, on_cpu=True, on_device_sampling_config=cfg)
        a, b = Sampler(nc, g, True), DataParallelSampler(nc, g, True)
        params = prepare_sampling_params(4, [1, 5, 8, 3], [1.0, 0.9, 0.5, 1.0], [1.0, 0.7, 1.3, 0.0])
        ta, tb = a(shard, params), b(shard, params)
        ok = ok and torch.equal(ta, tb)
        if cfg.top_k == 1 and not cfg.do_sample:
            ok = ok and torch.equal(ta, full.argmax(-1))
    if g.rank == 0:
        print('{"ok": %s}' % ("true" if ok else "false"), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
