"""Worker run under torchrun (gloo on CPU, nccl on GPU): tensor-parallel Llama vs the HF fp32 reference.
usage: torchrun --nproc-per-node N tests/mp/llama_tp_worker.py <ckpt_dir> <cpu|cuda> [dtype]"""
import faulthandler
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


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
    """``<ckpt>`` and ``MODEL_TYPE`` may be comma-separated lists of equal length: the families run one after the other in the same
    process group (saves the interpreter + import time of one torchrun per family)."""
    faulthandler.dump_traceback_later(int(os.environ.get('DUMP_AFTER', '240')), exit=True)
    ckpts, dev = sys.argv[1].split(","), sys.argv[2]
    dtype = sys.argv[3] if len(sys.argv) > 3 else ("float32" if dev == "cpu" else "bfloat16")
    types = os.environ.get("MODEL_TYPE", "llama").split(",")
    assert len(types) == len(ckpts)
    for ckpt, mt in zip(ckpts, types):
        run_one(ckpt, dev, dtype, mt)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.distributed.barrier()
        _hard_exit()


def run_one(ckpt, dev, dtype, model_type):
    from neuronx_distributed_inference_b200.config import NeuronConfig, OnDeviceSamplingConfig, load_pretrained_config
    from neuronx_distributed_inference_b200.utils.constants import get_model_cls
    app_cls = get_model_cls(model_type)          # any registered causal-lm family
    from neuronx_distributed_inference_b200.parallel import state as pstate
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if dev == "cuda":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    nc = app_cls.get_neuron_config_cls()(batch_size=2, seq_len=64, max_context_length=32, torch_dtype=dtype, tp_degree=world,
                      on_cpu=(dev == "cpu"), output_logits=True, on_device_sampling_config=OnDeviceSamplingConfig(top_k=1),
                      flash_decoding_enabled=os.environ.get("FLASH_DECODING", "0") == "1",
                      sequence_parallel_enabled=os.environ.get("SEQUENCE_PARALLEL", "0") == "1",
                      rolling_sliding_window_cache=os.environ.get("ROLLING_SWA", "0") == "1",
                      attention_dp_degree=int(os.environ.get("ATTENTION_DP", "1")), cp_degree=int(os.environ.get("CP", "1")),
                      strided_context_parallel_kernel_enabled=os.environ.get("STRIDED_CP", "0") == "1",
                      is_continuous_batching=int(os.environ.get("ATTENTION_DP", "1")) > 1)
    cfg = app_cls.get_config_cls()(nc, load_config=load_pretrained_config(ckpt))
    app = app_cls(ckpt, cfg)
    app.load(None, skip_warmup=True)
    if nc.flash_decoding_enabled:
        kvg = app.model.layers[0].self_attn.kv_group
        assert kvg is not None and kvg.size == world // cfg.num_key_value_heads, "flash decoding group not formed"
        assert app.model.kv_mgr.max_len == -(-nc.max_length // kvg.size), "KV cache is not sequence-sharded"
    if nc.attention_dp_degree > 1:
        a0 = app.model.layers[0].self_attn
        assert a0.dp_group is not None and a0.dp_group.size == nc.attention_dp_degree
        assert type(app.model.kv_mgr).__name__ == "DataParallelKVCacheManager" and app.model.kv_mgr.num_lines == 2 // nc.attention_dp_degree
    if nc.cp_degree > 1:
        assert app.model.layers[0].self_attn.cp_group is not None
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, cfg.vocab_size, (2, 12), generator=g)
    mask = torch.ones(2, 12, dtype=torch.int32)
    mask[1, 9:] = 0
    out = app(ids, attention_mask=mask)
    logits = [out.logits[:, 0].float().cpu()]
    toks = [out.tokens.cpu()]
    pos = mask.sum(-1).view(2, 1).int()
    tok = out.tokens.cpu()
    for _ in range(6):
        out = app(tok.view(2, 1), position_ids=pos)
        tok = out.tokens.cpu()
        logits.append(out.logits[:, 0].float().cpu())
        toks.append(tok)
        pos = pos + 1
    if rank == 0:
        from transformers import AutoModelForCausalLM
        hf = AutoModelForCausalLM.from_pretrained(ckpt, torch_dtype=torch.float32).eval()
        ok = True
        worst = 0.0
        with torch.no_grad():
            for b in range(2):
                n = int(mask[b].sum())
                seq = ids[b:b + 1, :n]
                for step in range(len(logits)):
                    ref = hf(seq).logits[0, -1]
                    got = logits[step][b]
                    err = ((got - ref).norm() / ref.norm()).item()
                    worst = max(worst, err)
                    # teacher-force OUR token so later steps stay comparable
                    seq = torch.cat([seq, toks[step][b].view(1, 1)], 1)
        tol = 2e-4 if dtype == "float32" else 4e-2
        ok = worst < tol
        print(json.dumps({"model_type": model_type, "world": world, "device": dev, "dtype": dtype, "worst_rel_err": worst, "ok": ok}), flush=True)
        if not ok:
            os._exit(1)


if __name__ == "__main__":
    main()
