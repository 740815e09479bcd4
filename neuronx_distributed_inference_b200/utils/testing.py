"""Test / benchmark helpers (role of reference utils/testing.py:20-432: build a module or a tiny model
with random weights, run it, compare with a CPU callable)."""
from __future__ import annotations

import json
import os
import tempfile
from typing import Callable, Optional

import torch


def build_random_llama(hf_overrides: dict, batch_size: int = 1, seq_len: int = 128, max_context_length: int = 64,
                       device: str = "cuda", tp_degree: int = 1, dtype="bfloat16", app_cls=None, config_cls=None,
                       skip_warmup: bool = True, seed: int = 0, fused_draft: Optional[dict] = None, **neuron_kwargs):
    """A Llama-architecture application with N(0,0.02) weights created directly on ``device``.
    ``fused_draft``: ``dict(hf=<overrides of the draft architecture>, neuron=<draft NeuronConfig overrides>)`` builds the
    application with fused speculation (draft inside the same application, reference FusedSpecNeuronConfig)."""
    from ..config import NeuronConfig, OnDeviceSamplingConfig
    from ..models.llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaForCausalLM
    app_cls = app_cls or NeuronLlamaForCausalLM
    config_cls = config_cls or app_cls.get_config_cls()
    hf = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
              num_key_value_heads=2, vocab_size=1024, max_position_embeddings=max(seq_len, 2048), rms_norm_eps=1e-5,
              rope_theta=10000.0, hidden_act="silu", tie_word_embeddings=False, pad_token_id=0)
    hf.update(hf_overrides)
    nk = dict(batch_size=batch_size, seq_len=seq_len, max_context_length=max_context_length, torch_dtype=dtype,
              tp_degree=tp_degree, on_cpu=(device == "cpu"),
              on_device_sampling_config=OnDeviceSamplingConfig(top_k=1))
    nk.update(neuron_kwargs)
    nc = config_cls.get_neuron_config_cls()(**nk)

    def load_config(cfg):
        for k, v in hf.items():
            setattr(cfg, k, v)
    fsc = None
    if fused_draft is not None:
        from ..config import FusedSpecNeuronConfig
        dhf = dict(hf)
        dhf.update(fused_draft.get("hf", {}))
        dnk = dict(nk)
        dnk.update(fused_draft.get("neuron", {}))
        for drop in ("enable_fused_speculation", "enable_eagle_speculation", "token_tree_config", "is_medusa",
                     "num_medusa_heads", "medusa_speculation_length", "medusa_tree"):
            dnk.pop(drop, None) if drop not in fused_draft.get("neuron", {}) else None
        dnc = config_cls.get_neuron_config_cls()(**dnk)

        def load_draft(c):
            for k, v in dhf.items():
                setattr(c, k, v)
        fsc = FusedSpecNeuronConfig(app_cls._model_cls, draft_config=config_cls(dnc, load_config=load_draft))
    cfg = config_cls(nc, load_config=load_config, fused_spec_config=fsc)
    app = app_cls("<random>", cfg)
    app.load(None, skip_warmup=skip_warmup, random_weights=True, seed=seed)
    return app


def perturb_constant_vectors(model, std: float = 0.1, seed: int = 1234):
    """Hugging Face initialises every norm weight to exactly 1 (or 0 for the ``1 + w`` kind) and every bias to 0; a port that forgets to
    load one of them, or applies it on the wrong side of a rotation, would still match.  Give every CONSTANT vector-shaped (1-D, or [1, n]) floating
    parameter / buffer a random perturbation so the comparison can see it."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, t in list(model.named_parameters()) + list(model.named_buffers()):
            if t.numel() == max(t.shape, default=0) and t.is_floating_point() and t.numel() > 1 and bool((t == t.flatten()[0]).all()) and float(t.flatten()[0]) in (0.0, 1.0):
                t.add_(torch.randn(t.shape, generator=g).to(t.dtype) * std)
    return model


def save_random_hf_checkpoint(hf_config, out_dir: Optional[str] = None, seed: int = 0, dtype=torch.float32,
                              perturb: Optional[bool] = None) -> str:
    """``AutoModelForCausalLM.from_config(...).save_pretrained`` — how the reference's integration tests
    make checkpoints without network access (test/integration/utils/test_utils.py:15-48).  ``perturb``: see
    :func:`perturb_constant_vectors`; default: on for the fp32 CPU comparisons, off on a CUDA box (the bf16 GPU tolerances were set
    on the plain Hugging Face initialisation).  ``B200_PERTURB_CKPT=0/1`` overrides."""
    from transformers import AutoModelForCausalLM
    torch.manual_seed(seed)
    model = AutoModelForCausalLM.from_config(hf_config).to(dtype).eval()
    if perturb is None:
        env = os.environ.get("B200_PERTURB_CKPT")
        perturb = (env == "1") if env in ("0", "1") else not torch.cuda.is_available()
    if perturb:
        perturb_constant_vectors(model)
    out_dir = out_dir or tempfile.mkdtemp(prefix="nxdi_b200_ckpt_")
    model.save_pretrained(out_dir)
    return out_dir


def validate_accuracy(fn: Callable, ref_fn: Callable, inputs, rtol: float = 1e-2, atol: float = 1e-3):
    got, exp = fn(*inputs), ref_fn(*inputs)
    torch.testing.assert_close(got.float().cpu(), exp.float().cpu(), rtol=rtol, atol=atol)
    return got


def init_cpu_env(tp_degree: int = 1):
    from ..parallel import state
    state.initialize_model_parallel(tensor_model_parallel_size=tp_degree, skip_collective_init=True)


def destroy_cpu_env():
    from ..parallel import state
    state.destroy_model_parallel()
