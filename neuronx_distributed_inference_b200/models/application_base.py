"""Host-side application objects.

``NeuronApplicationBase``: compile / load / shard weights / quantized-checkpoint generation
(reference models/application_base.py:68-822).  ``NeuronBaseForCausalLM``: owns the sub-model
runners (context encoding, token generation, speculation, fused speculation), dispatches each call
to prefill or decode, derives async next-step inputs, builds the output object
(reference models/model_base.py:3024-3945).

B200 mapping of the reference life-cycle:
    compile(path)  -> save ``neuron_config.json`` (+ optional pre-sharded per-rank safetensors);
                      nothing is traced — kernels are AOT-compiled by ``__graft_entry__.build``.
    load(path)     -> join process groups, build the module on the GPU, load+convert+shard the HF
                      checkpoint, allocate the KV cache, attach the NVLink symmetric workspace,
                      warm up (captures one CUDA graph per decode bucket).
    forward(...)   -> prefill eagerly (bucket-padded), decode by CUDA-graph replay.
"""
from __future__ import annotations

import copy
import logging
import os
import time
from typing import Dict, List, Optional, Union

import torch
import torch.nn as nn

from .. import ops
from ..config import InferenceConfig, NeuronConfig, to_torch_dtype
from ..modules import autobucketing
from ..modules.checkpoint import (load_sharded, load_state_dict, prune_state_dict, save_state_dict_safetensors,
                                  shard_state_dict)
from ..modules.sampling import prepare_sampling_params, validate_sampling_params
from ..parallel import state as pstate
from ..runtime.runner import SubModelRunner
from .model_base import ModelOutput

logger = logging.getLogger("b200infer")

CONTEXT_ENCODING_MODEL_TAG = "context_encoding_model"
TOKEN_GENERATION_MODEL_TAG = "token_generation_model"
SPECULATION_MODEL_TAG = "speculation_model"
MEDUSA_MODEL_TAG = "medusa_speculation_model"
FUSED_SPECULATION_MODEL_TAG = "fused_speculation_model"
VISION_ENCODER_MODEL_TAG = "vision_encoder_model"


class CausalLMOutput:
    """Mirror of ``CausalLMOutputWithPast`` + the reference's extra attributes
    (model_base.py:3821-3862)."""

    def __init__(self, logits=None, tokens=None, hidden_states=None, fused_outputs=None, past_key_values=None,
                 captured_tensors=None, medusa_tokens=None):
        self.logits = logits
        self.tokens = tokens
        self.hidden_states = hidden_states
        self.fused_outputs = fused_outputs
        self.past_key_values = past_key_values if past_key_values is not None else []
        self.captured_tensors = captured_tensors
        self.medusa_tokens = medusa_tokens
        self.async_should_stop = False


class NeuronApplicationBase(nn.Module):
    _model_cls = None
    _STATE_DICT_MODEL_PREFIX = "model."
    _NEW_STATE_DICT_MODEL_PREFIX = ""

    def __init__(self, model_path: str, config: Optional[InferenceConfig] = None,
                 neuron_config: Optional[NeuronConfig] = None):
        super().__init__()
        if config is None:
            config = self.get_config_cls().load(model_path)
        if neuron_config is not None:
            config.neuron_config = neuron_config
        self.validate_config(config)
        self.config = config
        self.neuron_config = config.neuron_config
        self.model_path = model_path
        self.is_compiled = False
        self.is_loaded_to_neuron = False
        self.models: List[SubModelRunner] = []
        self.model = None          # the device nn.Module shared by all runners
        self.device = None
        self._external_state_dict = None

    # ---- class hooks ---------------------------------------------------------------------------
    @classmethod
    def get_config_cls(cls):
        return InferenceConfig

    @classmethod
    def get_neuron_config_cls(cls):
        return cls.get_config_cls().get_neuron_config_cls()

    @classmethod
    def validate_config(cls, config):
        if not hasattr(config, "neuron_config") or config.neuron_config is None:
            raise ValueError("config.neuron_config is required")

    @staticmethod
    def load_hf_model(model_path):
        from transformers import AutoModelForCausalLM
        return AutoModelForCausalLM.from_pretrained(model_path)

    @staticmethod
    def convert_hf_to_neuron_state_dict(state_dict: dict, config: InferenceConfig) -> dict:
        return state_dict

    @staticmethod
    def update_state_dict_for_tied_weights(state_dict):
        pass

    # ---- device / groups ------------------------------------------------------------------------
    def _init_runtime(self):
        nc = self.neuron_config
        if nc.on_cpu or not torch.cuda.is_available():
            self.device = torch.device("cpu")
            if nc.tp_degree > 1:
                pstate.init_distributed("gloo")
        else:
            pstate.init_distributed("nccl")
            self.device = torch.device("cuda", torch.cuda.current_device())
        if not pstate.model_parallel_is_initialized() or pstate.get_tensor_model_parallel_size() != nc.tp_degree:
            ep = getattr(nc, "moe_ep_degree", 1) or 1
            pstate.initialize_model_parallel(
                tensor_model_parallel_size=nc.tp_degree, pipeline_model_parallel_size=nc.pp_degree,
                expert_model_parallel_size=ep, context_parallel_size=nc.cp_degree,
                attention_dp_size=nc.attention_dp_degree, moe_tp_size=getattr(nc, "moe_tp_degree", None),
                kv_shared_size=(getattr(self.config.get_text_config(), "num_cores_per_group", 1)
                                if (nc.flash_decoding_enabled or nc.attention_dp_degree > 1 or nc.cp_degree > 1) else 1))

    # ---- compile / load -----------------------------------------------------------------------
    def compile(self, compiled_model_path: str, debug: bool = False, pre_shard_weights_hook=None,
                dry_run: bool = False):
        """Persist the artifact directory.  Nothing is traced on B200; with
        ``save_sharded_checkpoint`` the per-rank weight shards are written so ``load`` is a plain
        file read (reference application_base.py:240-265,292-316)."""
        os.makedirs(compiled_model_path, exist_ok=True)
        self.config.save(compiled_model_path)
        if dry_run:
            return
        if self.neuron_config.save_sharded_checkpoint:
            self.shard_weights(compiled_model_path)
        self.is_compiled = True

    def shard_weights(self, compiled_model_path: str):
        self._init_runtime()
        model = self._build_module(torch.device("meta"))
        sd = self.checkpoint_loader_fn()
        wdir = os.path.join(compiled_model_path, "weights")
        g = pstate.get_tensor_model_parallel_group()
        ranks = range(self.neuron_config.tp_degree) if g.size == 1 else [g.rank]
        for r in ranks:
            shard = shard_state_dict(model, sd, strict=False, rank_override=r if g.size == 1 else None)
            save_state_dict_safetensors(shard, os.path.join(wdir, f"tp{r}_sharded_checkpoint"))

    def _build_module(self, device):
        nc = self.neuron_config
        prev = torch.get_default_dtype()
        try:
            with torch.device(device):
                model = self._model_cls(self.config, device=device)
        finally:
            torch.set_default_dtype(prev)
        return model.eval()

    def load(self, compiled_model_path: Optional[str] = None, start_rank_id=None, local_ranks_size=None,
             skip_warmup: bool = False, state_dict: Optional[dict] = None, random_weights: bool = False,
             seed: int = 0):
        """Build the model on this rank's device and load weights (``random_weights``: N(0, 0.02) init
        directly on the device — benchmarks / smoke tests without a checkpoint)."""
        nc = self.neuron_config
        self._init_runtime()
        t0 = time.time()
        self.model = self._build_module(self.device)
        if self.neuron_config.quantized:
            from ..quantization.convert import convert
            convert(self.model, nc)
        self._post_build(self.model)
        pre = None
        if compiled_model_path is not None:
            g = pstate.get_tensor_model_parallel_group()
            p = os.path.join(compiled_model_path, "weights", f"tp{g.rank}_sharded_checkpoint")
            if os.path.isdir(p) and nc.save_sharded_checkpoint:
                pre = load_state_dict(p)
        if random_weights:
            self.init_random_weights(seed)
        elif pre is not None:
            self.model.load_state_dict(pre, strict=False)
        else:
            sd = state_dict if state_dict is not None else self.checkpoint_loader_fn()
            load_sharded(self.model, sd, nc.torch_dtype, strict=False)
            rep = self.load_report = getattr(self.model, "_load_report", {"missing": [], "unexpected": []})
            if rep["missing"] or rep["unexpected"]:     # a port that forgot a tensor must not go unnoticed (the values stay at their init)
                logger.warning("checkpoint load: %d model tensors not in the checkpoint %s; %d checkpoint tensors unused %s",
                               len(rep["missing"]), rep["missing"][:6], len(rep["unexpected"]), rep["unexpected"][:6])
                if rep["missing"] and os.environ.get("B200_STRICT_LOAD") == "1":      # the test-suite setting
                    raise KeyError(f"model tensors missing from the converted checkpoint: {rep['missing'][:12]}")
        self._post_load(self.model)
        self._attach_symmetric_workspace()
        self._build_speculation(random_weights, seed)
        self._build_runners()
        self.is_loaded_to_neuron = True
        logger.info("model loaded in %.1fs on %s", time.time() - t0, self.device)
        if not (skip_warmup or nc.skip_warmup):
            self.warmup()
        return self

    def init_random_weights(self, seed: int = 0):
        gen = torch.Generator(device=self.device)
        gen.manual_seed(seed + 1000 * pstate.get_tensor_model_parallel_group().rank)
        with torch.no_grad():
            for name, p in self.model.named_parameters():
                if p.dim() == 1 and ("norm" in name or "layernorm" in name):
                    p.fill_(1.0)
                elif p.dtype in (torch.int8, torch.float8_e4m3fn, torch.float8_e5m2):
                    p.copy_((torch.randn(p.shape, device=self.device, generator=gen) * 20).clamp(-100, 100).to(p.dtype))
                elif name.endswith(".scale"):
                    p.fill_(0.001)
                elif p.is_floating_point():
                    tmp = torch.empty(p.shape, device=self.device, dtype=torch.float32 if p.numel() < (1 << 28)
                                      else p.dtype)
                    tmp.normal_(0.0, 0.02, generator=gen)
                    p.copy_(tmp)
                    del tmp
            for m in self.model.modules():      # heads stored zero-padded to a kernel width (modules/gqa.py)
                if hasattr(m, "zero_head_padding"):
                    m.zero_head_padding()

    def to_cpu(self):
        """CPU execution path (reference application_base.py:556-628)."""
        self.neuron_config.on_cpu = True
        return self.load(None, skip_warmup=True)

    def _post_build(self, model):
        pass

    def _post_load(self, model):
        lc = self.neuron_config.lora_config
        if lc is not None and hasattr(model, "layers"):
            from ..modules.lora import LoraModel, LoraModelManager
            model.lora = LoraModel(model, lc, device=self.device)
            self.lora_manager = LoraModelManager(model.lora, lc)

    def _build_speculation(self, random_weights: bool = False, seed: int = 0):
        pass

    def _attach_symmetric_workspace(self):
        nc = self.neuron_config
        # W8A8: dynamic per-token fp8 activations into the tcgen05 kind::f8f6f4 GEMM (prefill-sized token counts); the reference's
        # quantized_mlp_kernel_enabled / activation_quantization_type="dynamic" / rmsnorm_quantize_kernel_enabled switches
        # (models/config.py:219-242 of the reference)
        if self.device.type == "cuda" and (nc.quantized_mlp_kernel_enabled or nc.rmsnorm_quantize_kernel_enabled
                                           or nc.activation_quantization_type == "dynamic"):
            ops.set_activation_quant(True)
        g = pstate.get_tensor_model_parallel_group()
        if g.size > 1 and self.device.type == "cuda" and nc.fused_collectives and g.symm is None:
            from ..parallel.symm import SymmetricWorkspace
            hidden = getattr(self.config, "hidden_size", 8192)
            g.symm = SymmetricWorkspace.create(g, self.device, max_tokens=ops.GEMV_MAX_TOKENS, max_width=hidden)
        if (g.size > 1 and self.device.type == "cuda" and nc.fused_collectives and g.symm is not None and getattr(g, "heap", None) is None
                and os.environ.get("NXDI_B200_SYMM_HEAP", "1") != "0"):
            # prefill-sized collectives: VMM symmetric heap + NVLS multicast (in-switch all-reduce / reduce-scatter / all-gather)
            from ..parallel.symm_heap import SymmetricHeap
            try:
                g.heap = SymmetricHeap(g, self.device, int(os.environ.get("NXDI_B200_SYMM_HEAP_MB", "512")) << 20)
                if not g.heap.has_multicast:
                    logger.warning("symmetric heap: NVLS multicast not available on this system; prefill collectives stay on NCCL")
            except Exception as e:      # no fd passing / no VMM support: keep NCCL
                logger.warning("symmetric heap unavailable (%s); prefill collectives stay on NCCL", str(e).split("\n")[0])
                g.heap = None

    def _build_runners(self):
        raise NotImplementedError

    def warmup(self):
        for r in self.models:
            try:
                r.warmup()
            except RuntimeError as e:  # reference swallows per-bucket warmup errors (:356-371)
                logger.warning("warmup of %s failed: %s", r.tag, e)

    # ---- weights ---------------------------------------------------------------------------------
    def checkpoint_loader_fn(self, mmap: bool = False) -> dict:
        """Full (unsharded) converted state dict (reference application_base.py:630-683)."""
        nc = self.neuron_config
        path = self.model_path
        if nc.quantized:
            path = nc.quantized_checkpoints_path
            sd = load_state_dict(path)
            sd = {self._strip(k): v for k, v in sd.items()}
            sd = {(k[:-len("weight_scale")] + "scale" if k.endswith("weight_scale") else k): v for k, v in sd.items()}
            return sd
        return self.get_state_dict(path, self.config)

    @classmethod
    def _strip(cls, k: str) -> str:
        p = cls._STATE_DICT_MODEL_PREFIX
        if p and k.startswith(p):
            return cls._NEW_STATE_DICT_MODEL_PREFIX + k[len(p):]
        return k

    @classmethod
    def get_state_dict(cls, model_name_or_path: str, config: InferenceConfig) -> dict:
        """HF checkpoint -> our naming (reference application_base.py:692-739)."""
        if os.path.isdir(model_name_or_path) or os.path.isfile(model_name_or_path):
            sd = load_state_dict(model_name_or_path)
        else:
            sd = cls.load_hf_model(model_name_or_path).state_dict()
        sd = {cls._strip(k): v for k, v in sd.items()}
        sd = {(k[:-len("weight_scale")] + "scale" if k.endswith("weight_scale") else k): v for k, v in sd.items()}
        sd = cls.convert_hf_to_neuron_state_dict(sd, config)
        if getattr(config, "tie_word_embeddings", False):
            cls.update_state_dict_for_tied_weights(sd)
        return sd

    # ---- quantized checkpoint generation ----------------------------------------------------
    @classmethod
    def save_quantized_state_dict(cls, model_path: str, config: InferenceConfig):
        """Quantise the HF weights offline and store them at ``quantized_checkpoints_path``
        (reference application_base.py:746-799)."""
        from ..quantization.convert import quantize_state_dict
        nc = config.neuron_config
        sd = cls.get_state_dict(model_path, config)
        qsd = quantize_state_dict(sd, nc)
        out = nc.quantized_checkpoints_path
        if out.endswith(".pt"):
            os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
            torch.save(prune_state_dict(qsd), out)
        else:
            save_state_dict_safetensors(prune_state_dict(qsd), out)
        return qsd

    generate_quantized_state_dict = save_quantized_state_dict

    def reset(self):
        if self.model is not None:
            self.model.reset()
        for r in self.models:
            r.reset()


class NeuronBaseForCausalLM(NeuronApplicationBase):
    """CTE / TKG / speculation dispatch."""

    def __init__(self, model_path: str, config=None, neuron_config=None):
        super().__init__(model_path, config, neuron_config)
        self.text_config = self.config.get_text_config()
        self.vocab_size = getattr(self.text_config, "vocab_size", None)
        self.padding_side = self.neuron_config.padding_side
        self.kv_cache_populated = False
        self.sampler = None
        self.context_encoding_model: Optional[SubModelRunner] = None
        self.token_generation_model: Optional[SubModelRunner] = None
        self.speculation_model: Optional[SubModelRunner] = None
        self._async_prev = None

    # ---- runners -----------------------------------------------------------------------------
    def _build_runners(self):
        nc = self.neuron_config
        self.models = []
        self.enable_context_encoding()
        self.enable_token_generation()
        if nc.speculation_length > 0 and not nc.enable_fused_speculation:
            self.enable_speculation()

    def enable_context_encoding(self):
        nc = self.neuron_config
        buckets = autobucketing.generate_buckets_for_cte(self.config)
        self.context_encoding_model = SubModelRunner(
            CONTEXT_ENCODING_MODEL_TAG, self.model, self.config, batch_size=nc.ctx_batch_size, buckets=buckets,
            n_active_tokens=nc.max_context_length, is_prefill=True, device=self.device)
        self.models.append(self.context_encoding_model)

    def enable_token_generation(self):
        nc = self.neuron_config
        buckets = autobucketing.generate_buckets_for_tkg(self.config)
        self.token_generation_model = SubModelRunner(
            TOKEN_GENERATION_MODEL_TAG, self.model, self.config, batch_size=nc.tkg_batch_size, buckets=buckets,
            n_active_tokens=1, is_prefill=False, device=self.device)
        self.models.append(self.token_generation_model)

    def enable_speculation(self):
        nc = self.neuron_config
        buckets = autobucketing.generate_buckets_for_speculation(self.config)
        self.speculation_model = SubModelRunner(
            SPECULATION_MODEL_TAG, self.model, self.config, batch_size=nc.spec_batch_size, buckets=buckets,
            n_active_tokens=nc.speculation_length, is_prefill=False, device=self.device)
        self.models.append(self.speculation_model)

    # ---- speculation ------------------------------------------------------------------------
    def _build_speculation(self, random_weights: bool = False, seed: int = 0):
        """Fused speculation (draft + target in one device step; reference ``enable_fused_spec`` model_base.py:3078-3105 and
        the ``draft_model.`` / ``target_model.`` checkpoint split :3843-3900), EAGLE / EAGLE-3 / token tree, Medusa."""
        nc = self.neuron_config
        self.fused_spec_model = None
        self.medusa_model = None
        self.draft_model = None
        if nc.is_medusa:
            from ..generation.medusa import MedusaSpeculativeModel
            self.medusa_model = MedusaSpeculativeModel(self.model, nc.medusa_tree)
        fsc = getattr(self.config, "fused_spec_config", None)
        if not nc.enable_fused_speculation or fsc is None:
            return
        dcfg = fsc.draft_config
        draft_cls = fsc.draft_model_cls or fsc.worker_cls or self._model_cls
        with torch.device(self.device):
            draft = draft_cls(dcfg, device=self.device).eval()
        if random_weights:
            main, self.model = self.model, draft
            try:
                self.init_random_weights(seed + 17)
            finally:
                self.model = main
        else:
            sd = self._load_draft_state_dict(fsc, dcfg)
            load_sharded(draft, sd, dcfg.neuron_config.torch_dtype, strict=False)
        if dcfg.neuron_config.is_eagle_draft and not getattr(dcfg, "draft_vocab_size", None):
            # EAGLE heads reuse the target's embedding and output projection (they ship without their own)
            with torch.no_grad():
                if random_weights or not getattr(self, "_draft_has_embed", False):
                    draft.embed_tokens.load_state_dict(self.model.embed_tokens.state_dict())
                if random_weights or not getattr(self, "_draft_has_lm_head", False):
                    draft.lm_head.load_state_dict(self.model.lm_head.state_dict())
        self.draft_model = draft
        k = nc.speculation_length
        if nc.enable_eagle_speculation:
            from ..generation.eagle import EagleSpeculativeModel
            tree = None
            if nc.token_tree_config is not None:
                from ..modules.eagle.token_tree import TokenTree
                tree = TokenTree(nc.token_tree_config)
            self.fused_spec_model = EagleSpeculativeModel(self.model, draft, k, nc.max_batch_size, tree)
        else:
            from ..generation.speculative import FusedSpeculativeModel
            self.fused_spec_model = FusedSpeculativeModel(self.model, draft, k)

    def _load_draft_state_dict(self, fsc, dcfg) -> dict:
        from ..modules.checkpoint import load_state_dict as _load
        sd = _load(fsc.draft_model_path)
        sd = {self._strip(k): v for k, v in sd.items()}
        self._draft_has_embed = any(k.startswith("embed_tokens.") for k in sd)
        self._draft_has_lm_head = any(k.startswith("lm_head.") for k in sd)
        sd = self.convert_hf_to_neuron_state_dict(sd, dcfg)
        if getattr(dcfg, "tie_word_embeddings", False) and "lm_head.weight" not in sd and "embed_tokens.weight" in sd:
            self.update_state_dict_for_tied_weights(sd)
        return sd

    # ---- forward -----------------------------------------------------------------------------
    def _infer_attention_mask(self, position_ids):
        """mask[b, j] = j <= max position of row b (reference model_base.py:3485-3504)."""
        S = int(position_ids.max().item()) + 1
        ar = torch.arange(S, device=position_ids.device).unsqueeze(0)
        return (ar <= position_ids.max(-1, keepdim=True).values).to(torch.int32)

    def preprocess_inputs(self, input_ids, attention_mask, position_ids, seq_ids, sampling_params):
        nc = self.neuron_config
        B = input_ids.shape[0]
        if position_ids is None:
            if attention_mask is not None and attention_mask.shape[-1] == input_ids.shape[-1]:
                position_ids = (attention_mask.long().cumsum(-1) - 1).clamp_min(0)
            else:
                position_ids = torch.arange(input_ids.shape[-1]).unsqueeze(0).expand(B, -1)
        if attention_mask is None and input_ids.shape[-1] > 1 and position_ids.device.type == "cpu":
            attention_mask = self._infer_attention_mask(position_ids)   # decode (T == 1) never needs a mask
        if seq_ids is None:
            seq_ids = torch.arange(B, dtype=torch.int32)
        if sampling_params is None:
            c = nc.on_device_sampling_config
            sampling_params = prepare_sampling_params(B, c.top_k, c.top_p, c.temperature) if c is not None \
                else prepare_sampling_params(B)
        elif nc.on_device_sampling_config is not None and sampling_params.device.type == "cpu":
            validate_sampling_params(sampling_params, nc.on_device_sampling_config)
        return input_ids, attention_mask, position_ids, seq_ids, sampling_params

    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None, seq_ids: Optional[torch.Tensor] = None,
                sampling_params: Optional[torch.Tensor] = None, prev_hidden=None, adapter_ids=None,
                slot_mapping=None, block_table=None, full_context_lens=None, computed_context_lens=None,
                vision_embeddings=None, vision_mask=None, rotary_position_ids=None, output_logits=None,
                return_dict: bool = True, **kwargs) -> CausalLMOutput:
        """Same calling convention as the reference ``NeuronBaseForCausalLM.forward``
        (model_base.py:3314-3460).  Tensors may live on the host (copied through pinned staging
        buffers) or already on the device."""
        input_ids, attention_mask, position_ids, seq_ids, sampling_params = self.preprocess_inputs(
            input_ids, attention_mask, position_ids, seq_ids, sampling_params)
        nc = self.neuron_config
        T = input_ids.shape[-1]
        extra = dict(prev_hidden=prev_hidden, adapter_ids=adapter_ids, slot_mapping=slot_mapping,
                     block_table=block_table, vision_embeddings=vision_embeddings, vision_mask=vision_mask,
                     rotary_position_ids=rotary_position_ids, output_logits=output_logits)
        extra.update(kwargs)
        if computed_context_lens is not None:
            extra["has_prefix"] = bool((computed_context_lens > 0).any())
        is_prefill = T > 1 and self._is_prefill(position_ids, computed_context_lens)
        if getattr(self, "fused_spec_model", None) is not None and adapter_ids is None:
            return self._fused_speculation_forward(input_ids, attention_mask, position_ids, seq_ids, sampling_params, is_prefill)
        w = nc.windowed_context_encoding_size
        if is_prefill and w and T > w and computed_context_lens is None:
            return self._windowed_context_encoding(input_ids, attention_mask, position_ids, seq_ids, sampling_params, w, extra)
        if is_prefill:
            out = self.context_encoding_model(input_ids, attention_mask, position_ids, seq_ids, sampling_params, **extra)
            self.kv_cache_populated = True
        elif T == nc.speculation_length and self.speculation_model is not None and T > 1:
            out = self.speculation_model(input_ids, attention_mask, position_ids, seq_ids, sampling_params, **extra)
        else:
            out = self.token_generation_model(input_ids, attention_mask, position_ids, seq_ids, sampling_params, **extra)
        return self._construct_output(out)

    def _fused_speculation_forward(self, input_ids, attention_mask, position_ids, seq_ids, sampling_params, is_prefill):
        """``forward`` of an application built with ``enable_fused_speculation`` (reference model_base.py:2857-3021,3821-3862):
        the prompt goes through both models; every later call takes the last accepted token ``[B,1]`` at ``position_ids`` and
        returns ``tokens`` = accepted tokens padded with -1 ``[B,k]`` plus
        ``fused_outputs = [accepted_tokens, next_input_ids, None, next_position_ids, n_accepted]`` (device tensors)."""
        dev = self.device
        fused = self.fused_spec_model
        ids = input_ids.to(dev)
        pos = position_ids.to(dev, torch.int32)
        sid = seq_ids.to(dev, torch.int32)
        if is_prefill:
            mask = attention_mask.to(dev) if attention_mask is not None else torch.ones_like(ids)
            out = fused.prefill(ids, mask, pos, sid, sampling_params.to(dev) if sampling_params is not None else None)
            n = mask.long().sum(-1, keepdim=True)
            self._spec_prev = ids.gather(1, (n - 1).clamp_min(0))
            self.kv_cache_populated = True
            return self._construct_output(out)
        from ..generation.speculative import FusedSpeculativeModel
        if isinstance(fused, FusedSpeculativeModel) and sampling_params is not None:
            # do_sample: the draft samples and the target verifies by rejection sampling (speculative.py::forward_sampling)
            accepted, n_acc, nxt, npos, prev, out_t = fused(ids, pos, sid, self._spec_prev, sampling_params.to(dev))
        else:
            accepted, n_acc, nxt, npos, prev, out_t = fused(ids, pos, sid, self._spec_prev)
        self._spec_prev = prev
        res = CausalLMOutput(tokens=accepted, logits=out_t.logits, hidden_states=None)
        res.fused_outputs = [accepted, nxt, None, npos, n_acc]
        return res

    def _windowed_context_encoding(self, input_ids, attention_mask, position_ids, seq_ids, sampling_params, w, extra):
        """Long prompts are encoded window by window (reference model_base.py windowed CTE loop, config.py
        ``windowed_context_encoding_size``): window i attends to the cache written by windows < i plus itself, so the
        activation footprint is that of a ``w``-token prefill whatever the prompt length.  Each row's result comes from the
        window that contains its last real token."""
        B, T = input_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        n_valid = attention_mask.long().sum(-1)
        final = None
        for s in range(0, T, w):
            sl = slice(s, min(s + w, T))
            m = attention_mask[:, sl]
            if not bool(m.any()):
                break
            kw = dict(extra)
            if s > 0:
                kw["has_prefix"] = True
            out = self.context_encoding_model(input_ids[:, sl], m, position_ids[:, sl], seq_ids, sampling_params, **kw)
            res = self._construct_output(out)
            if final is None:
                final = res
            else:
                here = ((n_valid - 1) >= s).to(res.tokens.device if res.tokens is not None else res.logits.device)
                for name in ("tokens", "logits", "hidden_states"):
                    a, b = getattr(final, name), getattr(res, name)
                    if a is not None and b is not None:
                        sel = here.view(-1, *([1] * (a.dim() - 1)))
                        setattr(final, name, torch.where(sel, b, a))
        self.kv_cache_populated = True
        return final

    @staticmethod
    def _is_prefill(position_ids, computed_context_lens=None) -> bool:
        if computed_context_lens is not None:
            return True
        # whole tensor, like the reference (model_base.py:3546): with left padding column 0 holds the pad position (1)
        return int(position_ids.min()) == 0

    def _construct_output(self, out: ModelOutput) -> CausalLMOutput:
        res = CausalLMOutput(logits=out.logits, tokens=out.tokens, hidden_states=out.hidden_states,
                             captured_tensors=out.captured)
        if out.extras:
            res.fused_outputs = out.extras.get("fused_outputs")
            res.medusa_tokens = out.extras.get("medusa_tokens")
        return res

    def reset(self):
        super().reset()
        self.kv_cache_populated = False
        if getattr(self, "fused_spec_model", None) is not None:
            self.fused_spec_model.reset()

    def reset_kv_cache(self):
        self.reset()

    def get_required_kwargs(self) -> List[str]:
        return []
