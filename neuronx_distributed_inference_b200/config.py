"""Engine + model configuration.

API parity target: ``NeuronConfig`` / ``InferenceConfig`` / ``MoENeuronConfig`` /
``OnDeviceSamplingConfig`` / ``FusedSpecNeuronConfig`` of the reference
(src/neuronx_distributed_inference/models/config.py:84-1202).  The *names* of the
knobs are kept so a user of the reference can pass the same kwargs; the
implementation is table driven (one spec table -> defaults, validation,
(de)serialisation) rather than a long sequence of ``kwargs.pop`` statements, and
the Neuron-only knobs (LNC, scratchpad, cc tiling, NKI kernel toggles ...) are
accepted and recorded but have no effect on B200 (see ``IGNORED_ON_B200``).
"""
from __future__ import annotations

import copy
import importlib
import json
import logging
import os
from typing import Any, Callable, Dict, List, Optional, Type, Union

import torch

logger = logging.getLogger("b200infer")

CONFIG_FILE = "neuron_config.json"

_DTYPES = {
    "float32": torch.float32, "fp32": torch.float32, "f32": torch.float32,
    "float16": torch.float16, "fp16": torch.float16, "f16": torch.float16,
    "bfloat16": torch.bfloat16, "bf16": torch.bfloat16,
    "float8_e4m3fn": torch.float8_e4m3fn, "f8e4m3": torch.float8_e4m3fn,
    "float8_e5m2": torch.float8_e5m2, "f8e5m2": torch.float8_e5m2,
    "int8": torch.int8,
}


def to_torch_dtype(x) -> torch.dtype:
    if isinstance(x, torch.dtype):
        return x
    if isinstance(x, str):
        key = x.replace("torch.", "")
        if key in _DTYPES:
            return _DTYPES[key]
    raise ValueError(f"unknown dtype {x!r}")


def dtype_name(d: torch.dtype) -> str:
    return str(d).replace("torch.", "")


# Knobs that only make sense for the Neuron compiler/runtime.  They are accepted for
# CLI / kwargs parity (reference inference_demo.py:299-334) and stored verbatim.
IGNORED_ON_B200 = {
    "logical_nc_config": 1, "logical_neuron_cores": None, "cc_pipeline_tiling_factor": 2,
    "seq_len_threshold_for_cc_tiling": 16384, "target": None, "enable_spill_reload_dge": False,
    "weights_to_skip_layout_optimization": [], "dma_order_config": None,
    "scratchpad_page_size": None, "enable_output_completion_notifications": False,
    "disable_numeric_cc_token": False, "switch_cc": False, "layer_boundary_markers": False,
    "enable_cte_modular_flow": False, "weight_gather_seq_len_threshold": 32768,
    "qkv_kernel_enabled": False, "qkv_nki_kernel_enabled": False,
    "qkv_cte_nki_kernel_fuse_rope": False, "qkv_kernel_nbsd_layout": False,
    "mlp_kernel_enabled": False, "mlp_tkg_nki_kernel_enabled": False,
    "fused_rmsnorm_skip_gamma": False, "mlp_kernel_fuse_residual_add": False,
    "qkv_kernel_fuse_residual_add": False, "out_proj_kernel_enabled": False,
    "attn_tkg_nki_kernel_enabled": False, "attn_tkg_builtin_kernel_enabled": False,
    "attn_block_tkg_nki_kernel_enabled": False, "attn_block_tkg_nki_kernel_cache_update": False,
    "attn_block_tkg_nki_kernel_cascaded_attention": False,
    "attn_block_tkg_nki_kernel_use_online_softmax": True,
    "attn_block_tkg_nki_kernel_disable_gpsimd_sb2sb": False,
    "attn_block_cte_nki_kernel_enabled": False,
    "moe_fused_nki_kernel_enabled": None, "router_topk_nki_kernel_enabled": None,
    "expert_mlp_nki_kernel_enabled": None, "shared_mlp_nki_kernel_enabled": None,
    "eagle_rolling_buffer_kernel_enabled": False, "disable_kv_cache_tiling": False,
    "disable_argmax_kernel": False, "is_full_model_shuffled": False,
}


class OnDeviceSamplingConfig:
    """reference: models/config.py:1064-1075."""

    def __init__(self, **kw):
        self.do_sample = kw.pop("do_sample", False)
        self.top_k = kw.pop("top_k", 1)
        self.top_p = kw.pop("top_p", 1.0)
        self.temperature = kw.pop("temperature", 1.0)
        self.dynamic = kw.pop("dynamic", False)
        self.deterministic = kw.pop("deterministic", False)
        self.global_topk = kw.pop("global_topk", 256)
        self.on_device_sampling_config = kw.pop("on_device_sampling_config", True)
        self.top_k_kernel_enabled = kw.pop("top_k_kernel_enabled", True)
        self.sampling_dp_degree = kw.pop("sampling_dp_degree", 1)
        self.seed = kw.pop("seed", 0)


class ChunkedPrefillConfig:
    """reference: models/config.py:1078-1093."""

    def __init__(self, **kw):
        self.max_num_seqs = kw.pop("max_num_seqs", 0)
        self.tkg_model_enabled = kw.pop("tkg_model_enabled", True)
        self.kernel_q_tile_size = kw.pop("kernel_q_tile_size", 128)
        self.kernel_kv_tile_size = kw.pop("kernel_kv_tile_size", 1024)


class HybridShardingConfig:
    """Different TP x EP factorisations for prefill (cte) and decode (tkg) MoE.
    reference: models/config.py:1096-1101."""

    def __init__(self, **kw):
        self.moe_cte_tp_degree = kw.pop("moe_cte_tp_degree", 1)
        self.moe_cte_ep_degree = kw.pop("moe_cte_ep_degree", 1)
        self.moe_tkg_tp_degree = kw.pop("moe_tkg_tp_degree", 1)
        self.moe_tkg_ep_degree = kw.pop("moe_tkg_ep_degree", 1)


class KVQuantizationConfig:
    """KV-cache quantisation: fp8 direct cast or static scales
    (reference kv_cache_manager.py:138-149,636-692)."""

    def __init__(self, **kw):
        self.dtype = kw.pop("dtype", "float8_e4m3fn")
        self.scale_mode = kw.pop("scale_mode", "direct_cast")  # direct_cast|per_tensor|per_head|per_channel
        self.k_scale = kw.pop("k_scale", 1.0)
        self.v_scale = kw.pop("v_scale", 1.0)


class TensorCaptureConfig:
    """reference: models/config.py:1121-1169."""

    def __init__(self, **kw):
        self.modules_to_capture = kw.pop("modules_to_capture", [])
        self.max_intermediate_tensors = kw.pop("max_intermediate_tensors", None)
        self.auto_capture_moe_tensors = kw.pop("auto_capture_moe_tensors", False)
        self.capture_inputs = kw.pop("capture_inputs", False)


class TensorReplacementConfig:
    """reference: models/config.py:1172-1202."""

    def __init__(self, **kw):
        self.ref_dir = kw.pop("ref_dir", None)
        self.neuron_dir = kw.pop("neuron_dir", None)
        self.tf_map = kw.pop("tf_map", None)
        self.module_map = kw.pop("module_map", None)


class LoraServingConfig:
    """reference: modules/lora_serving/config.py:9-224."""

    def __init__(self, **kw):
        self.max_loras = kw.pop("max_loras", 1)
        self.max_lora_rank = kw.pop("max_lora_rank", 16)
        self.max_cpu_loras = kw.pop("max_cpu_loras", 0)  # >0 -> dynamic multi-LoRA
        self.target_modules = kw.pop("target_modules", None)
        self.lora_ckpt_paths = kw.pop("lora_ckpt_paths", None) or {}
        self.lora_ckpt_paths_cpu = kw.pop("lora_ckpt_paths_cpu", None) or {}
        self.lora_dtype = kw.pop("lora_dtype", None)
        self.lora_alpha = kw.pop("lora_alpha", None)
        self.base_model_quantized = kw.pop("base_model_quantized", False)

    @property
    def dynamic_multi_lora(self):
        return self.max_cpu_loras > 0


_NESTED = {
    "on_device_sampling_config": OnDeviceSamplingConfig,
    "chunked_prefill_config": ChunkedPrefillConfig,
    "hybrid_sharding_config": HybridShardingConfig,
    "kv_quant_config": KVQuantizationConfig,
    "tensor_capture_config": TensorCaptureConfig,
    "tensor_replacement_config": TensorReplacementConfig,
    "lora_config": LoraServingConfig,
}


def _nest(key, value):
    cls = _NESTED[key]
    if value is None or isinstance(value, cls):
        return value
    if isinstance(value, dict):
        return cls(**value)
    raise TypeError(f"{key} must be a dict or {cls.__name__}")


class NeuronConfig:
    """Runtime/feature flags of the engine (kwargs compatible with the reference NeuronConfig).

    Everything the engine needs to size buffers, choose kernels and build process groups."""

    def __init__(self, **kw):
        g = kw.pop
        # ---- basic shapes (reference config.py:94-139)
        self.batch_size = g("batch_size", 1)
        self.padding_side = g("padding_side", "right")
        self.allow_input_truncation = g("allow_input_truncation", False)
        self.seq_len = g("seq_len", 128)
        self.n_active_tokens = g("n_active_tokens", self.seq_len)
        self.n_positions = g("n_positions", self.seq_len)
        self.on_cpu = g("on_cpu", False)
        self.output_logits = g("output_logits", False)
        self.torch_dtype = to_torch_dtype(g("torch_dtype", torch.bfloat16))
        self.overrides_torch_dtype = g("overrides_torch_dtype", True)
        self.cast_type = g("cast_type", "config")
        rpl = g("rpl_reduce_dtype", None)
        self.rpl_reduce_dtype = to_torch_dtype(rpl) if rpl is not None else None
        adt = g("attention_dtype", None)
        self.attention_dtype = to_torch_dtype(adt) if adt is not None else None
        self.max_context_length = g("max_context_length", self.seq_len)
        self.max_new_tokens = g("max_new_tokens", self.seq_len - self.max_context_length)
        if self.max_new_tokens == 0:
            self.max_new_tokens = None
        self.max_length = g("max_length", self.seq_len)
        self.vocab_parallel = g("vocab_parallel", False)
        self.fused_qkv = g("fused_qkv", True)  # B200: one QKV GEMM is always what we run
        self.sequence_parallel_enabled = g("sequence_parallel_enabled", False)
        self.attn_cls = g("attn_cls", "NeuronLlamaAttention")
        self.pad_token_id = g("pad_token_id", 0)

        # ---- continuous batching (config.py:161-170)
        self.ctx_batch_size = g("ctx_batch_size", self.batch_size)
        self.tkg_batch_size = g("tkg_batch_size", self.batch_size)
        self.max_batch_size = g("max_batch_size", self.batch_size)
        self.is_continuous_batching = g("is_continuous_batching", False)
        self.kv_cache_batch_size = g("kv_cache_batch_size", self.batch_size)
        self.kv_cache_padding_size = g("kv_cache_padding_size", 0)
        self.apply_seq_ids_mask = g("apply_seq_ids_mask", False)

        # ---- sampling / async
        self.on_device_sampling_config = _nest("on_device_sampling_config", g("on_device_sampling_config", None))
        self.async_mode = g("async_mode", False)

        # ---- bucketing (config.py:186-204)
        self.enable_bucketing = g("enable_bucketing", False)
        self.buckets = g("buckets", [self.seq_len])
        self.bucket_n_active_tokens = g("bucket_n_active_tokens", False)
        self.context_encoding_buckets = g("context_encoding_buckets", None)
        self.prefix_buckets = g("prefix_buckets", None)
        self.token_generation_buckets = g("token_generation_buckets", None)
        self.token_generation_batches = g("token_generation_batches", None)
        if self.token_generation_batches is not None:
            self.token_generation_batches = sorted(self.token_generation_batches)

        # ---- quantization (config.py:215-240,299-304,436-445,546-551)
        self.quantized = g("quantized", False)
        self.quantized_checkpoints_path = g("quantized_checkpoints_path", None)
        self.quantization_type = g("quantization_type", "per_tensor_symmetric")
        self.quantization_dtype = g("quantization_dtype", "int8")
        self.quantization_block_size = g("quantization_block_size", None)
        self.quantization_block_axis = g("quantization_block_axis", None)
        self.quantization_scale_dtype = g("quantization_scale_dtype", "f32")
        self.is_mxfp4_compute = g("is_mxfp4_compute", False)
        self.modules_to_not_convert = g("modules_to_not_convert", None)
        self.draft_model_modules_to_not_convert = g("draft_model_modules_to_not_convert", None)
        self.kv_cache_quant = g("kv_cache_quant", False)
        self.kv_quant_config = _nest("kv_quant_config", g("kv_quant_config", None))
        if self.kv_cache_quant and self.kv_quant_config is None:
            self.kv_quant_config = KVQuantizationConfig()
        self.quantized_mlp_kernel_enabled = g("quantized_mlp_kernel_enabled", False)
        self.activation_quantization_type = g("activation_quantization_type", None)
        self.rmsnorm_quantize_kernel_enabled = g("rmsnorm_quantize_kernel_enabled", False)
        self.quantize_clamp_bound = g("quantize_clamp_bound", float("inf"))

        # ---- speculation (config.py:243-271)
        self.speculation_length = g("speculation_length", 0)
        self.spec_batch_size = g("spec_batch_size", self.batch_size)
        self.enable_fused_speculation = g("enable_fused_speculation", False)
        self.enable_eagle_speculation = g("enable_eagle_speculation", False)
        self.is_eagle3 = g("is_eagle3", False)
        self.is_eagle_draft = g("is_eagle_draft", False)
        self.enable_eagle_draft_input_norm = g("enable_eagle_draft_input_norm", False)
        self.token_tree_config = g("token_tree_config", None)
        self.enable_token_tree = self.token_tree_config is not None
        self.is_medusa = g("is_medusa", False)
        self.medusa_speculation_length = g("medusa_speculation_length", 0)
        self.num_medusa_heads = g("num_medusa_heads", 0)
        self.medusa_tree = g("medusa_tree", None)
        if self.enable_eagle_speculation:
            self.enable_fused_speculation = True

        # ---- paged KV / prefix caching / long context
        self.is_block_kv_layout = g("is_block_kv_layout", False)
        self.pa_num_blocks = g("pa_num_blocks", self.batch_size)
        self.pa_block_size = g("pa_block_size", self.seq_len)
        self.is_prefix_caching = g("is_prefix_caching", False)
        self.windowed_context_encoding_size = g("windowed_context_encoding_size", None)
        self.chunked_prefill_config = _nest("chunked_prefill_config", g("chunked_prefill_config", None))
        self.is_chunked_prefill = self.chunked_prefill_config is not None
        # sliding-window layers keep only `sliding_window` KV slots per line, written modulo the window
        # (reference sizes SWA caches by the window: kv_cache_manager.py:194-236, gpt_oss_kv_cache_manager.py:30-396)
        self.rolling_sliding_window_cache = g("rolling_sliding_window_cache", False)
        self.k_cache_transposed = g("k_cache_transposed", False)
        self.flash_decoding_enabled = g("flash_decoding_enabled", False)
        self.attn_kernel_enabled = g("attn_kernel_enabled", None)
        self.qk_layernorm = g("qk_layernorm", False)
        self.pre_rope_rmsnorm = g("pre_rope_rmsnorm", False)

        # ---- debug tools
        self.tensor_replacement_config = _nest("tensor_replacement_config", g("tensor_replacement_config", None))
        self.tensor_capture_config = _nest("tensor_capture_config", g("tensor_capture_config", None))
        self.lora_config = _nest("lora_config", g("lora_config", None))

        # ---- parallelism (config.py:361-391)
        self.tp_degree = g("tp_degree", 1)
        self.cp_degree = g("cp_degree", 1)
        # context parallelism splits the query sequence in contiguous slices; strided = positions j, j + cp, ... (causal load balance)
        self.strided_context_parallel_kernel_enabled = g("strided_context_parallel_kernel_enabled", False)
        self.mlp_cp_degree = g("mlp_cp_degree", 1)
        self.attention_dp_degree = g("attention_dp_degree", 1)
        self.pp_degree = g("pp_degree", 1)
        self.ep_degree = g("ep_degree", 1)
        self.save_sharded_checkpoint = g("save_sharded_checkpoint", False)
        self.skip_sharding = g("skip_sharding", False)
        self.enable_ve_data_parallel = g("enable_ve_data_parallel", False)
        self.world_size = g("world_size", None)
        if self.world_size is None:
            self.world_size = self.tp_degree * self.pp_degree * self.ep_degree
        self.start_rank_id = g("start_rank_id", 0)
        self.local_ranks_size = g("local_ranks_size", None)
        if self.local_ranks_size is None:
            self.local_ranks_size = self.world_size
        self.ep_dispatch_cc_option = g("ep_dispatch_cc_option", "AR_AG")
        self.lm_head_pad = g("lm_head_pad", False)
        self.lm_head_pad_alignment_size = g("lm_head_pad_alignment_size", 1)
        self.padded_hidden_size = g("padded_hidden_size", None)
        self.padded_intermediate_size = g("padded_intermediate_size", None)
        self.skip_warmup = g("skip_warmup", False)
        self.skip_vision = g("skip_vision", False)

        # ---- B200-native knobs (no reference equivalent)
        self.cuda_graphs = g("cuda_graphs", True)          # capture decode steps per (batch, bucket)
        self.fused_collectives = g("fused_collectives", True)  # in-kernel P2P all-reduce instead of NCCL
        self.use_custom_kernels = g("use_custom_kernels", True)
        self.device = g("device", None)

        # ---- Neuron-only knobs: keep, ignore.
        self.ignored = {}
        for k, dflt in IGNORED_ON_B200.items():
            v = g(k, dflt)
            setattr(self, k, v)
            if v != dflt:
                self.ignored[k] = v
        if self.ignored:
            logger.warning("Neuron-only options have no effect on B200 and are ignored: %s",
                           sorted(self.ignored))
        if kw:
            raise TypeError(f"unknown NeuronConfig options: {sorted(kw)}")

        self._derive()
        self._validate()

    # derived state ------------------------------------------------------------------------
    def _derive(self):
        if self.is_medusa:
            assert self.medusa_speculation_length > 0 and self.num_medusa_heads > 0
        if self.attention_dp_degree > 1:
            # decode attention is data parallel: each DP group keeps batch/dp cache lines
            # (reference config.py:513-520)
            self.kv_cache_batch_size = self.tkg_batch_size // self.attention_dp_degree
        if self.rolling_sliding_window_cache:
            bad = [n for n, v in (("is_block_kv_layout", self.is_block_kv_layout), ("is_prefix_caching", self.is_prefix_caching),
                                  ("chunked_prefill_config", self.is_chunked_prefill), ("flash_decoding_enabled", self.flash_decoding_enabled),
                                  ("windowed_context_encoding_size", self.windowed_context_encoding_size),
                                  ("speculation_length", self.speculation_length), ("is_medusa", self.is_medusa),
                                  ("attention_dp_degree", self.attention_dp_degree > 1)) if v]
            if bad:
                raise ValueError(f"rolling_sliding_window_cache cannot be combined with {bad}: these read a positional prefix back "
                                 "from the cache or write several tokens per step")
        if self.is_prefix_caching:
            self.is_block_kv_layout = True
        if self.enable_fused_speculation and self.speculation_length == 0:
            raise ValueError("enable_fused_speculation requires speculation_length > 0")
        self.enable_long_context_mode = self.max_context_length >= 32 * 1024
        self.is_prefill_stage = None  # set on per-sub-model copies
        self.on_device_sampling = self.on_device_sampling_config is not None

    def _validate(self):
        if self.padding_side not in ("right", "left"):
            raise ValueError("padding_side must be 'right' or 'left'")
        if self.pp_degree != 1:
            raise ValueError("pp_degree != 1 is not supported (the reference exposes the knob only; "
                             "no pipeline schedule exists there either)")
        if self.max_context_length > self.seq_len:
            raise ValueError("max_context_length must be <= seq_len")
        if self.quantized:
            if self.quantization_type not in ("per_tensor_symmetric", "per_channel_symmetric",
                                              "expert_wise_per_channel_symmetric", "blockwise_symmetric"):
                raise ValueError(f"bad quantization_type {self.quantization_type}")
            if self.quantization_dtype not in ("int8", "f8e4m3", "f8e5m2", "mxfp8", "mxfp4"):
                raise ValueError(f"bad quantization_dtype {self.quantization_dtype}")
            if self.quantized_mlp_kernel_enabled and self.quantization_dtype != "f8e4m3":
                raise ValueError("quantized_mlp_kernel_enabled requires quantization_dtype=f8e4m3")
        if self.tp_degree % self.cp_degree != 0:
            raise ValueError("cp_degree must divide tp_degree")
        if self.attention_dp_degree > 1:
            # reference config.py:700-721
            if self.tp_degree % self.attention_dp_degree != 0:
                raise ValueError("attention_dp_degree must divide tp_degree")
            if self.tkg_batch_size % self.attention_dp_degree != 0:
                raise ValueError("tkg_batch_size must be divisible by attention_dp_degree")
            if not self.is_continuous_batching:
                raise ValueError("attention data parallel requires continuous batching")
        if self.token_generation_batches is not None:
            # reference config.py:645-689
            bad = [n for n, v in dict(
                speculation=self.speculation_length > 0, fused_speculation=self.enable_fused_speculation,
                medusa=self.is_medusa, token_tree=self.enable_token_tree,
                chunked_prefill=self.is_chunked_prefill, prefix_caching=self.is_prefix_caching,
                block_kv=self.is_block_kv_layout).items() if v]
            if bad:
                raise ValueError(f"batch bucketing (token_generation_batches) is incompatible with {bad}")
        if self.is_chunked_prefill and not self.is_block_kv_layout:
            raise ValueError("chunked prefill requires is_block_kv_layout")
        if self.flash_decoding_enabled and self.is_block_kv_layout:
            raise ValueError("flash decoding is not supported with block KV layout")

    # helpers --------------------------------------------------------------------------------
    def is_mlp_quantized(self):
        return self.quantized_mlp_kernel_enabled or self.activation_quantization_type is not None

    def copy(self, **overrides) -> "NeuronConfig":
        c = copy.deepcopy(self)
        for k, v in overrides.items():
            setattr(c, k, v)
        return c

    def to_dict(self):
        return to_dict(self)


class MoENeuronConfig(NeuronConfig):
    """reference: models/config.py:798-846."""

    def __init__(self, **kw):
        g = kw.pop
        self.capacity_factor = g("capacity_factor", None)  # None => dropless
        self.glu_mlp = g("glu_mlp", True)
        self.glu_type = g("glu_type", "glu")
        self.hidden_act_scaling_factor = g("hidden_act_scaling_factor", 1.0)
        self.hidden_act_bias = g("hidden_act_bias", 0.0)
        self.gate_clamp_upper_limit = g("gate_clamp_upper_limit", None)
        self.gate_clamp_lower_limit = g("gate_clamp_lower_limit", None)
        self.up_clamp_upper_limit = g("up_clamp_upper_limit", None)
        self.up_clamp_lower_limit = g("up_clamp_lower_limit", None)
        self.use_index_calc_kernel = g("use_index_calc_kernel", False)
        self.moe_mask_padded_tokens = g("moe_mask_padded_tokens", False)
        self.early_expert_affinity_modulation = g("early_expert_affinity_modulation", False)
        self.normalize_top_k_affinities = not g("disable_normalize_top_k_affinities", False)
        self.fused_shared_experts = g("fused_shared_experts", False)
        self.shared_experts_sequence_parallel_enabled = g("shared_experts_sequence_parallel_enabled", False)
        self.return_expert_index = g("return_expert_index", False)
        self.return_router_logits = g("return_router_logits", False)
        self.hybrid_sharding_config = _nest("hybrid_sharding_config", g("hybrid_sharding_config", None))
        self.moe_tp_degree = g("moe_tp_degree", None)
        self.moe_ep_degree = g("moe_ep_degree", None)
        self.transpose_shared_experts_weights = g("transpose_shared_experts_weights", False)
        self.blockwise_matmul_config = g("blockwise_matmul_config", {})
        rc = g("router_config", None)
        self.router_config_explicit = rc is not None      # a user-given router config overrides the family's default activation
        self.router_config = rc or {"dtype": "float32", "act_fn": "softmax"}
        super().__init__(**kw)
        if self.moe_tp_degree is None:
            self.moe_tp_degree = self.tp_degree // max(self.moe_ep_degree or 1, 1)
        if self.moe_ep_degree is None:
            self.moe_ep_degree = 1
        if self.moe_tp_degree * self.moe_ep_degree not in (self.tp_degree, self.world_size):
            raise ValueError("moe_tp_degree * moe_ep_degree must equal tp_degree (experts are re-sharded "
                             "over the same ranks that hold the attention TP shards)")


def to_dict(obj) -> Any:
    """Recursive JSON-able view (dtypes as strings, classes as module/name pairs)."""
    if obj is None or isinstance(obj, (bool, int, float, str)):
        return obj
    if isinstance(obj, torch.dtype):
        return dtype_name(obj)
    if isinstance(obj, type):
        return {"__module__": obj.__module__, "__name__": obj.__name__}
    if isinstance(obj, (list, tuple)):
        return [to_dict(x) for x in obj]
    if isinstance(obj, dict):
        return {str(k): to_dict(v) for k, v in obj.items()}
    if hasattr(obj, "__dict__"):
        return {k: to_dict(v) for k, v in vars(obj).items()
                if not k.startswith("_") and not callable(v) and k not in ("ignored",)}
    return str(obj)


_NEURON_DERIVED = {"enable_long_context_mode", "is_prefill_stage", "on_device_sampling",
                   "enable_token_tree", "is_chunked_prefill", "normalize_top_k_affinities"}


def _neuron_config_from_dict(cls, d: dict) -> NeuronConfig:
    d = {k: v for k, v in d.items() if k not in _NEURON_DERIVED}
    if "normalize_top_k_affinities" in d:
        d["disable_normalize_top_k_affinities"] = not d.pop("normalize_top_k_affinities")
    return cls(**d)


class FusedSpecNeuronConfig:
    """Pairs a target model class with a draft config (reference config.py:1045-1061)."""

    def __init__(self, worker_cls, draft_config: "InferenceConfig" = None, draft_model_path: str = None,
                 draft_model_cls=None):
        self.worker_cls = worker_cls
        self.draft_config = draft_config
        self.draft_model_path = draft_model_path
        self.draft_model_cls = draft_model_cls


class InferenceConfig:
    """HF model hyper-parameters + NeuronConfig.  reference: models/config.py:849-1042."""

    attribute_map: Dict[str, str] = {}

    def __init__(self, neuron_config: NeuronConfig, fused_spec_config=None, load_config: Callable = None,
                 metadata: Optional[Dict] = None, **kwargs):
        self.neuron_config = neuron_config
        self.fused_spec_config = fused_spec_config
        if load_config is not None:
            load_config(self)
        else:
            self.load_config()
        self.metadata = metadata
        for k, v in kwargs.items():
            setattr(self, k, v)
        self.add_derived_config()
        self.validate_config()

    def __setattr__(self, key, value):
        amap = type(self).attribute_map
        super().__setattr__(amap.get(key, key), value)

    def __getattr__(self, key):
        # only reached when normal lookup fails -> try alias
        amap = type(self).attribute_map
        if key in amap:
            return super().__getattribute__(amap[key])
        raise AttributeError(key)

    def add_derived_config(self):
        self.num_cores_per_group = 1

    def load_config(self):
        pass

    def get_required_attributes(self) -> List[str]:
        return []

    def validate_config(self):
        missing = [a for a in self.get_required_attributes() if not hasattr(self, a)]
        if missing:
            raise AssertionError(f"Config must define {missing}")
        nc = self.neuron_config
        chunk = getattr(self, "attention_chunk_size", None)
        if chunk is not None and chunk < nc.seq_len and nc.cp_degree > 1:
            assert chunk % nc.cp_degree == 0, "attention_chunk_size must be divisible by cp_degree"
        wce = nc.windowed_context_encoding_size
        sw = getattr(self, "sliding_window", None)
        if wce is not None and sw is not None:
            assert wce == sw, "windowed_context_encoding_size must equal sliding_window when both are set"

    def get_text_config(self):
        return getattr(self, "text_config", None) or self

    # ---- (de)serialisation ------------------------------------------------------------------
    def save(self, model_path: Union[str, os.PathLike]):
        os.makedirs(model_path, exist_ok=True)
        self.to_json_file(os.path.join(model_path, CONFIG_FILE))

    def to_json_file(self, path):
        with open(path, "w", encoding="utf-8") as f:
            f.write(self.to_json_string() + "\n")

    def to_json_string(self) -> str:
        return json.dumps(to_dict(self), indent=2, sort_keys=True)

    @classmethod
    def get_neuron_config_cls(cls) -> Type[NeuronConfig]:
        return NeuronConfig

    @classmethod
    def load(cls, model_path, **kwargs) -> "InferenceConfig":
        return cls.from_json_file(os.path.join(model_path, CONFIG_FILE), **kwargs)

    @classmethod
    def from_json_file(cls, path, **kwargs):
        with open(path, "r", encoding="utf-8") as f:
            return cls.from_json_string(f.read(), **kwargs)

    @classmethod
    def from_json_string(cls, s: str, **kwargs):
        d = json.loads(s)
        d.update(kwargs)
        if isinstance(d.get("neuron_config"), dict):
            d["neuron_config"] = _neuron_config_from_dict(cls.get_neuron_config_cls(), d["neuron_config"])
        fs = d.get("fused_spec_config")
        if isinstance(fs, dict):
            def _cls(ref):
                if not ref:
                    return None
                return getattr(importlib.import_module(ref["__module__"]), ref["__name__"])
            draft_cls = _cls(fs.get("draft_model_cls"))
            dc = fs.get("draft_config")
            if isinstance(dc, dict):
                cfg_cls = draft_cls.get_config_cls() if draft_cls is not None else cls
                if isinstance(dc.get("neuron_config"), dict):
                    dc["neuron_config"] = _neuron_config_from_dict(cfg_cls.get_neuron_config_cls(), dc["neuron_config"])
                dc.pop("fused_spec_config", None)
                dc = cfg_cls(**dc)
            d["fused_spec_config"] = FusedSpecNeuronConfig(
                worker_cls=_cls(fs.get("worker_cls")), draft_config=dc,
                draft_model_path=fs.get("draft_model_path"), draft_model_cls=draft_cls)
        d.pop("num_cores_per_group", None)
        return cls(**d)


def load_pretrained_config(model_path_or_name: Optional[str] = None, hf_config=None) -> Callable:
    """Return a ``load_config`` hook that copies HF ``config.json`` attributes onto an
    InferenceConfig.  reference: utils/hf_adapter.py:36-80."""

    def load_config(self: InferenceConfig):
        cfg = hf_config
        if cfg is None:
            from transformers import AutoConfig
            try:
                cfg = AutoConfig.from_pretrained(model_path_or_name)
            except (ValueError, KeyError) as e:
                # architectures that ship as remote code on the hub (MiniCPM, InternLM3, Orion...): read config.json as is
                path = os.path.join(str(model_path_or_name), "config.json")
                if not os.path.isfile(path):
                    raise
                logger.info("AutoConfig does not know this model type (%s); using raw config.json", e)
                with open(path) as f:
                    cfg = json.load(f)
        d = cfg.to_dict() if hasattr(cfg, "to_dict") else dict(cfg)
        td = d.get("torch_dtype", d.get("dtype"))
        if td is not None and not self.neuron_config.overrides_torch_dtype:
            self.neuron_config.torch_dtype = to_torch_dtype(td)
        d.pop("torch_dtype", None)
        d.pop("dtype", None)
        # nested configs (multimodal) become attribute namespaces
        for k, v in list(d.items()):
            if isinstance(v, dict) and k in ("text_config", "vision_config", "audio_config"):
                ns = InferenceConfig.__new__(InferenceConfig)
                object.__setattr__(ns, "neuron_config", self.neuron_config)
                for kk, vv in v.items():
                    object.__setattr__(ns, kk, vv)
                d[k] = ns
        for k, v in d.items():
            setattr(self, k, v)
        if hasattr(cfg, "attribute_map") and not type(self).attribute_map:
            for alias, real in cfg.attribute_map.items():
                if real in d and not hasattr(self, alias):
                    object.__setattr__(self, alias, d[real])
        if not hasattr(self, "pad_token_id") or self.pad_token_id is None:
            object.__setattr__(self, "pad_token_id", self.neuron_config.pad_token_id)

    return load_config
