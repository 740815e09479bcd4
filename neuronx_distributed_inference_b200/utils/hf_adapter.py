"""Hugging Face ``generate()`` on top of the engine.

Role of reference ``HuggingFaceGenerationAdapter`` (utils/hf_adapter.py:104-940): a ``PreTrainedModel + GenerationMixin`` subclass
whose ``_sample`` (:139-257) is a right-padding aware loop over the engine, ``prepare_inputs_for_generation`` (:259-334: positions from
the mask, last token only after prefill), attention-mask growth (:369-405) and the assisted-decoding variants (:495-915).

Two entry paths, same class:

* ``generate(input_ids, attention_mask, max_new_tokens=...)`` — the lean host loop below (one engine call + one D2H per token); used by
  the benchmarks and whenever nothing HF-specific is asked for.
* anything that needs Hugging Face's machinery — ``logits_processor``, ``prefix_allowed_tokens_fn``, ``streamer``, repetition /
  no-repeat-ngram / min-length / bad-words settings of a ``GenerationConfig``, ``StoppingCriteriaList`` objects, or
  ``use_hf_generate=True`` — goes through ``GenerationMixin.generate``: HF validates and prepares the config, builds the processor and
  stopping-criteria lists and dispatches to ``_sample``, overridden here to drive the engine (on-device sampled tokens are used
  as they are when no processor needs the logits; otherwise the step's logits are processed and sampled on the host).
"""
from __future__ import annotations

import copy
from typing import List, Optional, Union

import torch

from ..config import load_pretrained_config  # noqa: F401  (re-export, reference keeps it here)
from ..modules.sampling import prepare_sampling_params


def to_pretrained_config(config):
    """InferenceConfig -> transformers.PretrainedConfig (reference hf_adapter.py:91-101)."""
    from transformers import PretrainedConfig
    d = {k: v for k, v in vars(config).items() if k not in ("neuron_config", "fused_spec_config", "metadata")
         and not k.startswith("_") and isinstance(v, (int, float, str, bool, list, dict, type(None)))}
    return PretrainedConfig(**d)


class GenerateOutput(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def _hf_bases():
    from transformers import PreTrainedModel
    from transformers.generation.utils import GenerationMixin
    return PreTrainedModel, GenerationMixin


_PreTrainedModel, _GenerationMixin = _hf_bases()

# GenerationConfig fields that only Hugging Face's logits processors implement
_HF_ONLY_CONFIG = ("repetition_penalty", "no_repeat_ngram_size", "min_length", "min_new_tokens", "bad_words_ids", "forced_eos_token_id",
                   "suppress_tokens", "begin_suppress_tokens", "sequence_bias", "encoder_repetition_penalty", "typical_p", "epsilon_cutoff",
                   "eta_cutoff", "min_p", "exponential_decay_length_penalty", "renormalize_logits", "remove_invalid_values")


class HuggingFaceGenerationAdapter(_PreTrainedModel, _GenerationMixin):
    main_input_name = "input_ids"
    _supports_cache_class = False
    _is_stateful = True           # the KV cache lives inside the engine: no past_key_values objects travel through generate()

    def __init__(self, model, input_start_offsets=None):
        hf_config = to_pretrained_config(model.config)
        super().__init__(hf_config)
        self.neuron_model = model
        self.inference_config = model.config
        self.neuron_config = model.neuron_config
        self.padding_side = self.neuron_config.padding_side
        self.on_device_sampling = self.neuron_config.on_device_sampling_config is not None
        self.input_start_offsets = input_start_offsets
        self.prev_kv_cache_populated = False
        from transformers import GenerationConfig
        self.generation_config = GenerationConfig()
        for k in ("eos_token_id", "pad_token_id", "bos_token_id"):
            v = getattr(model.config, k, None)
            if v is not None:
                setattr(self.generation_config, k, v)

    # the engine owns the device; generate() takes and returns host tensors
    @property
    def device(self):
        return torch.device("cpu")

    def can_generate(self):
        return True

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, seq_ids=None, sampling_params=None, adapter_ids=None,
                pixel_values=None, vision_embeddings=None, vision_mask=None, image_sizes=None, return_dict=True, **kwargs):
        """One engine call in Hugging Face clothing (``CausalLMOutputWithPast`` with the step's logits when the engine returns them;
        ``tokens`` carries the on-device sampled ids)."""
        from transformers.modeling_outputs import CausalLMOutputWithPast
        extra = {k: v for k, v in dict(adapter_ids=adapter_ids, pixel_values=pixel_values, vision_embeddings=vision_embeddings,
                                       vision_mask=vision_mask, image_sizes=image_sizes).items() if v is not None}
        out = self.neuron_model(input_ids, attention_mask=attention_mask, position_ids=position_ids, seq_ids=seq_ids,
                                sampling_params=sampling_params, **extra)
        res = CausalLMOutputWithPast(logits=out.logits, past_key_values=None)
        res.tokens = out.tokens
        return res

    # ------------------------------------------------------------------------------------------------
    def _resolve(self, generation_config, kwargs):
        gc = copy.deepcopy(generation_config or self.generation_config)
        for k in list(kwargs.keys()):
            if gc is not None and hasattr(gc, k):
                setattr(gc, k, kwargs.pop(k))
        return gc

    @staticmethod
    def _eos_list(eos) -> List[int]:
        if eos is None:
            return []
        if isinstance(eos, int):
            return [eos]
        return list(eos)

    def _host_sample(self, logits: torch.Tensor, gc, generated: torch.Tensor) -> torch.Tensor:
        """logits [B,V] fp32 on host -> next tokens [B] (greedy or HF-style top-k/top-p/temperature)."""
        if not getattr(gc, "do_sample", False):
            return logits.argmax(-1)
        x = logits.float()
        t = getattr(gc, "temperature", 1.0) or 1.0
        x = x / t
        k = getattr(gc, "top_k", 0) or 0
        if k > 0:
            kth = torch.topk(x, min(k, x.shape[-1]), -1).values[:, -1:]
            x = x.masked_fill(x < kth, float("-inf"))
        p = getattr(gc, "top_p", 1.0) or 1.0
        if p < 1.0:
            sv, si = torch.sort(x, descending=True, dim=-1)
            cp = torch.softmax(sv, -1).cumsum(-1)
            drop = (cp - torch.softmax(sv, -1)) >= p
            sv = sv.masked_fill(drop, float("-inf"))
            x = torch.full_like(x, float("-inf")).scatter(1, si, sv)
        return torch.multinomial(torch.softmax(x, -1), 1).squeeze(1)

    def prepare_inputs_for_generation(self, input_ids, attention_mask=None, is_prefill: Optional[bool] = None, **kwargs):
        """positions from the mask; after prefill only the last token is fed (reference :259-334).  -> (input_ids, position_ids)"""
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        if is_prefill is None:
            is_prefill = not self.prev_kv_cache_populated
        position_ids = (attention_mask.long().cumsum(-1) - 1).clamp_min(0)
        if not is_prefill:
            position_ids = position_ids.amax(-1, keepdim=True)  # == valid length - 1
            # right padding: the newest token of a row sits at its last VALID index, not in the last column
            input_ids = input_ids.gather(1, position_ids.long())
        else:
            position_ids = position_ids.masked_fill(attention_mask == 0, 1)
        return input_ids, position_ids

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                 generation_config=None, max_new_tokens: Optional[int] = None, max_length: Optional[int] = None,
                 sampling_params: Optional[torch.Tensor] = None, seq_ids: Optional[torch.Tensor] = None,
                 assistant_model=None, return_dict_in_generate: bool = False, output_logits: bool = False,
                 output_scores: bool = False, stopping_criteria=None, adapter_ids=None, **kwargs):
        model = self.neuron_model
        nc = self.neuron_config
        if self._wants_hf_machinery(generation_config, stopping_criteria, kwargs):
            kwargs.pop("use_hf_generate", None)
            hf_kw = dict(kwargs)
            for k, v in dict(max_new_tokens=max_new_tokens, max_length=max_length, sampling_params=sampling_params, seq_ids=seq_ids,
                             adapter_ids=adapter_ids, stopping_criteria=stopping_criteria).items():
                if v is not None:
                    hf_kw[k] = v
            if return_dict_in_generate:
                hf_kw.update(return_dict_in_generate=True, output_logits=output_logits or None, output_scores=output_scores or None)
            return _GenerationMixin.generate(self, input_ids, generation_config=generation_config, attention_mask=attention_mask,
                                             assistant_model=None, **{k: v for k, v in hf_kw.items() if v is not None})
        gc = self._resolve(generation_config, kwargs)
        if gc is not None:
            max_new_tokens = max_new_tokens if max_new_tokens is not None else getattr(gc, "max_new_tokens", None)
            if max_length is None and max_new_tokens is None:
                max_length = getattr(gc, "max_length", None)
        B, P = input_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        if max_new_tokens is not None:
            max_length = P + max_new_tokens
        max_length = min(max_length or nc.max_length, nc.max_length)
        eos = self._eos_list(kwargs.pop("eos_token_id", getattr(gc, "eos_token_id", None) if gc else None))
        pad_id = kwargs.pop("pad_token_id", getattr(gc, "pad_token_id", None) if gc else None)
        if pad_id is None:
            pad_id = eos[0] if eos else (getattr(self.inference_config, "pad_token_id", 0) or 0)
        if nc.is_medusa and getattr(model, "medusa_model", None) is not None:
            from ..generation.medusa import medusa_generate
            return medusa_generate(self, input_ids, attention_mask, max_length, eos, pad_id,
                                   return_dict_in_generate=return_dict_in_generate)
        if nc.speculation_length > 0 or assistant_model is not None:
            from ..generation.speculative import assisted_generate
            return assisted_generate(self, input_ids, attention_mask, max_length, eos, pad_id, assistant_model,
                                     sampling_params=sampling_params, gc=gc,
                                     return_dict_in_generate=return_dict_in_generate)
        if sampling_params is None and self.on_device_sampling:
            c = nc.on_device_sampling_config
            if c.dynamic and gc is not None and getattr(gc, "do_sample", False):
                sampling_params = prepare_sampling_params(B, getattr(gc, "top_k", 1) or 0, getattr(gc, "top_p", 1.0),
                                                          getattr(gc, "temperature", 1.0))
            else:
                sampling_params = prepare_sampling_params(B, c.top_k, c.top_p, c.temperature)
        model.reset()
        sequences = input_ids.clone()
        mask = attention_mask.clone()
        unfinished = torch.ones(B, dtype=torch.bool)
        eos_t = torch.tensor(eos, dtype=torch.long) if eos else None
        all_logits = []
        want_logits = output_logits or output_scores or not self.on_device_sampling
        cur_len = P
        is_prefill = True
        extra = {}
        if adapter_ids is not None:
            extra["adapter_ids"] = adapter_ids
        for k in ("pixel_values", "vision_embeddings", "vision_mask", "image_sizes"):
            if k in kwargs and kwargs[k] is not None:
                extra[k] = kwargs[k]
        while cur_len < max_length:
            ids, pos = self.prepare_inputs_for_generation(sequences, mask, is_prefill)
            out = model(ids, attention_mask=mask if is_prefill else None, position_ids=pos.to(torch.int32),
                        seq_ids=seq_ids, sampling_params=sampling_params,
                        **(extra if is_prefill else {k: v for k, v in extra.items() if k == "adapter_ids"}))
            is_prefill = False
            if out.tokens is not None and self.on_device_sampling:
                nxt = out.tokens.reshape(B, -1)[:, -1].to("cpu", torch.long)
            else:
                lg = out.logits[:, -1].float().cpu()
                nxt = self._host_sample(lg, gc, sequences)
            if want_logits and out.logits is not None:
                all_logits.append(out.logits[:, -1].float().cpu())
            nxt = torch.where(unfinished, nxt, torch.full_like(nxt, pad_id))
            # right padding: the new token goes right after each row's last valid token
            sequences, mask = _append_tokens(sequences, mask, nxt, pad_id, self.padding_side)
            cur_len += 1
            if eos_t is not None:
                unfinished = unfinished & ~torch.isin(nxt, eos_t)
                if not bool(unfinished.any()):
                    break
            if stopping_criteria is not None and bool(torch.as_tensor(stopping_criteria(sequences, None)).all()):
                break
        if self.padding_side == "right":
            sequences = _compact_right_padded(sequences, mask, pad_id)
        if return_dict_in_generate:
            return GenerateOutput(sequences=sequences, logits=all_logits or None, scores=all_logits or None)
        return sequences

    # ---- Hugging Face machinery ---------------------------------------------------------------------------------------------------
    def _wants_hf_machinery(self, generation_config, stopping_criteria, kwargs) -> bool:
        if kwargs.get("use_hf_generate") is not None:
            return bool(kwargs["use_hf_generate"])
        nc = self.neuron_config
        if nc.is_medusa or nc.speculation_length > 0 or kwargs.get("assistant_model") is not None:
            return False      # speculative variants have their own device-side loops (generation/speculative.py, medusa.py)
        if any(kwargs.get(k) is not None for k in ("logits_processor", "prefix_allowed_tokens_fn", "streamer", "negative_prompt_ids")):
            return True
        try:
            from transformers import StoppingCriteriaList
            if isinstance(stopping_criteria, StoppingCriteriaList):
                return True
        except Exception:  # pragma: no cover
            pass
        from transformers import GenerationConfig
        default = GenerationConfig()
        for src in (generation_config, kwargs):
            for k in _HF_ONLY_CONFIG:
                v = src.get(k) if isinstance(src, dict) else getattr(src, k, None) if src is not None else None
                if v is not None and v != getattr(default, k, None):
                    return True
        return False

    def _validate_model_kwargs(self, model_kwargs):
        pass      # engine kwargs (seq_ids, sampling_params, adapter_ids, vision inputs) are forwarded as they are

    def _prepare_cache_for_generation(self, *a, **k):
        pass      # no Cache object: the engine's KV cache is addressed by seq_ids

    def _supports_default_dynamic_cache(self, *a, **k):
        return False

    @torch.no_grad()
    def _sample(self, input_ids, logits_processor, stopping_criteria, generation_config, synced_gpus=False, streamer=None,
                **model_kwargs):
        """Called by ``GenerationMixin.generate`` (greedy / multinomial modes) after HF prepared the config, the logits-processor
        list and the stopping criteria.  Reference: hf_adapter.py:139-257."""
        from transformers.generation.utils import GenerateDecoderOnlyOutput
        model, nc = self.neuron_model, self.neuron_config
        B = input_ids.shape[0]
        mask = model_kwargs.get("attention_mask")
        mask = torch.ones_like(input_ids) if mask is None else mask.clone()
        seq_ids, sampling_params = model_kwargs.get("seq_ids"), model_kwargs.get("sampling_params")
        extra = {k: model_kwargs[k] for k in ("adapter_ids", "pixel_values", "vision_embeddings", "vision_mask", "image_sizes")
                 if model_kwargs.get(k) is not None}
        pad = generation_config.pad_token_id
        eos = self._eos_list(generation_config.eos_token_id)
        if pad is None:
            pad = eos[0] if eos else 0
        pad = int(pad if not torch.is_tensor(pad) else pad.flatten()[0])
        needs_logits = len(logits_processor) > 0 or not self.on_device_sampling or generation_config.output_scores \
            or generation_config.output_logits
        if needs_logits and not (nc.output_logits or not self.on_device_sampling):
            raise ValueError("logits processors / output_scores need the logits on the host: build the model with output_logits=True "
                             "(or without on-device sampling)")
        if sampling_params is None and self.on_device_sampling:
            c = nc.on_device_sampling_config
            sampling_params = prepare_sampling_params(B, c.top_k, c.top_p, c.temperature)
        model.reset()
        sequences, unfinished = input_ids.clone(), torch.ones(B, dtype=torch.bool)
        scores, raw = [], []
        is_prefill = True
        while sequences.shape[1] < nc.max_length:
            ids, pos = self.prepare_inputs_for_generation(sequences, mask, is_prefill)
            out = model(ids, attention_mask=mask if is_prefill else None, position_ids=pos.to(torch.int32), seq_ids=seq_ids,
                        sampling_params=sampling_params,
                        **(extra if is_prefill else {k: v for k, v in extra.items() if k == "adapter_ids"}))
            is_prefill = False
            self.prev_kv_cache_populated = True
            if needs_logits:
                lg = out.logits[:, -1].float().cpu()
                # processors see the row's real tokens (right-padded rows: up to its last valid token)
                proc = logits_processor(sequences, lg.clone())
                if generation_config.output_logits:
                    raw.append(lg)
                if generation_config.output_scores:
                    scores.append(proc)
                nxt = torch.multinomial(torch.softmax(proc, -1), 1).squeeze(1) if generation_config.do_sample else proc.argmax(-1)
            else:
                nxt = out.tokens.reshape(B, -1)[:, -1].to("cpu", torch.long)
            nxt = torch.where(unfinished, nxt, torch.full_like(nxt, pad))
            sequences, mask = _append_tokens(sequences, mask, nxt, pad, self.padding_side)
            if streamer is not None:
                streamer.put(nxt)
            done = stopping_criteria(sequences, tuple(scores) if scores else None)
            unfinished = unfinished & ~torch.as_tensor(done, dtype=torch.bool).reshape(-1).expand(B)
            if not bool(unfinished.any()):
                break
        self.prev_kv_cache_populated = False
        if streamer is not None:
            streamer.end()
        if self.padding_side == "right":
            sequences = _compact_right_padded(sequences, mask, pad)
        if generation_config.return_dict_in_generate:
            return GenerateDecoderOnlyOutput(sequences=sequences, scores=tuple(scores) or None, logits=tuple(raw) or None,
                                             past_key_values=None)
        return sequences


def _append_tokens(sequences, mask, nxt, pad_id, padding_side):
    """Grow ``sequences`` / ``mask`` by one column; right-padded rows receive the token right after their last valid one."""
    B = sequences.shape[0]
    if padding_side == "right" and bool((mask.sum(-1) < sequences.shape[1]).any()):
        sequences = torch.cat([sequences, torch.full((B, 1), pad_id, dtype=sequences.dtype)], 1)
        mask = torch.cat([mask, torch.zeros(B, 1, dtype=mask.dtype)], 1)
        idx = mask.long().sum(-1)
        sequences[torch.arange(B), idx] = nxt
        mask[torch.arange(B), idx] = 1
        return sequences, mask
    return torch.cat([sequences, nxt.view(B, 1)], 1), torch.cat([mask, torch.ones(B, 1, dtype=mask.dtype)], 1)


def _compact_right_padded(sequences, mask, pad_id):
    """Rows were kept right padded during generation ([prompt, new..., pads]); return them as such."""
    out = torch.full_like(sequences, pad_id)
    for b in range(sequences.shape[0]):
        valid = sequences[b][mask[b].bool()]
        out[b, : valid.numel()] = valid
    width = int(mask.sum(-1).max())
    return out[:, :width]
