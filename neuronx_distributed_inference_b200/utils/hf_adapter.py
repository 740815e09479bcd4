"""Hugging Face-style ``generate()`` on top of the engine.

Role of reference ``HuggingFaceGenerationAdapter`` (utils/hf_adapter.py:104-940): right-padding aware sampling
loop (``_sample`` :139-257), ``prepare_inputs_for_generation`` (:259-334: positions from the mask, last token
only after prefill), attention-mask growth (:369-405) and the assisted-decoding variants (:495-915).

The reference subclasses ``PreTrainedModel + GenerationMixin`` and overrides private hooks; those hooks change
between transformers releases, so this adapter owns its loop and only borrows ``GenerationConfig`` /
``StoppingCriteria`` *data* from transformers.  Behaviour kept: ``generate(input_ids, attention_mask, ...)``
returns ``[B, prompt + new]`` sequences padded with ``pad_token_id`` after EOS; on-device sampling tokens are
used when the model samples on device, otherwise logits are sampled on the host with HF logits processors.
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


class HuggingFaceGenerationAdapter:
    def __init__(self, model, input_start_offsets=None):
        self.neuron_model = model
        self.config = model.config
        self.neuron_config = model.neuron_config
        self.padding_side = self.neuron_config.padding_side
        self.on_device_sampling = self.neuron_config.on_device_sampling_config is not None
        self.input_start_offsets = input_start_offsets
        self.generation_config = None
        self.prev_kv_cache_populated = False
        try:
            from transformers import GenerationConfig
            self.generation_config = GenerationConfig()
        except Exception:  # pragma: no cover
            pass

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

    def prepare_inputs_for_generation(self, input_ids, attention_mask, is_prefill: bool):
        """positions from the mask; after prefill only the last token is fed (reference :259-334)."""
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
            pad_id = eos[0] if eos else (getattr(self.config, "pad_token_id", 0) or 0)
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
            if self.padding_side == "right" and bool((mask.sum(-1) < sequences.shape[1]).any()):
                sequences = torch.cat([sequences, torch.full((B, 1), pad_id, dtype=sequences.dtype)], 1)
                mask = torch.cat([mask, torch.zeros(B, 1, dtype=mask.dtype)], 1)
                idx = mask.long().sum(-1)
                sequences[torch.arange(B), idx] = nxt
                mask[torch.arange(B), idx] = 1
            else:
                sequences = torch.cat([sequences, nxt.view(B, 1)], 1)
                mask = torch.cat([mask, torch.ones(B, 1, dtype=mask.dtype)], 1)
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

    __call__ = generate


def _compact_right_padded(sequences, mask, pad_id):
    """Rows were kept right padded during generation ([prompt, new..., pads]); return them as such."""
    out = torch.full_like(sequences, pad_id)
    for b in range(sequences.shape[0]):
        valid = sequences[b][mask[b].bool()]
        out[b, : valid.numel()] = valid
    width = int(mask.sum(-1).max())
    return out[:, :width]
