"""Multi-LoRA serving (reference modules/lora_serving/*: config.py, lora_layer.py, lora_module.py, lora_model.py
``LoraModel`` :36-201, ``AdapterCache`` :294-422, ``LoraModelManager`` :425-682, lora_checkpoint.py; doc examples/slora.md).

* static multi-LoRA: up to ``max_loras`` adapters resident on the device, stacked ``A [slots, r, in]`` / ``B [slots, out, r]``
  per target projection; every request row picks its adapter with ``adapter_ids`` (slot 0 can be an all-zero "no adapter");
* dynamic multi-LoRA (``max_cpu_loras > 0``): a host pool of adapters + an LRU :class:`AdapterCache` that swaps adapters into
  device slots on demand (``add_adapter / remove_adapter / pin_adapter / list_adapters`` — the vLLM hooks of the reference);
* targets follow the engine's fused projections: q/k/v adapters write disjoint row blocks of the fused ``qkv_proj`` delta,
  gate/up of ``gate_up_proj``; tensor-parallel sharding of A/B reuses the base parameter's sharding metadata.
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .checkpoint import load_state_dict, shard_tensor

TARGETS = ("qkv_proj", "o_proj", "gate_up_proj", "down_proj")
_HF_TO_FUSED = {"q_proj": ("qkv_proj", 0), "k_proj": ("qkv_proj", 1), "v_proj": ("qkv_proj", 2), "o_proj": ("o_proj", 0),
                "gate_proj": ("gate_up_proj", 0), "up_proj": ("gate_up_proj", 1), "down_proj": ("down_proj", 0)}


class LoraLayer(nn.Module):
    """delta(x) = scaling[a] * B[a] (A[a] x) with a = adapter id of the row."""

    def __init__(self, slots: int, rank: int, in_features: int, out_features: int, dtype, device=None):
        super().__init__()
        self.A = nn.Parameter(torch.zeros(slots, rank, in_features, dtype=dtype, device=device), requires_grad=False)
        self.B = nn.Parameter(torch.zeros(slots, out_features, rank, dtype=dtype, device=device), requires_grad=False)
        self.register_buffer("scaling", torch.zeros(slots, dtype=torch.float32, device=device), persistent=False)

    def forward(self, x: torch.Tensor, adapter_ids: torch.Tensor) -> torch.Tensor:
        ids = adapter_ids.long()
        u = torch.einsum("bti,bri->btr", x, self.A[ids])
        d = torch.einsum("btr,bor->bto", u, self.B[ids])
        return d * self.scaling[ids].view(-1, 1, 1).to(d.dtype)


class _LayerHook:
    def __init__(self, owner, layer_idx):
        self.owner, self.i = owner, layer_idx

    def __call__(self, target: str, x: torch.Tensor, adapter_ids: Optional[torch.Tensor]):
        key = f"{self.i}.{target}"
        if adapter_ids is None or key not in self.owner.layers:
            return 0
        return self.owner.layers[key](x, adapter_ids)

    def has(self, target):
        return f"{self.i}.{target}" in self.owner.layers


class LoraModel(nn.Module):
    """Owns the LoRA layers of a decoder model and knows how to load PEFT checkpoints into device slots."""

    def __init__(self, model: nn.Module, lora_config, dtype=None, device=None):
        super().__init__()
        self.cfg = lora_config
        self.slots = lora_config.max_loras
        self.rank = lora_config.max_lora_rank
        self.layers = nn.ModuleDict()
        self._base = {}
        tm = set(lora_config.target_modules or ["q_proj", "k_proj", "v_proj", "o_proj"])
        fused = {t for hf, (t, _) in _HF_TO_FUSED.items() if hf in tm}
        for i, layer in enumerate(model.layers):
            attn, mlp = getattr(layer, "self_attn", None), getattr(layer, "mlp", None)
            for t, mod in (("qkv_proj", getattr(attn, "qkv_proj", None)), ("o_proj", getattr(attn, "o_proj", None)),
                           ("gate_up_proj", getattr(mlp, "gate_up_proj", None)), ("down_proj", getattr(mlp, "down_proj", None))):
                if t not in fused or mod is None or not hasattr(mod, "weight"):
                    continue
                w = mod.weight
                mult = {"qkv_proj": 3, "gate_up_proj": 2}.get(t, 1)   # fused targets stack the ranks of their sub-projections
                self.layers[f"{i}.{t}".replace(".", "_")] = LoraLayer(self.slots, self.rank * mult, w.shape[1], w.shape[0],
                                                                     dtype or w.dtype, device or w.device)
                self._base[f"{i}.{t}"] = mod
        # ModuleDict keys cannot contain '.', keep a lookup with the dotted names the hooks use
        self.layers_by_name = {k.replace("_", ".", 1): v for k, v in self.layers.items()}
        self.slot_names: List[Optional[str]] = [None] * self.slots

    # hooks ----------------------------------------------------------------------------------------------------------
    def for_layer(self, i: int) -> _LayerHook:
        h = _LayerHook(self, i)
        h.owner = _View(self.layers_by_name)
        return h

    # loading --------------------------------------------------------------------------------------------------------
    @staticmethod
    def read_peft(path: str):
        cfg = {}
        cj = os.path.join(path, "adapter_config.json")
        if os.path.isfile(cj):
            with open(cj) as f:
                cfg = json.load(f)
        sd = load_state_dict(path)
        return cfg, sd

    def load_adapter(self, slot: int, name: str, path: Optional[str] = None, state_dict: Optional[dict] = None,
                     alpha: Optional[float] = None, rank_hint: Optional[int] = None):
        """Copy one PEFT adapter into device slot ``slot`` (zero-padding ranks below ``max_lora_rank``)."""
        cfg = {}
        if state_dict is None:
            cfg, state_dict = self.read_peft(path)
        alpha = alpha if alpha is not None else cfg.get("lora_alpha", self.cfg.lora_alpha)
        for key, lay in self.layers_by_name.items():
            i, t = key.split(".", 1)
            base = self._base[key]
            lay.A.data[slot].zero_()
            lay.B.data[slot].zero_()
            parts = [(hf, blk) for hf, (ft, blk) in _HF_TO_FUSED.items() if ft == t]
            r_used = 0
            fullA_rows, fullB_blocks = [], []
            for hf, blk in parts:
                a = _find(state_dict, int(i), hf, "lora_A")
                b = _find(state_dict, int(i), hf, "lora_B")
                fullA_rows.append(a)
                fullB_blocks.append(b)
            r = max([a.shape[0] for a in fullA_rows if a is not None], default=0)
            if r == 0:
                continue
            assert r <= self.rank, f"adapter rank {r} > max_lora_rank {self.rank}"
            self._install(lay, base, slot, t, parts, fullA_rows, fullB_blocks, r)
            lay.scaling[slot] = float(alpha or r) / r
        self.slot_names[slot] = name

    def _install(self, lay, base, slot, t, parts, As, Bs, r):
        w = base.weight
        g = getattr(w, "tp_group", None)
        rank, size = (g.rank, g.size) if g is not None else (0, 1)
        dt, dev = lay.A.dtype, lay.A.device
        n = len(parts)
        if t in ("qkv_proj", "gate_up_proj"):
            # column targets: each sub-projection has its own A (stacked along the rank axis) and writes its row block of B
            full_out = _full_out_features(base, t)
            A = torch.zeros(n * r, w.shape[1])
            Bfull = torch.zeros(sum(full_out), n * r)
            ofs = 0
            for j, (a, b) in enumerate(zip(As, Bs)):
                if a is not None:
                    A[j * r:(j + 1) * r] = a.float()
                    Bfull[ofs:ofs + full_out[j], j * r:(j + 1) * r] = b.float()
                ofs += full_out[j]
            Bloc = shard_tensor(Bfull, w, rank, size)
            assert n * r <= lay.A.shape[1], "max_lora_rank must cover the stacked q/k/v (or gate/up) ranks"
            lay.A.data[slot, : n * r] = A.to(dt).to(dev)
            lay.B.data[slot, :, : n * r] = Bloc.to(dt).to(dev)
        else:
            a, b = As[0], Bs[0]
            Aloc = shard_tensor(a.float(), w, rank, size) if getattr(w, "partition_dim", None) == 1 or hasattr(w, "shard_fn") else a.float()
            lay.A.data[slot, :r] = Aloc.to(dt).to(dev)
            lay.B.data[slot, :, :r] = b.float().to(dt).to(dev)

    def unload_slot(self, slot: int):
        for lay in self.layers_by_name.values():
            lay.A.data[slot].zero_()
            lay.B.data[slot].zero_()
            lay.scaling[slot] = 0
        self.slot_names[slot] = None


class _View:
    def __init__(self, d):
        self.layers = d


def _full_out_features(base, t):
    if t == "qkv_proj":
        p, D = base.plan, base.head_dim
        return [p.n_q * D, p.n_kv * D, p.n_kv * D]
    half = base.output_size // 2
    return [half, half]


def _find(sd: Dict[str, torch.Tensor], layer: int, proj: str, which: str):
    for k, v in sd.items():
        if f"layers.{layer}." in k and f".{proj}." in k and which in k and k.endswith("weight"):
            return v
    return None


class AdapterCache:
    """LRU map adapter-name -> device slot with pinning (reference lora_model.py:294-422)."""

    def __init__(self, num_slots: int):
        self.num_slots = num_slots
        self.lru: "OrderedDict[str, int]" = OrderedDict()
        self.pinned = set()

    def lookup(self, name: str) -> Optional[int]:
        if name in self.lru:
            self.lru.move_to_end(name)
            return self.lru[name]
        return None

    def allocate(self, name: str) -> (int, Optional[str]):
        """-> (slot, evicted adapter name or None)."""
        if len(self.lru) < self.num_slots:
            used = set(self.lru.values())
            slot = next(s for s in range(self.num_slots) if s not in used)
            self.lru[name] = slot
            return slot, None
        for victim in self.lru:
            if victim not in self.pinned:
                slot = self.lru.pop(victim)
                self.lru[name] = slot
                return slot, victim
        raise RuntimeError("all LoRA slots are pinned")

    def remove(self, name: str):
        self.pinned.discard(name)
        return self.lru.pop(name, None)


class LoraModelManager:
    """Static + dynamic multi-LoRA front end: resolves request adapter names to device slots, swapping adapters in from the
    host pool when needed."""

    def __init__(self, lora_model: LoraModel, lora_config):
        self.lm = lora_model
        self.cfg = lora_config
        self.cpu_pool: Dict[str, dict] = {}
        self.cache = AdapterCache(lora_model.slots)
        for name, path in (lora_config.lora_ckpt_paths or {}).items():
            self.add_adapter(name, path)
            self.get_slot(name)
        for name, path in (lora_config.lora_ckpt_paths_cpu or {}).items():
            self.add_adapter(name, path)

    def add_adapter(self, name: str, path: Optional[str] = None, state_dict: Optional[dict] = None, alpha=None):
        if state_dict is None:
            cfg, state_dict = LoraModel.read_peft(path)
            alpha = alpha if alpha is not None else cfg.get("lora_alpha")
        self.cpu_pool[name] = dict(sd=state_dict, alpha=alpha)
        if self.cfg.max_cpu_loras and len(self.cpu_pool) > max(self.cfg.max_cpu_loras, self.lm.slots):
            raise RuntimeError("host LoRA pool is full")

    def remove_adapter(self, name: str):
        slot = self.cache.remove(name)
        if slot is not None:
            self.lm.unload_slot(slot)
        self.cpu_pool.pop(name, None)

    def pin_adapter(self, name: str):
        self.get_slot(name)
        self.cache.pinned.add(name)

    def list_adapters(self):
        return {"device": dict(self.cache.lru), "host": sorted(self.cpu_pool)}

    def get_slot(self, name: str) -> int:
        slot = self.cache.lookup(name)
        if slot is not None:
            return slot
        if name not in self.cpu_pool:
            raise KeyError(f"unknown LoRA adapter {name!r}")
        slot, _ = self.cache.allocate(name)
        e = self.cpu_pool[name]
        self.lm.load_adapter(slot, name, state_dict=e["sd"], alpha=e["alpha"])
        return slot

    def adapter_ids(self, names: List[str], device=None) -> torch.Tensor:
        return torch.tensor([self.get_slot(n) for n in names], dtype=torch.int32, device=device)
