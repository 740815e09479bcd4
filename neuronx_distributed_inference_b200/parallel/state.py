"""Process-group registry: one process per GPU, ``torch.distributed`` (NCCL on GPU, gloo on CPU).

Replaces ``neuronx_distributed.parallel_layers.parallel_state`` as used by the reference
(models/application_base.py:61-65,591-596; modules/attention/attention_process_groups.py:11-166;
modules/moe_v2.py:130-161; utils/distributed.py:32-38).  On an NVSwitch box every peer is
equidistant so the reference's 8x8 torus-aware mesh tables disappear: all sub-groups are
plain contiguous / strided rank lists.

Layout of the world (size W = tp):   rank r
  TP group        : all W ranks (attention heads / MLP columns sharded)
  CP x TP' groups : TP split into cp groups of tp/cp contiguous ranks (prefill attention)
  DP x TP' groups : same factorisation for decode attention data parallel
  EP x MoE-TP     : experts sharded over ep groups (strided), each expert TP-sharded inside
  draft group     : first ``draft_tp`` ranks (speculative draft with a smaller TP degree)
A world of size 1 needs no ``torch.distributed`` initialisation at all.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


@dataclass
class Group:
    """A communicator: list of global ranks + the torch ProcessGroup (None when size == 1)."""
    ranks: List[int]
    pg: Optional[object] = None
    rank: int = 0  # my index inside ``ranks``
    symm: Optional[object] = None  # lazily attached SymmetricWorkspace (parallel/symm.py)
    heap: Optional[object] = None  # lazily attached SymmetricHeap + NVLS collectives (parallel/symm_heap.py)

    @property
    def size(self) -> int:
        return len(self.ranks)


@dataclass
class _State:
    initialized: bool = False
    world_size: int = 1
    rank: int = 0
    tp: Group = field(default_factory=lambda: Group([0]))
    ep: Group = field(default_factory=lambda: Group([0]))
    moe_tp: Group = field(default_factory=lambda: Group([0]))
    cp: Group = field(default_factory=lambda: Group([0]))        # ranks that share my TP' index
    cp_tp: Group = field(default_factory=lambda: Group([0]))     # TP' group inside my CP group
    dp: Group = field(default_factory=lambda: Group([0]))
    dp_tp: Group = field(default_factory=lambda: Group([0]))
    kv_shared: Group = field(default_factory=lambda: Group([0]))  # flash-decoding KV group
    cp_block: Group = field(default_factory=lambda: Group([0]))   # cp_degree CONTIGUOUS tp ranks (general context parallelism)
    dp_block: Group = field(default_factory=lambda: Group([0]))   # attention_dp_degree contiguous tp ranks (general attention DP)
    replica: Group = field(default_factory=lambda: Group([0]))    # same shard in the other replicas of the TP group
    draft: Optional[Group] = None
    world: Group = field(default_factory=lambda: Group([0]))
    extra: Dict[str, Group] = field(default_factory=dict)


_S = _State()


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def init_distributed(backend: Optional[str] = None, device: Optional[torch.device] = None):
    """Join the job described by RANK/WORLD_SIZE/MASTER_* (torchrun) if there is one."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or dist.is_initialized():
        return _world()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29512")
    kw = {}
    if backend == "nccl":
        lr = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
        torch.cuda.set_device(lr)
        kw["device_id"] = torch.device("cuda", lr)
    dist.init_process_group(backend=backend, **kw)
    return _world()


def _make(groups: List[List[int]], me: int) -> Group:
    """Create every group in ``groups`` (collectively) and return the one containing me."""
    mine = None
    for ranks in groups:
        pg = dist.new_group(ranks) if len(ranks) > 1 else None
        if me in ranks:
            mine = Group(list(ranks), pg, ranks.index(me))
    assert mine is not None
    return mine


def initialize_model_parallel(tensor_model_parallel_size: int = 1, pipeline_model_parallel_size: int = 1,
                              expert_model_parallel_size: int = 1, context_parallel_size: int = 1,
                              attention_dp_size: int = 1, moe_tp_size: Optional[int] = None,
                              kv_shared_size: int = 1, skip_collective_init: bool = False):
    """Build all groups.  Signature follows the reference call sites
    (application_base.py:591-596, utils/testing.py:33-37)."""
    global _S
    if pipeline_model_parallel_size != 1:
        raise ValueError("pipeline parallelism is not supported")
    ws, me = _world()
    tp = tensor_model_parallel_size
    if ws == 1:
        if tp != 1 and not skip_collective_init:
            raise RuntimeError(f"tp_degree={tp} needs {tp} processes (launch with torchrun); world size is 1")
        _S = _State(initialized=True)
        return _S
    if ws % tp != 0:
        raise ValueError(f"world size {ws} not divisible by tp {tp}")
    st = _State(initialized=True, world_size=ws, rank=me)
    st.world = Group(list(range(ws)), dist.group.WORLD, me)
    n_tp_groups = ws // tp
    st.tp = _make([list(range(g * tp, (g + 1) * tp)) for g in range(n_tp_groups)], me)
    # replicas of the TP group (world = dp x tp): ranks holding the SAME shard (FLUX CFG- / context-parallel, reference dp=2)
    st.replica = _make([[i + g * tp for g in range(n_tp_groups)] for i in range(tp)], me) if n_tp_groups > 1 else Group([me])

    def split(deg):
        """Within each TP group: ``deg`` blocks of tp/deg contiguous ranks -> (outer, inner)."""
        inner_sz = tp // deg
        inner, outer = [], []
        for g in range(n_tp_groups):
            base = g * tp
            for b in range(deg):
                inner.append([base + b * inner_sz + i for i in range(inner_sz)])
            for i in range(inner_sz):
                outer.append([base + b * inner_sz + i for b in range(deg)])
        return _make(outer, me), _make(inner, me)

    if context_parallel_size > 1:
        st.cp, st.cp_tp = split(context_parallel_size)
    else:
        st.cp, st.cp_tp = Group([me]), st.tp
    if attention_dp_size > 1:
        st.dp, st.dp_tp = split(attention_dp_size)
    else:
        st.dp, st.dp_tp = Group([me]), st.tp
    ep = expert_model_parallel_size
    if ep > 1:
        mtp = moe_tp_size or tp // ep
        assert ep * mtp == tp, "ep * moe_tp must equal tp"
        # experts: EP groups are strided (ranks with the same moe-tp index), MoE-TP groups contiguous
        st.ep, st.moe_tp = split(ep)
    else:
        st.ep, st.moe_tp = Group([me]), st.tp
    if kv_shared_size > 1:
        _, st.kv_shared = split(tp // kv_shared_size)
    else:
        st.kv_shared = Group([me])
    # blocks of adjacent TP ranks for CP / attention DP when the degree is not the KV replication factor (attention_base.py)
    st.cp_block = split(tp // context_parallel_size)[1] if context_parallel_size > 1 else Group([me])
    st.dp_block = split(tp // attention_dp_size)[1] if attention_dp_size > 1 else Group([me])
    _S = st
    return st


def initialize_speculative_draft_group(draft_tp: int):
    """Draft model may run at a smaller TP (reference application_base.py:61-65)."""
    ws, me = _world()
    tp = _S.tp.size
    if draft_tp >= tp or ws == 1:
        _S.draft = _S.tp
        return _S.draft
    groups = []
    for g in range(ws // tp):
        base = g * tp
        for b in range(tp // draft_tp):
            groups.append([base + b * draft_tp + i for i in range(draft_tp)])
    _S.draft = _make(groups, me)
    return _S.draft


def destroy_model_parallel():
    global _S
    _S = _State()


def model_parallel_is_initialized() -> bool:
    return _S.initialized


def state() -> _State:
    return _S


def get_tensor_model_parallel_group() -> Group:
    return _S.tp


def get_tensor_model_parallel_size() -> int:
    return _S.tp.size


def get_tensor_model_parallel_rank() -> int:
    return _S.tp.rank


def get_expert_model_parallel_group() -> Group:
    return _S.ep


def get_expert_model_parallel_size() -> int:
    return _S.ep.size


def get_expert_model_parallel_rank() -> int:
    return _S.ep.rank


def get_moe_tp_group() -> Group:
    return _S.moe_tp


def get_data_parallel_group() -> Group:
    """Ranks that hold the same tensor-parallel shard in different replicas of the TP group (size = world / tp)."""
    return _S.replica


def get_world_group() -> Group:
    return _S.world


def get_kv_shared_group() -> Group:
    return _S.kv_shared


def get_context_parallel_block_group() -> Group:
    return _S.cp_block


def get_attention_dp_block_group() -> Group:
    return _S.dp_block


def get_speculative_draft_group() -> Group:
    return _S.draft or _S.tp


def get_context_parallel_group() -> Group:
    return _S.cp


def get_context_parallel_tp_group() -> Group:
    return _S.cp_tp


def get_data_parallel_attention_group() -> Group:
    return _S.dp


def get_data_parallel_attention_tp_group() -> Group:
    return _S.dp_tp


def get_experts_for_expert_parallel_rank(num_experts: int, ep_rank: Optional[int] = None,
                                         ep_size: Optional[int] = None) -> List[int]:
    """Contiguous expert ownership (reference test_moe_ep.py:43-69)."""
    ep_size = ep_size or _S.ep.size
    ep_rank = _S.ep.rank if ep_rank is None else ep_rank
    per = num_experts // ep_size
    return list(range(ep_rank * per, (ep_rank + 1) * per))


def get_tp_group(config) -> Group:
    """Draft models use the draft group (reference utils/distributed.py:32-38)."""
    nc = getattr(config, "neuron_config", config)
    if getattr(nc, "is_draft_model", False) and _S.draft is not None:
        return _S.draft
    return _S.tp
