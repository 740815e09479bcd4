"""With sequence-parallel prefill the last real token of a row lives on one rank's sequence shard; the reference slices the
logits there and all-gathers one row (modules/generation/seq_parallel_logits_slice.py).  Here the residual stream is gathered
back before the last-token select (models/model_base.py), so the helper reduces to a gather of the selected rows."""
import torch


def seq_parallel_slice_last_token(hidden: torch.Tensor, position_ids: torch.Tensor, sequence_parallel_group=None, sequence_dimension: int = 1,
                                  batch_size=None, hidden_size=None, num_queries: int = 1, neuron_config=None, config=None):
    from ...parallel import mappings
    if sequence_parallel_group is not None and sequence_parallel_group.size > 1:
        hidden = mappings.all_gather(hidden.contiguous(), sequence_dimension, sequence_parallel_group)
    idx = position_ids.long().argmax(-1)
    return hidden[torch.arange(hidden.shape[0], device=hidden.device), idx].unsqueeze(1)
