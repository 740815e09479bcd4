"""Attention process groups (reference modules/attention/attention_process_groups.py: TP x CP and TP x DP sub-meshes).  Groups are
created by ``parallel.state.initialize_model_parallel``; this module keeps the reference's accessor names."""
from ...parallel.state import (get_context_parallel_group as get_context_parallel_attention_cp_group,  # noqa: F401
                               get_context_parallel_tp_group as get_context_parallel_attention_tp_group,
                               get_data_parallel_attention_group as get_data_parallel_attention_dp_group,
                               get_kv_shared_group, initialize_model_parallel)


def init_context_parallel_attention_process_groups(config):
    return get_context_parallel_attention_cp_group()


def init_data_parallel_attention_process_groups(config):
    return get_data_parallel_attention_dp_group()
