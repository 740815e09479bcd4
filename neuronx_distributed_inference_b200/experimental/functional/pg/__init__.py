from .context_parallel import (get_context_parallel_cp_group, get_context_parallel_cp_mesh, get_context_parallel_tp_group,  # noqa: F401
                               get_context_parallel_tp_mesh, get_cp_rank, initialize_context_parallel_process_groups)
from .data_parallel import get_dp_rank  # noqa: F401
