from ..sampling import *  # noqa: F401,F403
from ..sampling import (DataParallelSampler, Sampler, create_sampler, mask_padded_logits, prepare_sampling_params,  # noqa: F401
                        infer_sampling_params, rand_like, validate_sampling_params)
