from .bucketing_processor import (BucketingProcessor, collect_buckets, get_buckets_by_model_type,  # noqa: F401
                                  select_smallest_bucket)
