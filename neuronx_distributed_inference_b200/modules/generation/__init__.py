"""reference layout modules/generation/{sampling,seq_parallel_logits_slice}.py"""
