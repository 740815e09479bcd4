"""reference models/qwen2_vl/utils/input_processor.py: chat messages + images -> model inputs through the checkpoint's processor."""
from __future__ import annotations


def prepare_generation_inputs_hf(text_prompt, image_data, hf_processor, role: str = "user", config=None):
    images = [] if image_data is None else (list(image_data) if isinstance(image_data, (list, tuple)) else [image_data])
    content = [{"type": "image"} for _ in images] + [{"type": "text", "text": text_prompt}]
    text = hf_processor.apply_chat_template([{"role": role, "content": content}], add_generation_prompt=True)
    enc = hf_processor(text=[text], images=images or None, return_tensors="pt")
    vision = {k: v for k, v in enc.items() if k not in ("input_ids", "attention_mask")}
    return enc["input_ids"], enc["attention_mask"], vision
