"""Prompt + image(s) -> model inputs through a Hugging Face processor's chat template (reference
models/{llama4,pixtral,qwen2_vl}/utils/input_processor.py ``prepare_generation_inputs_hf``, one copy per model family there)."""
from __future__ import annotations

import base64
from io import BytesIO
from typing import List, Optional, Sequence, Union


def _image_entry(img) -> dict:
    if isinstance(img, str):
        if img.startswith(("http://", "https://", "data:")):
            return {"type": "image", "url": img}
        with open(img, "rb") as f:
            data = base64.b64encode(f.read()).decode("utf-8")
        return {"type": "image", "url": f"data:image/jpeg;base64,{data}"}
    if hasattr(img, "save"):                                  # PIL.Image
        buf = BytesIO()
        img.convert("RGB").save(buf, format="JPEG")
        return {"type": "image", "url": "data:image/jpeg;base64," + base64.b64encode(buf.getvalue()).decode("utf-8")}
    raise TypeError(f"image_data entries must be paths, URLs or PIL images, got {type(img).__name__}")


def build_messages(text_prompt: str, image_data: Optional[Union[object, Sequence[object]]] = None, role: str = "user") -> List[dict]:
    """One chat turn: every image first, then the text (the order the vision-language chat templates expect)."""
    content = []
    if image_data is not None:
        for img in (image_data if isinstance(image_data, (list, tuple)) else [image_data]):
            content.append(_image_entry(img))
    content.append({"type": "text", "text": text_prompt})
    return [{"role": role, "content": content}]


def prepare_generation_inputs_hf(text_prompt: str, image_data, hf_processor, role: str = "user", config=None, **template_kwargs):
    """-> (input_ids, attention_mask, extra) where ``extra`` holds whatever else the processor produced (``pixel_values``,
    ``image_sizes``, ``image_grid_thw``, ``aspect_ratio_ids`` ...) ready to be passed to the application's ``forward`` /
    ``HuggingFaceGenerationAdapter.generate`` as keyword arguments."""
    messages = build_messages(text_prompt, image_data, role)
    inputs = hf_processor.apply_chat_template(messages, add_generation_prompt=True, tokenize=True, return_dict=True, return_tensors="pt",
                                              **template_kwargs)
    inputs = dict(inputs)
    ids, mask = inputs.pop("input_ids"), inputs.pop("attention_mask", None)
    return ids, mask, {k: v for k, v in inputs.items() if v is not None}
