"""Image-to-text applications (reference: generate_qwen2_vl.py / generate_qwen3_vl.py / generation_pixtral.py /
generation_mllama.py / generation_llama4.py).  The processor of the checkpoint produces the model inputs; the application
takes them as they are:

    Qwen2-VL / Qwen3-VL : app(input_ids, attention_mask=..., pixel_values=[n_patches, C*t*p*p], image_grid_thw=[n_img, 3])
    Pixtral (Llava)     : app(input_ids, attention_mask=..., pixel_values=[n_img, C, H, W], image_sizes=[n_img, 2])
    Llama-4             : app(input_ids, attention_mask=..., pixel_values=[n_tiles, C, H, W])
    Mllama              : app(input_ids, attention_mask=..., pixel_values=[B, media, tiles, C, H, W], aspect_ratio_ids=...,
                              aspect_ratio_mask=..., cross_attention_mask=[B, T, media, tiles])
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

import argparse

import torch

from neuronx_distributed_inference_b200.config import load_pretrained_config
from neuronx_distributed_inference_b200.utils.constants import get_model_cls
from neuronx_distributed_inference_b200.utils.hf_adapter import HuggingFaceGenerationAdapter


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-type", required=True, choices=["qwen2_vl", "qwen3_vl", "pixtral", "llama4", "mllama"])
    ap.add_argument("--model-path", required=True)
    ap.add_argument("--image", required=True)
    ap.add_argument("--prompt", default="Describe this image.")
    ap.add_argument("--tp-degree", type=int, default=1)
    a = ap.parse_args()
    from PIL import Image
    from transformers import AutoProcessor
    proc = AutoProcessor.from_pretrained(a.model_path)
    msgs = [{"role": "user", "content": [{"type": "image"}, {"type": "text", "text": a.prompt}]}]
    text = proc.apply_chat_template(msgs, add_generation_prompt=True)
    enc = proc(text=[text], images=[Image.open(a.image)], return_tensors="pt")
    cls = get_model_cls(a.model_type, "image-text-to-text")
    nc = cls.get_neuron_config_cls()(tp_degree=a.tp_degree, batch_size=1, max_context_length=4096, seq_len=4608, torch_dtype="bfloat16")
    app = cls(a.model_path, cls.get_config_cls()(nc, load_config=load_pretrained_config(a.model_path)))
    app.load(None)
    extra = {k: v for k, v in enc.items() if k not in ("input_ids", "attention_mask")}
    out = HuggingFaceGenerationAdapter(app).generate(enc["input_ids"], attention_mask=enc["attention_mask"], max_new_tokens=128, **extra)
    print(proc.batch_decode(out[:, enc["input_ids"].shape[1]:], skip_special_tokens=True)[0])


if __name__ == "__main__":
    main()
