"""Text-to-image with FLUX.1 (reference examples/generate_flux.py).
    python examples/generate_flux.py --model-path /path/to/FLUX.1-dev --prompt "A cat holding a sign that says hello world" """
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

import argparse

import torch

from neuronx_distributed_inference_b200.config import NeuronConfig
from neuronx_distributed_inference_b200.models.diffusers.flux.application import NeuronFluxApplication
from neuronx_distributed_inference_b200.utils.diffusers_adapter import to_uint8_images


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-path", required=True)
    ap.add_argument("--prompt", required=True)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=28)
    ap.add_argument("--guidance", type=float, default=3.5)
    ap.add_argument("--tp-degree", type=int, default=1)
    ap.add_argument("--out", default="flux.png")
    a = ap.parse_args()
    from transformers import CLIPTokenizer, T5TokenizerFast
    clip_tok = CLIPTokenizer.from_pretrained(a.model_path, subfolder="tokenizer")
    t5_tok = T5TokenizerFast.from_pretrained(a.model_path, subfolder="tokenizer_2")
    app = NeuronFluxApplication(a.model_path, NeuronConfig(batch_size=1, torch_dtype="bfloat16", tp_degree=a.tp_degree), height=a.height,
                                width=a.width).load()
    clip_ids = clip_tok([a.prompt], padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
    t5_ids = t5_tok([a.prompt], padding="max_length", max_length=512, truncation=True, return_tensors="pt").input_ids
    img = app(clip_ids, t5_ids, num_inference_steps=a.steps, guidance_scale=a.guidance, generator=torch.Generator().manual_seed(0))
    from PIL import Image
    Image.fromarray(to_uint8_images(img)[0]).save(a.out)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
