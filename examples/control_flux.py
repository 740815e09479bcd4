"""FLUX.1 Canny / Depth control (reference examples/control_flux.py): the control image is VAE-encoded and concatenated to the
latents channel-wise at every denoising step.
    python examples/control_flux.py --model-path /path/to/FLUX.1-Canny-dev --prompt "a robot" --control-image canny.png"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

import argparse

import numpy as np
import torch

from neuronx_distributed_inference_b200.config import NeuronConfig
from neuronx_distributed_inference_b200.models.diffusers.flux.application import NeuronFluxApplication
from neuronx_distributed_inference_b200.utils.diffusers_adapter import to_uint8_images


def load_image(path):
    from PIL import Image
    return torch.from_numpy(np.asarray(Image.open(path).convert("RGB"))).permute(2, 0, 1).float().div(255).unsqueeze(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-path", required=True)
    ap.add_argument("--prompt", required=True)
    ap.add_argument("--control-image", required=True)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--guidance", type=float, default=30.0)
    ap.add_argument("--tp-degree", type=int, default=1)
    ap.add_argument("--out", default="flux_control.png")
    a = ap.parse_args()
    from transformers import CLIPTokenizer, T5TokenizerFast
    clip_tok = CLIPTokenizer.from_pretrained(a.model_path, subfolder="tokenizer")
    t5_tok = T5TokenizerFast.from_pretrained(a.model_path, subfolder="tokenizer_2")
    app = NeuronFluxApplication(a.model_path, NeuronConfig(batch_size=1, torch_dtype="bfloat16", tp_degree=a.tp_degree), height=a.height,
                                width=a.width, task="control").load()
    clip_ids = clip_tok([a.prompt], padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
    t5_ids = t5_tok([a.prompt], padding="max_length", max_length=512, truncation=True, return_tensors="pt").input_ids
    img = app(clip_ids, t5_ids, num_inference_steps=a.steps, guidance_scale=a.guidance, generator=torch.Generator().manual_seed(0),
              control_image=load_image(a.control_image))
    from PIL import Image
    Image.fromarray(to_uint8_images(img)[0]).save(a.out)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
