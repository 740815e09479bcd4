"""Speech-to-text with Whisper (reference examples/generation_whisper.py).
    python examples/generation_whisper.py --model-path /path/to/whisper-large-v3 --audio sample.wav"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

import argparse

import torch

from neuronx_distributed_inference_b200.config import NeuronConfig, load_pretrained_config
from neuronx_distributed_inference_b200.models.whisper.modeling_whisper import NeuronApplicationWhisper, WhisperInferenceConfig


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-path", required=True)
    ap.add_argument("--audio", required=True)
    a = ap.parse_args()
    import soundfile as sf
    from transformers import WhisperProcessor
    proc = WhisperProcessor.from_pretrained(a.model_path)
    wav, sr = sf.read(a.audio)
    feats = proc(wav, sampling_rate=sr, return_tensors="pt").input_features
    nc = NeuronConfig(batch_size=1, seq_len=448, max_context_length=8, torch_dtype="bfloat16")
    app = NeuronApplicationWhisper(a.model_path, WhisperInferenceConfig(nc, load_config=load_pretrained_config(a.model_path)))
    app.load(None)
    prompt = torch.tensor([proc.get_decoder_prompt_ids(language="en", task="transcribe")]).T[1].view(1, -1)
    start = torch.cat([torch.tensor([[app.config.decoder_start_token_id]]), prompt], 1)
    toks = app.generate(feats, start, max_new_tokens=200)
    print(proc.batch_decode(toks, skip_special_tokens=True)[0])


if __name__ == "__main__":
    main()
