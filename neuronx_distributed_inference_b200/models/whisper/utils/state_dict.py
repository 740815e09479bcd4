"""reference models/whisper/utils/state_dict.py: Hugging Face Whisper weights -> the application's names."""
from ..modeling_whisper import NeuronApplicationWhisper


def convert_hf_state_dict_to_neuron(hf_state_dict, config):
    return NeuronApplicationWhisper.convert_hf_to_neuron_state_dict(hf_state_dict, config)
