"""b200infer — a Blackwell-native tensor-parallel LLM inference engine with the capabilities and
API surface of aws-neuron/neuronx-distributed-inference (see SURVEY.md / DESIGN.md)."""
__version__ = "0.1.0"
