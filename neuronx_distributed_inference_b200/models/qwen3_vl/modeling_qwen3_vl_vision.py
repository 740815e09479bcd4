"""Import path of the reference (models/qwen3_vl/modeling_qwen3_vl_vision.py)."""
from .modeling_qwen3_vl import NeuronQwen3VLForCausalLM, NeuronQwen3VLVisionModel, Qwen3VLPatchMerger, Qwen3VLVisionBlock  # noqa: F401

NeuronQwen3VLVisionPatchMerger = Qwen3VLPatchMerger
NeuronQwen3VLVisionBlock = Qwen3VLVisionBlock


class _ImageEncodingApplication:
    """The reference ships the vision tower as its own application (``Neuron...ForImageEncoding``: compile / load / forward -> image
    embeddings).  Here the tower lives inside the image-to-text application; this wrapper exposes it under the reference's class name."""
    _app_cls = None

    def __init__(self, model_path, config=None, **kw):
        self.app = self._app_cls(model_path, config, **kw)
        self.config = self.app.config

    def compile(self, compiled_model_path, **kw):
        return self.app.compile(compiled_model_path, **kw)

    def load(self, compiled_model_path=None, **kw):
        self.app.load(compiled_model_path, **kw)
        return self

    def forward(self, pixel_values, **kw):
        return self.app.encode_images(pixel_values, **kw)

    __call__ = forward


class NeuronQwen3VLForImageEncoding(_ImageEncodingApplication):
    _app_cls = NeuronQwen3VLForCausalLM
