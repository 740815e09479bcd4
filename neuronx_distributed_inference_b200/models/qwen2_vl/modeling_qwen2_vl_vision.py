"""Import path of the reference (models/qwen2_vl/modeling_qwen2_vl_vision.py): vision tower classes and the image-encoding application."""
from .modeling_qwen2_vl import NeuronQwen2VLForCausalLM, NeuronQwen2VLVisionModel, PatchMerger, Qwen2VLVisionBlock  # noqa: F401

NeuronQwen2VisionModel = NeuronQwen2VLVisionModel


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


class NeuronQwen2VLForImageEncoding(_ImageEncodingApplication):
    """``forward(pixel_values [n_patches, C*t*p*p], image_grid_thw=[n_img, 3])`` -> merged image tokens ``[sum(t*h*w/4), hidden]``."""
    _app_cls = NeuronQwen2VLForCausalLM
