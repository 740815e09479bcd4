"""Import path of the reference (models/pixtral/modeling_pixtral_vision.py)."""
from .modeling_pixtral import NeuronPixtralForCausalLM, NeuronPixtralVisionModel, PixtralVisionLayer  # noqa: F401

NeuronPixtralAttentionLayer = PixtralVisionLayer


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


class NeuronPixtralForImageEncoding(_ImageEncodingApplication):
    """``forward(pixel_values [n_img, C, H, W], image_sizes=[n_img, 2])`` -> projected image tokens."""
    _app_cls = NeuronPixtralForCausalLM
