"""Layer boundary markers (reference models/layer_boundary_marker.py:11-63: identity custom calls that tell the compiler's
modular flow where a layer starts and ends).  Without a tracing compiler the only consumer of layer boundaries is the
profiler: the markers open / close an NVTX range and are the identity on tensors."""
import torch


class ModuleMarkerStartWrapper(torch.nn.Module):
    def __init__(self, name: str = "layer"):
        super().__init__()
        self.name = name

    def forward(self, *tensors):
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_push(self.name)
        return tensors if len(tensors) != 1 else tensors[0]


class ModuleMarkerEndWrapper(torch.nn.Module):
    def forward(self, *tensors):
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()
        return tensors if len(tensors) != 1 else tensors[0]
