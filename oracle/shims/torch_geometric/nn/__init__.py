import torch
from . import conv  # noqa


class LayerNorm(torch.nn.Module):
    """PyG 1.7 graph LayerNorm.  With batch=None (the only branch the reference reaches,
    gnn_embedder.py:194,199) statistics are taken over the WHOLE tensor:
        x = x - x.mean();  out = x / (x.std(unbiased=False) + eps)
    ("parity unpinned": restated from the published implementation, from memory)."""
    def __init__(self, in_channels, eps=1e-5, affine=True):
        super().__init__()
        self.in_channels = in_channels
        self.eps = eps
        if affine:
            self.weight = torch.nn.Parameter(torch.ones(in_channels))
            self.bias = torch.nn.Parameter(torch.zeros(in_channels))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)

    def forward(self, x, batch=None):
        if batch is not None:
            raise NotImplementedError("reference never passes batch")
        x = x - x.mean()
        out = x / (x.std(unbiased=False) + self.eps)
        if self.weight is not None:
            out = out * self.weight + self.bias
        return out


class GlobalAttention(torch.nn.Module):  # name only (learn_graph_embedding=False)
    def __init__(self, gate_nn, nn=None):
        super().__init__()
        self.gate_nn, self.nn = gate_nn, nn
