"""Normalisation layers of the hot path (API of reference src/nn/norm.py).

`UnitSphereNorm` and `GraphNorm` run hand-written CUDA segment kernels
(csrc/segment.cu, csrc/norm.cu).  `GraphNorm` reproduces
torch_geometric.nn.norm.GraphNorm (the norm selected by
configs/model/semantic/_attention.yaml:9-11 and spt.yaml:19-21) including its
parameter names (`weight`, `bias`, `mean_scale`) so reference checkpoints load.
"""
import torch
from torch import nn

from .. import ops

__all__ = ['BatchNorm', 'UnitSphereNorm', 'GraphNorm', 'LayerNorm', 'GroupNorm',
           'INDEX_BASED_NORMS']


class BatchNorm(nn.Module):
    """BatchNorm1d usable on [N, C] (and [B, N, C]) tensors; dense library op,
    kept as torch (reference src/nn/norm.py:20-50; not used by the default SPT
    configs, which select GraphNorm)."""

    def __init__(self, num_features, **kwargs):
        super().__init__()
        self.batch_norm = nn.BatchNorm1d(num_features, **kwargs)

    def forward(self, x):
        if x.dim() == 2:
            return self.batch_norm(x)
        if x.dim() == 3:
            return self.batch_norm(x.transpose(1, 2)).transpose(1, 2)
        raise ValueError(f"Non supported number of dimensions {x.dim()}")


class UnitSphereNorm(nn.Module):
    """Normalise node positions inside their parent segment to a unit-diameter
    sphere: per-segment bbox -> diameter, `w`-weighted centroid -> centre
    (reference src/nn/norm.py:53-138).

    forward(pos, idx, w=None, num_super=None) -> (pos_normalised, diameter[Np,1])
    """

    def __init__(self, log_diameter=False):
        super().__init__()
        self.log_diameter = log_diameter

    def forward(self, pos, idx, w=None, num_super=None):
        pos, diameter = ops.unit_sphere_norm(pos, idx, w=w, num_super=num_super)
        if self.log_diameter:
            diameter = torch.log(diameter + 1)
        return pos, diameter


class GraphNorm(nn.Module):
    """x -> weight * (x - mean_scale * mean_g) / sqrt(var_g + eps) + bias, with
    mean/var over the nodes of each graph g = batch[i] (PyG GraphNorm, called as
    `norm(x, batch=index)` from src/nn/transformer.py:258-265, src/nn/mlp.py:89-94)."""

    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.in_channels = in_channels
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))
        self.mean_scale = nn.Parameter(torch.ones(in_channels))

    def reset_parameters(self):
        nn.init.ones_(self.weight)
        nn.init.zeros_(self.bias)
        nn.init.ones_(self.mean_scale)

    def forward(self, x, batch=None, batch_size=None, act_slope=1.0):
        return ops.graph_norm(x, self.weight, self.bias, self.mean_scale, batch=batch,
                              batch_size=batch_size, eps=self.eps, act_slope=act_slope)

    def __repr__(self):
        return f'{self.__class__.__name__}({self.in_channels})'


class LayerNorm(nn.Module):
    """torch_geometric.nn.norm.LayerNorm: `mode='graph'` (its default, and the code
    default of reference src/nn/transformer.py:137) normalises over all nodes and
    channels of each graph with the segment kernels of csrc/norm.cu; `mode='node'` is
    the dense per-node torch op."""

    def __init__(self, in_channels, eps=1e-5, affine=True, mode='graph'):
        super().__init__()
        if mode not in ('graph', 'node'):
            raise ValueError(f"Unknown normalization mode: {mode}")
        self.in_channels = in_channels
        self.eps = eps
        self.mode = mode
        if affine:
            self.weight = nn.Parameter(torch.ones(in_channels))
            self.bias = nn.Parameter(torch.zeros(in_channels))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)

    def forward(self, x, batch=None, batch_size=None):
        if self.mode == 'graph':
            # PyG puts eps outside the sqrt when no `batch` is given (x / (std + eps))
            return ops.group_norm(x, self.weight, self.bias, batch=batch, batch_size=batch_size,
                                  num_groups=1, eps=self.eps, eps_outside=batch is None)
        return torch.nn.functional.layer_norm(
            x, (self.in_channels,), self.weight, self.bias, self.eps)

    def __repr__(self):
        return f'{self.__class__.__name__}({self.in_channels}, mode={self.mode})'


class GroupNorm(nn.Module):
    """Group normalisation on graphs (reference src/nn/norm.py:141-237): `mode='graph'`
    takes mean/variance over the nodes of each graph x the channels of each group."""

    def __init__(self, in_channels, num_groups=4, eps=1e-5, affine=True, mode='graph'):
        super().__init__()
        assert in_channels % num_groups == 0, \
            "`in_channels` must be a multiple of `num_groups`"
        self.in_channels = in_channels
        self.num_groups = num_groups
        self.group_channels = in_channels // num_groups
        self.eps = eps
        self.mode = mode
        if affine:
            self.weight = nn.Parameter(torch.ones(in_channels))
            self.bias = nn.Parameter(torch.zeros(in_channels))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)

    def forward(self, x, batch=None, batch_size=None):
        if self.mode == 'graph':
            return ops.group_norm(x, self.weight, self.bias, batch=batch, batch_size=batch_size,
                                  num_groups=self.num_groups, eps=self.eps)
        if self.mode == 'node' and batch is None:
            return nn.functional.group_norm(x, self.num_groups, weight=self.weight,
                                            bias=self.bias, eps=self.eps)
        raise ValueError(f"Unknown normalization mode: {self.mode}")

    def __repr__(self):
        return (f'{self.__class__.__name__}(in_channels={self.in_channels}, '
                f'num_groups={self.num_groups}, mode={self.mode})')


INDEX_BASED_NORMS = (LayerNorm, GraphNorm, GroupNorm)
