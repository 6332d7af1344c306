"""MLP / FFN / Classifier (API and state-dict layout of reference src/nn/mlp.py).

The Linear layers are dense GEMMs and stay on the library path (cuBLAS); the
norm between them is the CUDA GraphNorm.  Module indices inside `self.mlp`
match the reference ModuleList ([Linear, norm, act] per layer, Linear bias only
when no norm follows; src/nn/mlp.py:37-57) so checkpoints are interchangeable.
"""
from torch import nn

from .linear import Linear

from .norm import BatchNorm, GraphNorm, INDEX_BASED_NORMS

__all__ = ['MLP', 'FFN', 'Classifier']


def _layers(dims, activation, last_activation, norm, last_norm, drop, device):
    assert len(dims) >= 2
    mods = []
    n = len(dims) - 1
    for i in range(n):
        is_last = i == n - 1
        mods.append(Linear(dims[i], dims[i + 1], bias=norm is None, device=device))
        if norm is not None and (last_norm or not is_last):
            mods.append(norm(dims[i + 1]).to(device))
        if activation is not None and (last_activation or not is_last):
            mods.append(activation.to(device))
    if drop is not None and drop > 0:
        mods.append(nn.Dropout(drop, inplace=True))
    return nn.ModuleList(mods)


class MLP(nn.Module):
    def __init__(self, dims, activation=nn.LeakyReLU(), last_activation=True,
                 norm=BatchNorm, last_norm=True, drop=None, device='cpu'):
        super().__init__()
        self.mlp = _layers(dims, activation, last_activation, norm, last_norm, drop, device)
        self.out_dim = dims[-1]

    def forward(self, x, batch=None):
        mods = list(self.mlp)
        i = 0
        fused_out = False   # x is the saved output of a fused norm+activation
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if isinstance(m, GraphNorm) and isinstance(nxt, nn.LeakyReLU) and x.is_cuda:
                # norm + LeakyReLU in one CUDA pass (forward and backward); the backward
                # reads the sign of the saved OUTPUT, so nothing may modify it in place
                x = m(x, batch=batch, act_slope=nxt.negative_slope)
                fused_out = True
                i += 2
                continue
            if isinstance(m, nn.Dropout) and m.inplace and fused_out:
                # the reference's trailing Dropout(inplace=True) (src/nn/mlp.py:56-57) would
                # overwrite the tensor the fused backward needs: same mask semantics, new tensor
                x = nn.functional.dropout(x, m.p, self.training, inplace=False)
            else:
                x = m(x, batch=batch) if isinstance(m, INDEX_BASED_NORMS) else m(x)
            fused_out = False
            i += 1
        return x


class FFN(MLP):
    """Two Linear layers, no norm, activation only in between
    (reference src/nn/mlp.py:97-125)."""

    def __init__(self, dim, hidden_dim=None, out_dim=None, activation=nn.LeakyReLU(),
                 drop=None, device='cpu'):
        super().__init__([dim, hidden_dim or dim, out_dim or dim], activation=activation,
                         last_activation=False, norm=None, last_norm=False, drop=drop,
                         device=device)


class Classifier(nn.Module):
    """Single Linear head (reference src/nn/mlp.py:128-142)."""

    def __init__(self, in_dim, num_classes, bias=True, device='cpu'):
        super().__init__()
        self.classifier = Linear(in_dim, num_classes, bias=bias, device=device)

    def forward(self, x):
        return self.classifier(x)
