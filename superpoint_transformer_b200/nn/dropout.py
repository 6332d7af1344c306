"""Stochastic depth (API of reference src/nn/dropout.py:7-40)."""
from torch import nn

__all__ = ['DropPath']


class DropPath(nn.Module):
    """Per-sample (per-row) path dropping on the residual branch; identity in
    eval mode or when drop_prob == 0 (reference src/nn/dropout.py:7-21)."""

    def __init__(self, drop_prob: float = 0., scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if not self.training or self.drop_prob == 0.:
            return x
        keep = 1. - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0. and self.scale_by_keep:
            mask.div_(keep)
        return x * mask

    def extra_repr(self):
        return f'drop_prob={round(self.drop_prob, 3):0.3f}'
