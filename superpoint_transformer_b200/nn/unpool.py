"""Index un-pooling (API of reference src/nn/unpool.py:7-13)."""
from torch import nn

from .. import ops

__all__ = ['IndexUnpool']


class IndexUnpool(nn.Module):
    """x_parent[idx]: redistributes level-(i+1) features to level-i nodes.
    Forward is a CUDA row gather; backward is a deterministic CSR segment-sum
    (the reference's autograd uses an atomic index_add)."""

    def forward(self, x, idx):
        return ops.index_unpool(x, idx)
