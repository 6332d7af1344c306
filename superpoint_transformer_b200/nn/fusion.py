"""Feature fusion operators (API of reference src/nn/fusion.py:8-57)."""
import torch
from torch import nn

__all__ = ['CatFusion', 'AdditiveFusion', 'TakeFirstFusion', 'TakeSecondFusion',
           'fusion_factory']

_CAT = ('cat', 'concatenate', 'concatenation', '|')
_ADD = ('residual', 'additive', '+')
_FIRST = ('first', '1', '1st')
_SECOND = ('second', '2', '2nd')


class _Fusion(nn.Module):
    """Binary fusion where a missing operand (None) is the identity
    (reference src/nn/fusion.py:25-37)."""

    def forward(self, x1, x2):
        if x1 is None:
            return x2
        if x2 is None:
            return x1
        return self._fuse(x1, x2)


class CatFusion(_Fusion):
    def _fuse(self, x1, x2):
        return torch.cat((x1, x2), dim=1)


class AdditiveFusion(_Fusion):
    def _fuse(self, x1, x2):
        return x1 + x2


class TakeFirstFusion(_Fusion):
    def _fuse(self, x1, x2):
        return x1


class TakeSecondFusion(_Fusion):
    def _fuse(self, x1, x2):
        return x2


def fusion_factory(mode):
    """String -> fusion module (reference src/nn/fusion.py:8-22)."""
    for names, cls in ((_CAT, CatFusion), (_ADD, AdditiveFusion),
                       (_FIRST, TakeFirstFusion), (_SECOND, TakeSecondFusion)):
        if mode in names:
            return cls()
    raise NotImplementedError(f"Unknown mode='{mode}'")
