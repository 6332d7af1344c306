"""Sampling transforms that run on the device right before the hot path (names, constructor
keywords and behaviour of reference src/transforms/sampling.py).  They draw indices and hand
them to `NAG.select` (csrc/select.cu)."""
import torch

from ..data import NAG

__all__ = ['SampleSubNodes', 'SampleSegments']


class SampleSubNodes:
    """Sample `low`-level elements by the `high`-level segment they belong to: at least `n_min`
    and at most `n_max` per segment, without replacement, then `nag.select(low, idx)`
    (reference src/transforms/sampling.py:656-715).  `low == high`: identity.  The per-segment
    draw is csrc/sample.cu (`seed`: optional fixed seed, otherwise torch's global generator)."""

    def __init__(self, high=1, low=0, n_max=32, n_min=16, mask=None, seed=None):
        assert isinstance(high, int)
        assert isinstance(low, int)
        assert isinstance(n_max, int)
        assert isinstance(n_min, int)
        self.high, self.low, self.n_max, self.n_min, self.mask = high, low, n_max, n_min, mask
        self.seed = seed

    def __call__(self, nag):
        assert isinstance(nag, NAG)
        if self.low == self.high:
            return nag
        idx = nag.get_sampling(high=self.high, low=self.low, n_max=self.n_max,
                               n_min=self.n_min, mask=self.mask, return_pointers=False,
                               seed=self.seed)
        return nag.select(self.low, idx)


class SampleSegments:
    """Drop a `ratio` of the nodes of every level >= 1, top level first, keeping all indices
    consistent through `NAG.select` (reference src/transforms/sampling.py:718-807).  `by_size`
    favours keeping large segments, `by_class` segments holding rare classes; the draw itself is
    `torch.multinomial` without replacement on the nodes' device, as in the reference."""

    def __init__(self, ratio=0.2, by_size=False, by_class=False):
        assert isinstance(ratio, list) and all(0 <= r < 1 for r in ratio) \
               or (0 <= ratio < 1)
        self.ratio, self.by_size, self.by_class = ratio, by_size, by_class

    def weights(self, nag, i_level):
        """Sampling weight of every node of `i_level` (sampling.py:771-798)."""
        num_nodes = nag[i_level].num_nodes
        weights = torch.ones(num_nodes, device=nag.device)
        if self.by_size:
            node_size = nag.get_sub_size(i_level, low=0)
            size_weights = node_size ** 0.333
            size_weights /= size_weights.sum()
            weights += size_weights
        if self.by_class and nag[i_level].y is not None:
            counts = nag[i_level].y.sum(dim=0).sqrt()
            scores = 1 / (counts + 1)
            scores /= scores.sum()
            mask = nag[i_level].y.gt(0)
            class_weights = (mask * scores.view(1, -1)).max(dim=1).values
            class_weights /= class_weights.sum()
            weights += class_weights.squeeze()
        weights /= weights.sum()
        return weights

    def __call__(self, nag):
        assert isinstance(nag, NAG)
        if not isinstance(self.ratio, list):
            ratio = [self.ratio] * (nag.end_i_level - max(0, nag.start_i_level - 1))
        else:
            ratio = self.ratio
        for i_level in range(nag.end_i_level, max(0, nag.start_i_level - 1), -1):
            if ratio[i_level - 1] <= 0:
                continue
            num_nodes = nag[i_level].num_nodes
            num_keep = num_nodes - int(num_nodes * ratio[i_level - 1])
            idx = torch.multinomial(self.weights(nag, i_level), num_keep, replacement=False)
            nag = nag.select(i_level, idx)
        return nag
