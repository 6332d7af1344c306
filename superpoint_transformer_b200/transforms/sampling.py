"""Sampling transforms that run on the device right before the hot path (names, constructor
keywords and behaviour of reference src/transforms/sampling.py).  They draw indices and hand
them to `NAG.select` (csrc/select.cu)."""
import torch

from ..data import NAG, NAGBatch
from ..data.nag import _fill_levels

__all__ = ['SampleSubNodes', 'SampleSegments', 'SampleEdges', 'RestrictSize', 'NAGRestrictSize',
           'SampleKHopSubgraphs', 'SampleRadiusSubgraphs']


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
        return _node_weights(nag, i_level, self.by_size, self.by_class)

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


def _node_weights(nag, i_level, by_size, by_class):
    """Sampling weight of every node of `i_level`: uniform, plus a size term (level-0 size to
    the power 0.333) and a rare-class term (reference sampling.py:771-798 and :893-918, the
    same expression in both places)."""
    num_nodes = nag[i_level].num_nodes
    weights = torch.ones(num_nodes, device=nag.device)
    if by_size:
        node_size = nag.get_sub_size(i_level, low=0)
        size_weights = node_size ** 0.333
        size_weights /= size_weights.sum()
        weights += size_weights
    if by_class and nag[i_level].y is not None:
        counts = nag[i_level].y.sum(dim=0).sqrt()
        scores = 1 / (counts + 1)
        scores /= scores.sum()
        mask = nag[i_level].y.gt(0)
        class_weights = (mask * scores.view(1, -1)).max(dim=1).values
        class_weights /= class_weights.sum()
        weights += class_weights.squeeze()
    weights /= weights.sum()
    return weights


def _take_edges(data, idx):
    """Keep the edges `idx` of `data` in place: edge_index, edge_attr and every `edge_*` key
    (reference sampling.py:1303-1307), one gather launch for all of them."""
    from .. import ops
    keys = (['edge_attr'] if data.edge_attr is not None and data.edge_attr.shape[0] > 0
            else []) + data.edge_keys
    outs = ops.take_rows_multi(
        [data.edge_index[0], data.edge_index[1]] + [data[k] for k in keys], idx)
    data.edge_index = torch.stack(outs[:2])
    for k, v in zip(keys, outs[2:]):
        data[k] = v
    return data


class SampleEdges:
    """Sample the edges of the chosen levels by source node: at least `n_min` and at most
    `n_max` edges per node, within what the node has, without replacement (reference
    src/transforms/sampling.py:1234-1312; `sparse_sample` on `edge_index[0]` = csrc/sample.cu).
    Modifies the NAG in place and returns it, like the reference.

    `level`: int, 'all', 'i+' or 'i-'.  The transform is applied at exactly those levels (the
    reference's per-level closures all capture the values of the LAST level, sampling.py:1283,
    so it only samples when the last level is selected; per-level lists of n_min / n_max fail
    in the reference and are refused here)."""

    def __init__(self, level='1+', n_min=16, n_max=32, seed=None):
        assert isinstance(level, (int, str))
        if not isinstance(n_min, int) or not isinstance(n_max, int):
            raise NotImplementedError("per-level lists of n_min / n_max (they raise in the "
                                      "reference as well)")
        self.level, self.n_min, self.n_max, self.seed = level, n_min, n_max, seed

    def __call__(self, nag):
        assert isinstance(nag, NAG)
        flags = _fill_levels(self.level, False, True, nag.absolute_num_levels,
                             nag.start_i_level)
        for i_level in nag.level_range:
            if flags[i_level]:
                self._process_single_level(nag[i_level], self.n_min, self.n_max, self.seed)
        return nag

    @staticmethod
    def _process_single_level(data, n_min, n_max, seed=None):
        if n_min < 0 or n_max < 0 or not data.has_edges:
            return data
        from .. import ops
        idx = ops.sparse_sample(data.edge_index[0], n_max=n_max, n_min=n_min,
                                return_pointers=False, num_segments=data.num_nodes, seed=seed)
        return _take_edges(data, idx)


class RestrictSize:
    """At most `num_nodes` nodes and `num_edges` edges, drawn uniformly (torch.multinomial, as
    the reference): `Data.select` then an edge gather (reference sampling.py:1315-1348, which
    keeps the whole tuple `Data.select` returns and fails on the next line; the Data is used
    here)."""

    def __init__(self, num_nodes=0, num_edges=0):
        self.num_nodes, self.num_edges = num_nodes, num_edges

    def __call__(self, data):
        if data.num_nodes > self.num_nodes and self.num_nodes > 0:
            weights = torch.ones(data.num_nodes, device=data.device)
            idx = torch.multinomial(weights, self.num_nodes, replacement=False)
            data = data.select(idx)[0]
        if data.num_edges > self.num_edges and self.num_edges > 0:
            weights = torch.ones(data.num_edges, device=data.device)
            idx = torch.multinomial(weights, self.num_edges, replacement=False)
            _take_edges(data, idx)
        return data


class NAGRestrictSize:
    """Per level: at most `num_nodes` nodes (through `NAG.select`, all levels stay consistent)
    and `num_edges` edges (reference src/transforms/sampling.py:1351-1423).  `level`: int,
    'all', 'i+' or 'i-'; values <= 0 disable the restriction."""

    def __init__(self, level='1+', num_nodes=0, num_edges=0):
        assert isinstance(level, (int, str))
        assert isinstance(num_nodes, int) and isinstance(num_edges, int), \
            "per-level lists raise in the reference too (a list is compared with an int)"
        self.level, self.num_nodes, self.num_edges = level, num_nodes, num_edges

    def __call__(self, nag):
        assert isinstance(nag, NAG)
        if isinstance(self.level, int):
            return self._restrict_level(nag, self.level, self.num_nodes, self.num_edges)
        n = nag.absolute_num_levels
        level_num_nodes = _fill_levels(self.level, -1, self.num_nodes, n, nag.start_i_level)
        level_num_edges = _fill_levels(self.level, -1, self.num_edges, n, nag.start_i_level)
        for i_level in nag.level_range:
            nag = self._restrict_level(nag, i_level, level_num_nodes[i_level],
                                       level_num_edges[i_level])
        return nag

    @staticmethod
    def _restrict_level(nag, i_level, num_nodes, num_edges):
        if nag[i_level].num_nodes > num_nodes and num_nodes > 0:
            weights = torch.ones(nag[i_level].num_nodes, device=nag.device)
            idx = torch.multinomial(weights, num_nodes, replacement=False)
            nag = nag.select(i_level, idx)
        if nag[i_level].num_edges > num_edges and num_edges > 0:
            weights = torch.ones(nag[i_level].num_edges, device=nag.device)
            idx = torch.multinomial(weights, num_edges, replacement=False)
            _take_edges(nag[i_level], idx)
        return nag


class _BaseSampleSubgraphs:
    """Pick `k` seed nodes of `i_level` (uniformly, or favouring large segments / rare classes;
    with `use_batch` spread over the items of a batch), grow a node set around them
    (`_sample_subgraphs_from_seeds`) and `NAG.select` it — one NAG holding all the sets, or
    with `disjoint` a NAGBatch of one NAG per seed (reference sampling.py:810-1000)."""

    def __init__(self, i_level=1, k=1, by_size=False, by_class=False, use_batch=True,
                 disjoint=True):
        self.i_level, self.k, self.by_size, self.by_class = i_level, k, by_size, by_class
        self.use_batch, self.disjoint = use_batch, disjoint

    def seeds(self, nag, i_level):
        """The seed draw (sampling.py:884-953), torch.multinomial like the reference."""
        k = self.k if self.k < nag[i_level].num_nodes else 1
        weights = _node_weights(nag, i_level, self.by_size, self.by_class)
        batch = nag[i_level].batch
        if batch is None or not self.use_batch:
            return torch.multinomial(weights, k, replacement=False)
        idx_list = []
        batch_indices = batch.unique()
        num_batch = batch_indices.numel()
        batch_indices = batch_indices[torch.randperm(num_batch)]
        num_sampled = 0
        k_batch = max(k // num_batch, 1)
        for i_step, i_batch in enumerate(batch_indices):
            if i_step >= num_batch - 1:
                k_batch = k - num_sampled
            mask = torch.where(i_batch == batch)[0]
            idx_ = torch.multinomial(weights[mask], k_batch, replacement=False)
            idx_list.append(mask[idx_])
            num_sampled += k_batch
            if num_sampled >= k:
                break
        return torch.cat(idx_list)

    def __call__(self, nag):
        assert isinstance(nag, NAG)
        if self.i_level is None or self.k <= 0:
            return nag
        if self.i_level == -1:
            i_level = nag.end_i_level
        elif nag.start_i_level <= self.i_level < nag.absolute_num_levels:
            i_level = self.i_level
        else:
            raise ValueError(
                f"Invalid i_level: {self.i_level}. Must be in range [{nag.start_i_level}, "
                f"{nag.absolute_num_levels - 1}],\nor -1 for the highest level available.")
        idx_seed = self.seeds(nag, i_level)
        if self.disjoint:
            idx_subgraphs = [self._sample_subgraphs_from_seeds(nag, i_level, i.view(1))
                             for i in idx_seed]
            if all(idx is None for idx in idx_subgraphs):
                idx_subgraphs = None
        else:
            idx_subgraphs = self._sample_subgraphs_from_seeds(nag, i_level, idx_seed)
        if isinstance(idx_subgraphs, list):
            return NAGBatch.from_nag_list([nag.select(i_level, idx) for idx in idx_subgraphs])
        return nag.select(i_level, idx_subgraphs)

    def _sample_subgraphs_from_seeds(self, nag, i_level, idx_seed):
        raise NotImplementedError


class SampleKHopSubgraphs(_BaseSampleSubgraphs):
    """Seeds plus everything within `hops` edges of them in the graph of `i_level`, edges taken
    in both directions (reference sampling.py:1003-1091; the hop expansion is csrc/select.cu).
    `hops` None or negative: the NAG is returned as it is."""

    def __init__(self, hops=2, i_level=1, k=1, by_size=False, by_class=False, use_batch=True,
                 disjoint=False):
        super().__init__(i_level=i_level, k=k, by_size=by_size, by_class=by_class,
                         use_batch=use_batch, disjoint=disjoint)
        self.hops = hops

    def _sample_subgraphs_from_seeds(self, nag, i_level, idx_seed):
        if self.hops is None or self.hops < 0:
            return None
        assert nag[i_level].has_edges, \
            "Expected Data object to have edges for k-hop subgraph sampling"
        from .. import ops
        return ops.khop_nodes(nag[i_level].edge_index, idx_seed, self.hops,
                              nag[i_level].num_nodes)


class SampleRadiusSubgraphs(_BaseSampleSubgraphs):
    """Seeds plus every node of `i_level` within `r` of one of them (a sphere, or with
    `cylindrical` a cylinder around z; never across batch items; at most the `k_max` closest
    per seed) — reference sampling.py:1094-1231.  The neighbour search is one pass over the
    nodes (csrc/select.cu) instead of the reference's full sort of the distances.  `r` None or
    <= 0: the NAG is returned as it is."""

    def __init__(self, r=2, k_max=10000, i_level=1, k=1, by_size=False, by_class=False,
                 use_batch=True, disjoint=False, cylindrical=False):
        super().__init__(i_level=i_level, k=k, by_size=by_size, by_class=by_class,
                         use_batch=use_batch, disjoint=disjoint)
        self.r, self.k_max, self.cylindrical = r, k_max, cylindrical

    def _sample_subgraphs_from_seeds(self, nag, i_level, idx_seed):
        if self.r is None or self.r <= 0:
            return None
        from .. import ops
        return ops.radius_nodes(nag[i_level].pos, idx_seed, self.r, k_max=self.k_max,
                                batch=nag[i_level].batch, cylindrical=self.cylindrical)
