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
        for name, value in (('high', high), ('low', low), ('n_max', n_max), ('n_min', n_min)):
            assert isinstance(value, int), f'{name} must be an int'
        self.high, self.low, self.n_max, self.n_min, self.mask = high, low, n_max, n_min, mask
        self.seed = seed

    def __call__(self, nag):
        assert isinstance(nag, NAG)
        if self.high == self.low:          # identity, as in the reference
            return nag
        kept = nag.get_sampling(high=self.high, low=self.low, n_max=self.n_max,
                                n_min=self.n_min, mask=self.mask, return_pointers=False,
                                seed=self.seed)
        return nag.select(self.low, kept)


class SampleSegments:
    """Drop a `ratio` of the nodes of every level >= 1, top level first, keeping all indices
    consistent through `NAG.select` (reference src/transforms/sampling.py:718-807).  `by_size`
    favours keeping large segments, `by_class` segments holding rare classes; the draw itself is
    `torch.multinomial` without replacement on the nodes' device, as in the reference."""

    def __init__(self, ratio=0.2, by_size=False, by_class=False):
        ratios = ratio if isinstance(ratio, list) else [ratio]
        assert all(0 <= r < 1 for r in ratios), "ratios must lie in [0, 1)"
        self.ratio, self.by_size, self.by_class = ratio, by_size, by_class

    def weights(self, nag, i_level):
        """Sampling weight of every node of `i_level` (sampling.py:771-798)."""
        return _node_weights(nag, i_level, self.by_size, self.by_class)

    def __call__(self, nag):
        assert isinstance(nag, NAG)
        lowest = max(0, nag.start_i_level - 1)
        per_level = self.ratio if isinstance(self.ratio, list) \
            else [self.ratio] * (nag.end_i_level - lowest)
        for level in range(nag.end_i_level, lowest, -1):      # top level first
            drop = per_level[level - 1]
            if drop <= 0:
                continue
            total = nag[level].num_nodes
            keep = total - int(total * drop)
            chosen = torch.multinomial(self.weights(nag, level), keep, replacement=False)
            nag = nag.select(level, chosen)
        return nag


def _uniform_draw(total, count, device):
    """`count` of `total` positions, uniformly, without replacement — drawn like the reference
    (torch.multinomial over unit weights), so that a torch seed fixes it the same way."""
    return torch.multinomial(torch.ones(total, device=device), count, replacement=False)


def _node_weights(nag, i_level, by_size, by_class):
    """Sampling weight of every node of `i_level`: uniform, plus a size term (level-0 size to
    the power 0.333) and a rare-class term (reference sampling.py:771-798 and :893-918, the
    same expression in both places; the tensor ops and their order are the reference's, so
    that the weights — and with them the seeded draw — agree to the bit)."""
    level = nag[i_level]
    w = torch.ones(level.num_nodes, device=nag.device)
    if by_size:
        size_term = nag.get_sub_size(i_level, low=0) ** 0.333
        size_term /= size_term.sum()
        w += size_term
    y = level['y'] if by_class else None          # (absent -> None, like the reference's Data.y)
    if y is not None:
        rarity = 1 / (y.sum(dim=0).sqrt() + 1)
        rarity /= rarity.sum()
        class_term = (y.gt(0) * rarity.view(1, -1)).max(dim=1).values
        class_term /= class_term.sum()
        w += class_term.squeeze()
    w /= w.sum()
    return w


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
        if 0 < self.num_nodes < data.num_nodes:
            data = data.select(_uniform_draw(data.num_nodes, self.num_nodes, data.device))[0]
        if 0 < self.num_edges < data.num_edges:
            _take_edges(data, _uniform_draw(data.num_edges, self.num_edges, data.device))
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
        if 0 < num_nodes < nag[i_level].num_nodes:
            nag = nag.select(i_level, _uniform_draw(nag[i_level].num_nodes, num_nodes,
                                                    nag.device))
        level = nag[i_level]
        if 0 < num_edges < level.num_edges:
            _take_edges(level, _uniform_draw(level.num_edges, num_edges, nag.device))
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
        # spread the seeds over the batch items, visited in random order: an equal share each
        # (at least one), the last item visited takes what is left
        items = batch.unique()
        items = items[torch.randperm(items.numel())]
        share, taken, picked = max(k // items.numel(), 1), 0, []
        for step, item in enumerate(items):
            take = k - taken if step >= items.numel() - 1 else share
            members = torch.where(item == batch)[0]
            picked.append(members[torch.multinomial(weights[members], take,
                                                    replacement=False)])
            taken += take
            if taken >= k:
                break
        return torch.cat(picked)

    def __call__(self, nag):
        assert isinstance(nag, NAG)
        if self.k <= 0 or self.i_level is None:
            return nag
        i_level = nag.end_i_level if self.i_level == -1 else self.i_level
        if not nag.start_i_level <= i_level < nag.absolute_num_levels:
            raise ValueError(
                f"Invalid i_level: {self.i_level}. Must be in range [{nag.start_i_level}, "
                f"{nag.absolute_num_levels - 1}],\nor -1 for the highest level available.")
        seeds = self.seeds(nag, i_level)
        if not self.disjoint:
            return nag.select(i_level, self._sample_subgraphs_from_seeds(nag, i_level, seeds))
        # one node set, hence one NAG, per seed
        node_sets = [self._sample_subgraphs_from_seeds(nag, i_level, s.view(1)) for s in seeds]
        if all(ns is None for ns in node_sets):
            return nag.select(i_level, None)
        return NAGBatch.from_nag_list([nag.select(i_level, ns) for ns in node_sets])

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
