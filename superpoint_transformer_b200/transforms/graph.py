"""On-device NAG transforms that feed the hot path every batch (API of the
corresponding classes in reference src/transforms/graph.py):
  NodeSize                          :1475-1498  (int64 exact)
  OnTheFlyHorizontalEdgeFeatures    :1063-1277  (18-column edge_attr, both directions)
  NAGAddSelfLoops                   :1419-1452  (N zero-feature self-loops appended)
`OnTheFlyHorizontalEdgeFeatures(add_self_loops=True)` fuses the last two into the
single CUDA pass of csrc/edge_features.cu.
"""
import torch

from .. import ops

__all__ = ['NodeSize', 'OnTheFlyHorizontalEdgeFeatures', 'OnTheFlyVerticalEdgeFeatures',
           'NAGAddSelfLoops', 'ON_THE_FLY_HORIZONTAL_FEATURES', 'ON_THE_FLY_VERTICAL_FEATURES',
           'minimalistic_horizontal_edge_features', 'cluster_point_features']

# columns of each key in the 18-column output (reference f_list assembly,
# src/transforms/graph.py:1188-1266: mean_off is PREPENDED, the others appended in this order)
_H_COLUMNS = {'mean_off': (0, 1, 2), 'std_off': (3, 4, 5), 'mean_dist': (6,),
              'angle_source': (7,), 'angle_target': (8,), 'normal_angle': (9,),
              'log_length': (10,), 'log_surface': (11,), 'log_volume': (12,), 'log_size': (13,),
              'centroid_dir': (14, 15, 16), 'centroid_dist': (17,)}

# column order of the reference's f_list assembly
ON_THE_FLY_HORIZONTAL_FEATURES = [
    'mean_off', 'std_off', 'mean_dist', 'angle_source', 'angle_target', 'normal_angle',
    'log_length', 'log_surface', 'log_volume', 'log_size', 'centroid_dir', 'centroid_dist']


class NodeSize:
    """nag[i].node_size = number of level-`low` elements under each level-i node."""

    def __init__(self, low=0):
        assert isinstance(low, int) and low >= -1
        self.low = low

    def __call__(self, nag):
        start = nag.start_i_level
        low = max(self.low, start - 1)
        for i_level in range(max(self.low + 1, start), nag.absolute_num_levels):
            d = nag[i_level]
            if i_level == start and low < start and d.sub is None and d.node_size is not None:
                continue  # first loaded level of a nano NAG: sizes come with the data
            d.node_size = nag.get_sub_size(i_level, low=low)
        return nag


class OnTheFlyHorizontalEdgeFeatures:
    """Symmetrise the trimmed graph and build the 18 handcrafted edge features.

    Input per level (>= 1): trimmed `edge_index` [2, Eh] (i<j once), `edge_attr`
    [Eh, 7] (mean_off 3, std_off 3, mean_dist 1; fp16 or fp32), `pos`, `normal`,
    `log_length/log_surface/log_volume/log_size` [N, 1].  Output: `edge_index`
    [2, 2Eh (+N)] ordered [i->j | j->i (| self-loops)], `edge_attr` fp32 18 columns.
    A subset of `keys` keeps the columns of those keys (in the reference's assembly order);
    no key at all leaves `edge_attr = None` like the reference.

    `csr_order=True` (an extension; the model is invariant to the order of the edges)
    additionally groups the output edges by source node, stably, so that the attention
    blocks consume `edge_attr` in place instead of permuting it (and un-permuting its
    gradient) once per stage and step.
    """

    def __init__(self, keys=None, use_mean_normal=False, add_self_loops=False,
                 csr_order=False):
        keys = ON_THE_FLY_HORIZONTAL_FEATURES if keys is None else list(keys)
        unknown = [k for k in keys if k not in _H_COLUMNS]
        if unknown:
            raise ValueError(f"unknown horizontal edge feature keys {unknown}")
        self.keys = [k for k in ['mean_off'] + [k for k in ON_THE_FLY_HORIZONTAL_FEATURES
                                                if k != 'mean_off'] if k in keys]
        cols = [c for k in self.keys for c in _H_COLUMNS[k]]
        self.columns = None if cols == list(range(18)) else cols
        self.normal_key = 'mean_normal' if use_mean_normal else 'normal'
        self.add_self_loops = add_self_loops
        self.csr_order = csr_order

    def __call__(self, nag):
        for i_level in nag.level_range:
            if i_level == 0:
                continue
            d = nag[i_level]
            if d.edge_index is None:
                continue
            ei, ea = ops.horizontal_edge_features(
                d.edge_index, d.edge_attr, d.pos, d[self.normal_key], d['log_length'],
                d['log_surface'], d['log_volume'], d['log_size'], d.num_nodes,
                add_self_loops=self.add_self_loops)
            if self.columns is not None:   # key subset: the columns of the requested keys
                ea = ea[:, torch.tensor(self.columns, device=ea.device)].contiguous() \
                    if self.columns else None
            if self.csr_order:
                seg = ops.group_index(ei[0], d.num_nodes)
                ei = ei.index_select(1, seg.perm.long())
                if ea is not None:
                    ea = ops._gather_rows(ea, seg.perm)
                ops.mark_csr_ordered(ei)
            d.edge_index, d.edge_attr = ei, ea
        return nag


ON_THE_FLY_VERTICAL_FEATURES = [
    'centroid_dir', 'centroid_dist', 'normal_angle', 'log_length', 'log_surface', 'log_volume',
    'log_size']


class OnTheFlyVerticalEdgeFeatures:
    """child -> parent edge features `v_edge_attr` [Nc, 9] for every loaded level that has
    a parent (reference src/transforms/graph.py:1280-1416; consumed by the attentive pools).
    Only the full default key set is built by the kernel."""

    def __init__(self, keys=None, use_mean_normal=False):
        keys = ON_THE_FLY_VERTICAL_FEATURES if keys is None else list(keys)
        self.enabled = len(keys) > 0
        if self.enabled and sorted(keys) != sorted(ON_THE_FLY_VERTICAL_FEATURES):
            raise NotImplementedError(
                "the CUDA vertical edge-feature kernel builds the full 9-column default set")
        self.normal_key = 'mean_normal' if use_mean_normal else 'normal'

    def __call__(self, nag):
        if not self.enabled:
            return nag
        for i_level in range(nag.start_i_level + 1, nag.absolute_num_levels):
            child, parent = nag[i_level - 1], nag[i_level]
            child.v_edge_attr = ops.vertical_edge_features(child, parent, self.normal_key)
        return nag


class NAGAddSelfLoops:
    """Append one self-loop per node with zero features (PyG add_self_loops,
    fill_value=0).  Concatenation only — pure data movement, done with torch.cat;
    prefer OnTheFlyHorizontalEdgeFeatures(add_self_loops=True) which fuses it."""

    def __call__(self, nag):
        for i_level in range(max(nag.start_i_level, 1), nag.absolute_num_levels):
            d = nag[i_level]
            if not d.has_edges:
                continue
            n, dev = d.num_nodes, d.edge_index.device
            loops = torch.arange(n, device=dev, dtype=d.edge_index.dtype)
            d.edge_index = torch.cat((d.edge_index, torch.stack((loops, loops))), dim=1)
            if d.edge_attr is not None:
                pad = torch.zeros((n, d.edge_attr.shape[1]), dtype=d.edge_attr.dtype,
                                  device=dev)
                d.edge_attr = torch.cat((d.edge_attr, pad), dim=0)
        return nag


def minimalistic_horizontal_edge_features(data, points, se_point_index, se_id, keys=None):
    """Superedge features from the level-0 sub-edges of every superedge (reference
    `_minimalistic_horizontal_edge_features`, src/transforms/graph.py:950-1060): sets
    `data.edge_attr = [mean_off | std_off | mean_dist]` ([E/2, 7]) for the trimmed
    `data.edge_index`.  `points` [N0, 3] level-0 positions, `se_point_index` [2, Es] the
    level-0 end points of each sub-edge, `se_id` [Es] the superedge each sub-edge belongs to.
    One CUDA pass per superedge (csrc/segment.cu:k_superedge_features) instead of 3 scatters,
    3 gathers and ~15 elementwise launches."""
    if keys is not None and not all(k in keys for k in ('mean_off', 'std_off', 'mean_dist')):
        raise NotImplementedError(
            "'mean_off', 'std_off' and 'mean_dist' must all be computed (same restriction as "
            "the reference, src/transforms/graph.py:996-1001)")
    data.edge_attr = ops.superedge_features(points, se_point_index, se_id,
                                            data.edge_index.shape[1])
    return data


def cluster_point_features(nag, i_level, mean_keys=(), std_keys=(), strict=True):
    """The scatter parts of `_compute_cluster_features` (reference
    src/transforms/graph.py:262-285): `mean_<key>` / `std_<key>` of level-0 point attributes
    over the level-`i_level` clusters (torch_scatter mean / unbiased std).  The sampled
    geometric features of :221-258 need `pgeof` and stay on the host (out of scope)."""
    data = nag[i_level]
    super_index = nag.get_super_index(i_level)
    n = data.num_nodes
    seg = ops.segment_index(super_index, n)
    for key in dict.fromkeys(list(mean_keys) + list(std_keys)):
        f = nag[0][key] if key in nag[0].keys else None
        if f is None:
            if strict:
                raise ValueError(f"No point key `{key}` to build 'mean_{key}' / 'std_{key}'")
            continue
        if key == 'normal' and key in mean_keys:
            raise NotImplementedError("mean_normal needs scatter_mean_orientation (PCA), "
                                      "a preprocessing-only helper outside SURVEY §8")
        mean, std = ops.segment_mean_std(f.float(), super_index, n, want_mean=key in mean_keys,
                                         want_std=key in std_keys, seg=seg)
        if key in mean_keys:
            data[f'mean_{key}'] = mean
        if key in std_keys:
            data[f'std_{key}'] = std
    return nag
