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
           'NAGAddSelfLoops', 'ON_THE_FLY_HORIZONTAL_FEATURES', 'ON_THE_FLY_VERTICAL_FEATURES']

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
    Only the full default key set is built by the kernel.

    `csr_order=True` (an extension; the model is invariant to the order of the edges)
    additionally groups the output edges by source node, stably, so that the attention
    blocks consume `edge_attr` in place instead of permuting it (and un-permuting its
    gradient) once per stage and step.
    """

    def __init__(self, keys=None, use_mean_normal=False, add_self_loops=False,
                 csr_order=False):
        keys = ON_THE_FLY_HORIZONTAL_FEATURES if keys is None else list(keys)
        if sorted(keys) != sorted(ON_THE_FLY_HORIZONTAL_FEATURES):
            raise NotImplementedError(
                "the CUDA edge-feature kernel builds the full 18-column default set")
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
            if self.csr_order:
                seg = ops.group_index(ei[0], d.num_nodes)
                ei = ei.index_select(1, seg.perm.long())
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
