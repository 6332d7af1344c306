"""CPU restatement of the reference's node selection (TEST INFRASTRUCTURE — see
oracle/__init__.py): NAG.select -> Data.select -> Cluster.select -> CSRData.__getitem__,
written with the reference's own algorithm (sort-based `consecutive_cluster`, `torch.where`
edge compaction, repeat_interleave pointer expansion) on plain dicts of CPU tensors.

A level is a dict {key: tensor} whose `sub` entry, when present, is a dict
{'pointers': ..., 'points': ...}; a NAG is (list of levels, start_i_level).  The product
(superpoint_transformer_b200/data) is never imported here.

Pinned by tests/golden/select.pt, generated from the reference's own source files by
oracle/make_golden_select.py through oracle/reference_data.py (`consecutive_cluster` and the
PyG `Data` base class are third-party and restated there => same "leaves unpinned" caveat as
oracle/__init__.py).  Paths below are relative to /root/reference.
"""
import copy

import numpy as np
import torch


def tensor_idx(idx, device='cpu'):
    """src/utils/tensor.py:33-70"""
    if idx is None:
        return None
    if isinstance(idx, int):
        idx = torch.tensor([idx], dtype=torch.long)
    elif isinstance(idx, slice):
        idx = torch.arange(idx.start, idx.stop)
    elif isinstance(idx, np.ndarray):
        idx = torch.from_numpy(idx)
    if idx.dtype == torch.bool:
        idx = torch.where(idx)[0]
    return idx.long()


def is_arange(a, n):
    """src/utils/tensor.py:73-79"""
    return a.equal(torch.arange(n))


def consecutive_cluster(src):
    """torch_geometric.nn.pool.consecutive.consecutive_cluster (PyG 2.3.0; third-party, absent
    from /root/reference): (inverse of the sorted unique values, one position per value)."""
    unique, inv = torch.unique(src, sorted=True, return_inverse=True)
    perm = torch.arange(inv.size(0), dtype=inv.dtype)
    perm = inv.new_empty(unique.size(0)).scatter_(0, inv, perm)
    return inv, perm


def index_select_pointers(pointers, indices):
    """src/data/csr.py:328-356"""
    assert indices.max() <= pointers.shape[0] - 2
    pointers_new = torch.cat([
        torch.zeros(1, dtype=pointers.dtype),
        torch.cumsum(pointers[indices + 1] - pointers[indices], 0)])
    sizes = pointers_new[1:] - pointers_new[:-1]
    val_idx = torch.arange(int(pointers_new[-1]))
    val_idx -= torch.arange(int(pointers_new[-1]) + 1)[pointers_new[:-1]].repeat_interleave(sizes)
    val_idx += pointers[indices].repeat_interleave(sizes)
    return pointers_new, val_idx


def to_super_index(cluster):
    """src/data/cluster.py:67-77"""
    sizes = cluster['pointers'][1:] - cluster['pointers'][:-1]
    out = torch.empty(cluster['points'].shape[0], dtype=torch.long)
    out[cluster['points']] = torch.arange(sizes.shape[0]).repeat_interleave(sizes)
    return out


def cluster_from_dense(indices, points):
    """Cluster(indices, points, dense=True): src/data/csr.py:83-85 + src/utils/sparse.py:23-41.
    The reference's torch.sort is not stable (order inside a cluster unspecified); the stable
    order is used here and the comparisons canonicalise that order."""
    order = torch.sort(indices, stable=True).indices
    s = indices[order]
    pointers = torch.cat([torch.tensor([0]), torch.where(s[1:] > s[:-1])[0] + 1,
                          torch.tensor([s.shape[0]])])
    return {'pointers': pointers, 'points': points[order]}


def cluster_select(cluster, idx, update_sub=True):
    """src/data/cluster.py:79-140 -> cluster, (idx_sub, sub_super)"""
    idx = tensor_idx(idx)
    num_clusters = cluster['pointers'].shape[0] - 1
    if idx is None or is_arange(idx, num_clusters):
        # (the reference returns an EMPTY CSRData on this branch, csr.py:371-378; it is never
        # reached from Data.select / NAG.select, which clone before getting here)
        return copy.deepcopy(cluster), (None, None)
    pointers, val_idx = index_select_pointers(cluster['pointers'], idx)
    out = {'pointers': pointers, 'points': cluster['points'][val_idx]}
    if not update_sub:
        return out, (None, None)
    new_points, perm = consecutive_cluster(out['points'])
    idx_sub = out['points'][perm]
    out['points'] = new_points
    return out, (idx_sub, to_super_index(out))


def _num_nodes(level):
    for k in ('x', 'pos', 'batch'):
        if k in level:
            return level[k].shape[0]
    raise ValueError('level without node-level tensor')


def data_select(level, idx, update_sub=True, update_super=True):
    """src/data/data.py:286-470 -> level, (idx_sub, sub_super), (idx_super, super_sub)"""
    idx = tensor_idx(idx)
    num_nodes = _num_nodes(level)
    if idx is None or is_arange(idx, num_nodes):
        return copy.deepcopy(level), (None, None), (None, None)
    out = {}
    has_edges = 'edge_index' in level and level['edge_index'].shape[1] > 0
    num_edges = level['edge_index'].shape[1] if 'edge_index' in level else 0
    idx_edge = None
    if has_edges:                                                  # data.py:356-371
        reindex = torch.full((num_nodes,), -1, dtype=torch.int64)
        reindex = reindex.scatter_(0, idx, torch.arange(idx.shape[0]))
        edge_index = reindex[level['edge_index'].long()]
        idx_edge = torch.where((edge_index != -1).all(dim=0))[0]
        out['edge_index'] = edge_index[:, idx_edge]
    out_sub = (None, None)
    if 'sub' in level:                                             # data.py:377-386
        out['sub'], out_sub = cluster_select(level['sub'], idx, update_sub=update_sub)
    out_super = (None, None)
    if 'super_index' in level:                                     # data.py:393-417
        out['super_index'] = level['super_index'][idx]
        if update_super:
            new_super_index, perm = consecutive_cluster(out['super_index'])
            idx_super = out['super_index'][perm]
            out['super_index'] = new_super_index
            super_sub = cluster_from_dense(new_super_index, torch.arange(idx.shape[0]))
            out_super = (idx_super, super_sub)
    edge_keys = ['edge_attr'] + [k for k in level if k.startswith('edge_')
                                 and k not in ('edge_index', 'edge_attr')]
    for key, item in level.items():                                # data.py:420-463
        if key in ('edge_index', 'sub', 'super_index', 'neighbor_index', 'neighbor_distance'):
            continue
        is_node_size = item.shape[0] == num_nodes
        is_edge_size = item.shape[0] == num_edges
        if is_node_size and key.startswith('v_edge_'):
            out[key] = item[idx]
        elif has_edges and is_edge_size and key in edge_keys:
            out[key] = item[idx_edge]
        elif is_node_size:
            out[key] = item[idx]
        else:
            out[key] = item.clone()
    return out, out_sub, out_super


def nag_select(levels, start_i_level, i_level, idx):
    """src/data/nag.py:306-399 (without the InstanceData bookkeeping of :385-394).  `levels[j]`
    is absolute level start_i_level + j.  Returns the list of selected levels."""
    idx = tensor_idx(idx)
    absolute = start_i_level + len(levels)
    get = lambda i: levels[i - start_i_level]
    if idx is None or is_arange(idx, _num_nodes(get(i_level))):
        return copy.deepcopy(levels)
    out = [None] * absolute
    out[i_level], out_sub, out_super = data_select(get(i_level), idx, True, True)
    for i in range(i_level - 1, start_i_level - 1, -1):            # nag.py:359-370
        idx_sub, sub_super = out_sub
        out[i], out_sub, _ = data_select(get(i), idx_sub, True, False)
        if sub_super is None:
            out[i].pop('super_index', None)     # the reference stores None here (nag.py:370)
        else:
            out[i]['super_index'] = sub_super
    for i in range(i_level + 1, absolute):                         # nag.py:372-383
        idx_super, super_sub = out_super
        out[i], _, out_super = data_select(get(i), idx_super, False, True)
        if super_sub is None:
            out[i].pop('sub', None)             # idem (nag.py:383)
        else:
            out[i]['sub'] = super_sub
    return out[start_i_level:]


def canonical(level):
    """Level with the points of every `sub` cluster in ascending order (the order inside a
    cluster carries no information: src/data/csr.py:438-443)."""
    out = dict(level)
    if 'sub' in level:
        ptr, pts = level['sub']['pointers'], level['sub']['points']
        sizes = ptr[1:] - ptr[:-1]
        group = torch.arange(sizes.shape[0]).repeat_interleave(sizes)
        order = torch.sort(group * (int(pts.max()) + 1 if pts.numel() else 1) + pts).indices
        out['sub'] = {'pointers': ptr, 'points': pts[order]}
    return out
