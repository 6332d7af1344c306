"""CPU restatement of the reference's per-segment sampling (TEST INFRASTRUCTURE — see
oracle/__init__.py): `sparse_sample` (src/utils/sparse.py:142-243) and the node weights of
`SampleSegments` (src/transforms/sampling.py:771-798), on CPU tensors.

Pinned by tests/golden/sampling.pt (oracle/make_golden_select.py, from the reference's own
files): the number of samples per segment (deterministic) bit-exactly, the SampleSegments
output under a fixed torch seed bit-exactly.  Which elements a segment keeps is random: the
product is held to the sampling LAW (distinct elements of the right segment, every subset
equally likely), not to the reference's random stream.  Paths relative to /root/reference.
"""
import torch


def sampling_counts(size, n_max, n_min):
    """src/utils/sparse.py:176-189 (fp32 tanh heuristic, then clamp to [n_min, size])"""
    if n_max > 0:
        n_samples = (n_max * torch.tanh(size / n_max)).floor().long()
    else:
        n_samples = size.sqrt().round().long()
    return n_samples.clamp(min=n_min).clamp(max=size)


def sparse_sample(idx, n_max=32, n_min=1, mask=None, generator=None):
    """src/utils/sparse.py:142-243 -> (idx_samples, ptr_samples); `mask`: LongTensor of
    positions or BoolTensor."""
    assert 0 <= n_min <= n_max
    size = idx.bincount()
    num_segments = int(idx.max()) + 1
    n_samples = sampling_counts(size, n_max, n_min)
    sample_idx = torch.arange(idx.shape[0])
    if mask is not None:
        if mask.dtype == torch.bool:
            mask = torch.where(mask)[0]
        sample_idx = sample_idx[mask]
        idx = idx[mask]
        size = idx.bincount(minlength=num_segments)
        n_samples = n_samples.clamp(max=size)
    perm = torch.randperm(sample_idx.shape[0], generator=generator)     # :218-220
    idx, sample_idx = idx[perm], sample_idx[perm]
    idx, order = idx.sort()                                              # :225-226
    sample_idx = sample_idx[order]
    offset = torch.cat((torch.zeros(1, dtype=torch.long), size[:-1])).cumsum(0)
    ptr = torch.cat((torch.zeros(1, dtype=torch.long), n_samples)).cumsum(0)
    take = torch.cat([torch.arange(int(o), int(o) + int(n))
                      for o, n in zip(offset, n_samples)]) if len(n_samples) else ptr[:0]
    return sample_idx[take], ptr


def segment_weights(y, node_size, by_size, by_class):
    """src/transforms/sampling.py:771-798: sampling weight of every node of a level from its
    level-0 size and its label histogram `y` [N, classes]."""
    weights = torch.ones(node_size.shape[0])
    if by_size:
        size_weights = node_size ** 0.333
        size_weights /= size_weights.sum()
        weights += size_weights
    if by_class and y is not None:
        counts = y.sum(dim=0).sqrt()
        scores = 1 / (counts + 1)
        scores /= scores.sum()
        mask = y.gt(0)
        class_weights = (mask * scores.view(1, -1)).max(dim=1).values
        class_weights /= class_weights.sum()
        weights += class_weights.squeeze()
    weights /= weights.sum()
    return weights


def radius_nodes(pos, seeds, r, k_max=10000, batch=None, cylindrical=False):
    """Neighbour search of SampleRadiusSubgraphs: src/transforms/sampling.py:1196-1231 with
    knn_brute_force (src/utils/neighbors.py:245-295) -> sorted unique node ids."""
    mask = torch.tensor([[1, 1, 0 if cylindrical else 1]])
    x_search, x_query = pos * mask, pos[seeds] * mask
    if batch is not None:
        hi = max(x_search[:, 2].max(), x_query[:, 2].max())
        lo = min(x_search[:, 2].min(), x_query[:, 2].min())
        z_offset = hi - lo + r + 1
        off_s, off_q = torch.zeros_like(x_search), torch.zeros_like(x_query)
        off_s[:, 2] = batch * z_offset
        off_q[:, 2] = batch[seeds] * z_offset
        x_search, x_query = x_search + off_s, x_query + off_q
    distances = (x_search.unsqueeze(0) - x_query.unsqueeze(1)).norm(dim=2)
    distances, neighbors = distances.sort(dim=1)
    distances, neighbors = distances[:, :k_max], neighbors[:, :k_max]
    neighbors = neighbors.clone()
    neighbors[distances > r] = -1
    return neighbors[neighbors >= 0].unique()


def khop_nodes(edge_index, seeds, hops, num_nodes):
    """torch_geometric.utils.k_hop_subgraph(seeds, hops, to_undirected(edge_index))[0] (PyG
    2.3.0, third-party): src/transforms/sampling.py:1080-1091."""
    row = torch.cat((edge_index[0], edge_index[1]))
    col = torch.cat((edge_index[1], edge_index[0]))
    subsets = [seeds.view(-1)]
    for _ in range(hops):
        node_mask = torch.zeros(num_nodes, dtype=torch.bool)
        node_mask[subsets[-1]] = True
        subsets.append(col[node_mask[row]])
    return torch.cat(subsets).unique()
