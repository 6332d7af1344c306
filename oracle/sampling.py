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
