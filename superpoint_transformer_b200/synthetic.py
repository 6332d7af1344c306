"""Seeded synthetic NAG generator (SURVEY.md §8d / Appendix C invariants).

Produces NAGs shaped like the reference's preprocessed + sampled batches right
BEFORE the on-device transforms: per level >= 1 a *trimmed* horizontal graph
(i<j once, coalesced, no self-loops) with 7 raw edge attributes, node positions,
normals, log-size features, 12 handcrafted segment features, `super_index` and
the `sub` Cluster.  All laws (level ratios, degree distribution) are this
project's assumptions — the reference publishes none.
"""
import numpy as np
import torch

from .data import Data, NAG, Cluster

__all__ = ['make_nag', 'CONFIGS']

# BASELINE.json configs (level-1.. node counts)
CONFIGS = {
    'cfg1': dict(levels=[1000, 100], mean_degree=8, seed=0),
    'cfg2': dict(levels=[100_000, 20_000, 4_000], mean_degree=16, seed=1),
    'cfg3': dict(levels=[500_000, 100_000, 20_000], mean_degree=16, seed=2),
    'cfg4_scene': dict(levels=[50_000, 10_000, 2_000], mean_degree=16, seed=100),
    'cfg5': dict(levels=[1_000_000, 200_000, 40_000], mean_degree=16, seed=3),
}

NUM_HF_SEGMENT = 12
SEGMENT_HF = ['hf']  # single [N, 12] block standing for the 12 handcrafted columns


def _trimmed_graph(rng, n, mean_degree, k_max=30):
    """Random symmetric-degree ~ clamp(Poisson(mean_degree), 1, k_max) graph,
    returned trimmed: each undirected pair once with i < j, lexicographically
    sorted (what `to_trimmed` + coalesce give, reference src/utils/graph.py:466-521)."""
    if n < 2:
        return np.zeros((2, 0), dtype=np.int64)
    half = np.clip(rng.poisson(mean_degree / 2.0, size=n), 1, k_max // 2)
    src = np.repeat(np.arange(n, dtype=np.int64), half)
    dst = rng.integers(0, n - 1, size=src.shape[0], dtype=np.int64)
    dst = dst + (dst >= src)  # never a self-loop
    lo, hi = np.minimum(src, dst), np.maximum(src, dst)
    key = np.unique(lo * n + hi)
    return np.stack((key // n, key % n))


def _unit(v):
    return v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)


def make_nag(levels, mean_degree=16, seed=0, start_i_level=1, device='cpu',
             dtype=torch.float32):
    """levels: node counts of NAG levels start_i_level, start_i_level+1, ...
    Node ids are a random permutation (worst-case gather locality)."""
    rng = np.random.default_rng(seed)
    datas = []
    for li, n in enumerate(levels):
        d = {}
        d['pos'] = rng.uniform(0, 50, size=(n, 3)).astype(np.float32)
        d['normal'] = _unit(rng.normal(size=(n, 3))).astype(np.float32)
        for k in ('log_length', 'log_surface', 'log_volume', 'log_size'):
            d[k] = rng.normal(size=(n, 1)).astype(np.float32)
        d['hf'] = rng.normal(size=(n, NUM_HF_SEGMENT)).astype(np.float32)
        se = _trimmed_graph(rng, n, mean_degree)
        eh = se.shape[1]
        mean_off = rng.normal(size=(eh, 3)).astype(np.float32)
        std_off = np.abs(rng.normal(size=(eh, 3))).astype(np.float32)
        mean_dist = np.abs(rng.normal(size=(eh, 1))).astype(np.float32)
        d['edge_index'] = se
        d['edge_attr'] = np.concatenate((mean_off, std_off, mean_dist), axis=1)
        if li + 1 < len(levels):
            n_up = levels[li + 1]
            sup = np.concatenate((np.arange(n_up), rng.integers(0, n_up, size=n - n_up)))
            d['super_index'] = rng.permutation(sup).astype(np.int64)
        datas.append(d)
    out = []
    for li, d in enumerate(datas):
        td = Data(**{k: torch.from_numpy(v) for k, v in d.items()})
        if li > 0:
            sup = out[li - 1].super_index
            td.sub = Cluster.from_super_index(sup, levels[li])
        out.append(td)
    # level-1 node_size ~ 1 + Poisson(30) points per superpoint (no level 0 loaded)
    out[0].node_size = torch.from_numpy(1 + rng.poisson(30, size=levels[0])).long()
    nag = NAG(out, start_i_level=start_i_level)
    if device != 'cpu':
        nag = nag.to(device)
    return nag
