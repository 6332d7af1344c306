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


def _trimmed_graph(rng, n, mean_degree, k_max=30, locality=None):
    """Random symmetric-degree ~ clamp(Poisson(mean_degree), 1, k_max) graph,
    returned trimmed: each undirected pair once with i < j, lexicographically
    sorted (what `to_trimmed` + coalesce give, reference src/utils/graph.py:466-521).
    `locality = (order, window)`: neighbours are drawn within `window` places of the node in
    the spatial ordering `order` (a permutation of the node ids) instead of uniformly — the
    spatially coherent variant used where a scene is cut into tiles (cfg 5)."""
    if n < 2:
        return np.zeros((2, 0), dtype=np.int64)
    half = np.clip(rng.poisson(mean_degree / 2.0, size=n), 1, k_max // 2)
    src = np.repeat(np.arange(n, dtype=np.int64), half)
    if locality is not None:
        order, window = locality
        place = np.empty(n, dtype=np.int64)
        place[order] = np.arange(n, dtype=np.int64)
        off = rng.integers(1, window + 1, size=src.shape[0]) * rng.choice((-1, 1), size=src.shape[0])
        tgt = place[src] + off
        tgt = np.where((tgt < 0) | (tgt >= n), place[src] - off, tgt)   # reflect at the ends
        dst = order[np.clip(tgt, 0, n - 1)]
        keep = dst != src
        src, dst = src[keep], dst[keep]
        lo, hi = np.minimum(src, dst), np.maximum(src, dst)
        key = np.unique(lo * n + hi)
        return np.stack((key // n, key % n))
    dst = rng.integers(0, n - 1, size=src.shape[0], dtype=np.int64)
    dst = dst + (dst >= src)  # never a self-loop
    lo, hi = np.minimum(src, dst), np.maximum(src, dst)
    key = np.unique(lo * n + hi)
    return np.stack((key // n, key % n))


def _unit(v):
    return v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)


def make_nag(levels, mean_degree=16, seed=0, start_i_level=1, device='cpu',
             dtype=torch.float32, spatial=False):
    """levels: node counts of NAG levels start_i_level, start_i_level+1, ...
    Node ids are a random permutation (worst-case gather locality).  `spatial=True`: nodes sit
    near their top-level ancestor and edges join nodes whose ancestors are close along x
    (ids stay random): the graph can be cut into spatial tiles that keep ~99 % of the edges."""
    rng = np.random.default_rng(seed)
    datas = []
    sup_all, top_pos, anc_top = [], None, None
    if spatial:
        # hierarchy first (same law as below), then positions / edge locality from the top level
        srng = np.random.default_rng(seed + 7919)
        for li, n in enumerate(levels[:-1]):
            n_up = levels[li + 1]
            sup = np.concatenate((np.arange(n_up), srng.integers(0, n_up, size=n - n_up)))
            sup_all.append(srng.permutation(sup).astype(np.int64))
        top_pos = srng.uniform(0, 50, size=(levels[-1], 3)).astype(np.float32)
        anc_top = [None] * len(levels)
        anc_top[-1] = np.arange(levels[-1])
        for li in range(len(levels) - 2, -1, -1):
            anc_top[li] = anc_top[li + 1][sup_all[li]]
        top_rank = np.empty(levels[-1], dtype=np.int64)
        top_rank[np.argsort(top_pos[:, 0], kind='stable')] = np.arange(levels[-1])
    for li, n in enumerate(levels):
        d = {}
        if spatial:
            d['pos'] = (top_pos[anc_top[li]] +
                        rng.normal(scale=0.5, size=(n, 3))).astype(np.float32)
        else:
            d['pos'] = rng.uniform(0, 50, size=(n, 3)).astype(np.float32)
        d['normal'] = _unit(rng.normal(size=(n, 3))).astype(np.float32)
        for k in ('log_length', 'log_surface', 'log_volume', 'log_size'):
            d[k] = rng.normal(size=(n, 1)).astype(np.float32)
        d['hf'] = rng.normal(size=(n, NUM_HF_SEGMENT)).astype(np.float32)
        if spatial:
            order = np.argsort(top_rank[anc_top[li]] + rng.uniform(0, 1, size=n), kind='stable')
            se = _trimmed_graph(rng, n, mean_degree, locality=(order, max(n // 400, 20)))
        else:
            se = _trimmed_graph(rng, n, mean_degree)
        eh = se.shape[1]
        mean_off = rng.normal(size=(eh, 3)).astype(np.float32)
        std_off = np.abs(rng.normal(size=(eh, 3))).astype(np.float32)
        mean_dist = np.abs(rng.normal(size=(eh, 1))).astype(np.float32)
        d['edge_index'] = se
        d['edge_attr'] = np.concatenate((mean_off, std_off, mean_dist), axis=1)
        if li + 1 < len(levels):
            if spatial:
                d['super_index'] = sup_all[li]
            else:
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
