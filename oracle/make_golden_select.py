"""Generate tests/golden/select.pt from the REFERENCE's own data containers (build container
only):   python -m oracle.make_golden_select

Runs `NAG.select`, `Data.select`, `Cluster.select`, `CSRData.index_select_pointers` of
/root/reference/src/data (loaded verbatim by oracle/reference_data.py) on small seeded nested
partitions and stores inputs and outputs as plain dicts of tensors.  TEST INFRASTRUCTURE.
"""
import os

import numpy as np
import torch

from . import reference_data as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   'tests', 'golden', 'select.pt')


def make_levels(sizes, seed, start=0):
    """Plain-dict levels of a random nested partition: every parent keeps >= 1 child."""
    g = torch.Generator().manual_seed(seed)
    levels = []
    for li, n in enumerate(sizes):
        lv = {'pos': torch.rand(n, 3, generator=g),
              'x': torch.rand(n, 5, generator=g),
              'y': torch.randint(0, 9, (n, 7), generator=g),
              'node_size': torch.randint(1, 50, (n,), generator=g)}
        if li + start >= 1:
            e = 6 * n
            lv['edge_index'] = torch.randint(0, n, (2, e), generator=g)
            lv['edge_attr'] = torch.rand(e, 4, generator=g)
        if li + 1 < len(sizes):
            up = sizes[li + 1]
            sup = torch.cat((torch.arange(up), torch.randint(0, up, (n - up,), generator=g)))
            lv['super_index'] = sup[torch.randperm(n, generator=g)]
            lv['v_edge_attr'] = torch.rand(n, 3, generator=g)
        levels.append(lv)
    for li in range(len(sizes)):
        if li == 0 and start == 0:
            continue
        if li == 0:      # nano NAG: `sub` of the first loaded level points at absent atoms
            n_atoms = 5 * sizes[0]
            sup = torch.cat((torch.arange(sizes[0]),
                             torch.randint(0, sizes[0], (n_atoms - sizes[0],), generator=g)))
            sup = sup[torch.randperm(n_atoms, generator=g)]
        else:
            sup = levels[li - 1]['super_index']
        order = torch.sort(sup, stable=True).indices
        ptr = torch.zeros(sizes[li] + 1, dtype=torch.long)
        ptr[1:] = torch.bincount(sup, minlength=sizes[li]).cumsum(0)
        levels[li]['sub'] = {'pointers': ptr, 'points': order}
    return levels


def to_reference(ns, levels, start):
    datas = []
    for lv in levels:
        kw = {k: v.clone() for k, v in lv.items() if k != 'sub'}
        if 'sub' in lv:
            kw['sub'] = ns.Cluster(lv['sub']['pointers'].clone(), lv['sub']['points'].clone())
        datas.append(ns.Data(**kw))
    return ns.NAG(datas, start_i_level=start)


def cluster_dict(c):
    return {'pointers': c.pointers.clone(), 'points': c.points.clone()}


def level_dict(ns, data):
    out = {}
    for k, v in data:
        if v is None or k == 'num_nodes':
            continue
        out[k] = cluster_dict(v) if isinstance(v, ns.CSRData) else v.clone()
    return out


def make_idx(kind, n, g):
    keep = max(1, (2 * n) // 3)
    perm = torch.randperm(n, generator=g)
    if kind == 'sorted':
        return perm[:keep].sort().values
    if kind == 'shuffled':
        return perm[:keep]
    if kind == 'mask':
        m = torch.zeros(n, dtype=torch.bool)
        m[perm[:keep]] = True
        return m
    if kind == 'int':
        return int(perm[0])
    if kind == 'numpy':
        return perm[:keep].numpy().copy()
    if kind == 'arange':
        return torch.arange(n)
    if kind == 'few':
        return perm[:max(1, n // 10)]
    raise ValueError(kind)


def main():
    ns = R.load_data()
    g = torch.Generator().manual_seed(1234)
    out = {'nags': {}, 'nag_cases': [], 'data_cases': [], 'cluster_cases': [],
           'pointer_cases': [], 'consecutive_cases': [], 'reference_drops': []}
    specs = {'full4': ([400, 80, 16, 4], 0, 11), 'nano3': ([120, 30, 6], 1, 12),
             'two': ([60, 9], 0, 13)}
    for name, (sizes, start, seed) in specs.items():
        levels = make_levels(sizes, seed, start)
        out['nags'][name] = {'start': start, 'levels': levels}
        nag = to_reference(ns, levels, start)
        for li, n in enumerate(sizes):
            i_level = li + start
            for kind in ('sorted', 'shuffled', 'mask', 'int', 'numpy', 'arange', 'few'):
                idx = make_idx(kind, n, g)
                res = nag.select(i_level, idx)
                res_levels = [level_dict(ns, res[i]) for i in range(start, start + len(sizes))]
                # the reference hands None across levels when a neighbouring level needed no
                # re-indexing (nag.py:370, 383) and thereby DROPS that level's super_index /
                # sub: those cases are stored apart (the product keeps the attribute)
                dropped = any(('super_index' in a) != ('super_index' in b) or
                              ('sub' in a) != ('sub' in b)
                              for a, b in zip(levels, res_levels))
                case = {'nag': name, 'i_level': i_level, 'kind': kind,
                        'idx': torch.from_numpy(idx) if kind == 'numpy' else idx,
                        'out': res_levels}
                out['reference_drops' if dropped else 'nag_cases'].append(case)

        # Data.select alone, every flag combination, on a middle level
        li = 1 if len(sizes) > 2 else 0
        data = nag[li + start]
        for upd_sub in (True, False):
            for upd_super in (True, False):
                idx = make_idx('shuffled', sizes[li], g)
                d, (idx_sub, sub_super), (idx_super, super_sub) = data.select(
                    idx, update_sub=upd_sub, update_super=upd_super)
                out['data_cases'].append({
                    'nag': name, 'i_level': li + start, 'idx': idx, 'update_sub': upd_sub,
                    'update_super': upd_super, 'out': level_dict(ns, d), 'idx_sub': idx_sub,
                    'sub_super': sub_super, 'idx_super': idx_super,
                    'super_sub': None if super_sub is None else cluster_dict(super_sub)})

        # Cluster.select / CSRData.index_select_pointers
        for li in range(len(sizes)):
            if 'sub' not in levels[li]:
                continue
            cl = nag[li + start].sub
            for kind in ('sorted', 'shuffled', 'few'):
                idx = make_idx(kind, sizes[li], g)
                for upd in (True, False):
                    c, (idx_sub, sub_super) = cl.select(idx, update_sub=upd)
                    out['cluster_cases'].append({
                        'nag': name, 'i_level': li + start, 'idx': idx, 'update_sub': upd,
                        'out': cluster_dict(c), 'idx_sub': idx_sub, 'sub_super': sub_super})
                p, v = ns.index_select_pointers(cl.pointers, idx)
                out['pointer_cases'].append({'pointers': cl.pointers.clone(), 'idx': idx,
                                             'pointers_new': p, 'val_idx': v})

    # consecutive_cluster as restated for the loader (third-party; documents the contract the
    # product's bitmap + scan relabel is held to)
    for n, hi in ((1, 1), (50, 10), (300, 1000), (2000, 500)):
        src = torch.randint(0, hi, (n,), generator=g)
        inv, perm = ns.consecutive_cluster(src)
        out['consecutive_cases'].append({'src': src, 'num_ids': hi, 'inv': inv,
                                         'unique': src[perm]})
    torch.save(out, OUT)
    print(OUT, {k: len(v) for k, v in out.items()})
    make_sampling(ns, out['nags'])


def make_sampling(ns, nags):
    """tests/golden/sampling.pt: `sparse_sample` (deterministic part: samples per segment; the
    drawn indices are kept for the record) and SampleSegments under a fixed torch seed."""
    g = torch.Generator().manual_seed(4321)
    out = {'sparse': [], 'segments': [], 'restrict': [], 'edges': [], 'subgraphs': []}
    for n, hi, n_max, n_min, masked in (
            (1, 1, 4, 1, False), (500, 40, 8, 2, False), (500, 40, 8, 2, True),
            (3000, 7, 32, 16, False), (3000, 7, 32, 1, True), (2000, 300, 4, 4, False),
            (2000, 50, 0, 0, False), (5000, 3, 64, 1, False), (800, 100, 32, 16, 'bool')):
        idx = torch.cat((torch.arange(hi), torch.randint(0, hi, (n - hi,), generator=g)))
        idx = idx[torch.randperm(n, generator=g)]
        mask = None
        if masked:
            mask = torch.randperm(n, generator=g)[:n // 2]
            if masked == 'bool':
                m = torch.zeros(n, dtype=torch.bool)
                m[mask] = True
                mask = m
        samples, ptr = ns.sparse_sample(idx, n_max=n_max, n_min=n_min, mask=mask,
                                        return_pointers=True)
        out['sparse'].append({'idx': idx, 'n_max': n_max, 'n_min': n_min, 'mask': mask,
                              'idx_samples': samples, 'ptr_samples': ptr})
    for name, spec in nags.items():
        if spec['start'] != 0:
            continue
        for ratio, by_size, by_class, seed in ((0.3, False, False, 1), (0.25, True, False, 2),
                                               ([0.2, 0.5, 0.0][:len(spec['levels']) - 1],
                                                True, True, 3)):
            nag = to_reference(ns, spec['levels'], spec['start'])
            torch.manual_seed(seed)
            res = ns.SampleSegments(ratio=ratio, by_size=by_size, by_class=by_class)(nag)
            out['segments'].append({
                'nag': name, 'ratio': ratio, 'by_size': by_size, 'by_class': by_class,
                'seed': seed, 'out': [level_dict(ns, res[i]) for i in range(res.num_levels)]})
        for level, num_nodes, num_edges, seed in (('1+', 10, 40, 4), (1, 25, 0, 5),
                                                  ('all', 30, 100, 6)):
            nag = to_reference(ns, spec['levels'], spec['start'])
            torch.manual_seed(seed)
            res = ns.NAGRestrictSize(level=level, num_nodes=num_nodes, num_edges=num_edges)(nag)
            out['restrict'].append({
                'nag': name, 'level': level, 'num_nodes': num_nodes, 'num_edges': num_edges,
                'seed': seed, 'out': [level_dict(ns, res[i]) for i in range(res.num_levels)]})
        for n_min, n_max in ((1, 2), (2, 4), (0, 3)):
            nag = to_reference(ns, spec['levels'], spec['start'])
            res = ns.SampleEdges(level='1+', n_min=n_min, n_max=n_max)(nag)
            out['edges'].append({
                'nag': name, 'level': '1+', 'n_min': n_min, 'n_max': n_max,
                'degree': [torch.bincount(res[i].edge_index[0], minlength=res[i].num_nodes)
                           if res[i].edge_index is not None else None
                           for i in range(res.num_levels)]})
        # subgraph sampling: seeds from torch's global CPU generator (fixed seed), then the
        # reference's neighbour search + NAG.select.  `batch`: two halves of every level along x
        for kind, kw in (('radius', dict(r=0.35, i_level=1, k=1)),
                         ('radius', dict(r=0.3, i_level=1, k=3, by_size=True)),
                         ('radius', dict(r=0.25, i_level=2, k=2, cylindrical=True,
                                         by_class=True)),
                         ('radius', dict(r=0.2, i_level=1, k=3, k_max=101)),
                         ('radius_batch', dict(r=0.4, i_level=1, k=2, use_batch=True)),
                         ('khop', dict(hops=1, i_level=1, k=2)),
                         ('khop', dict(hops=2, i_level=2, k=1, by_size=True))):
            if kw['i_level'] >= len(spec['levels']):
                continue
            nag = to_reference(ns, spec['levels'], spec['start'])
            batches = None
            if kind == 'radius_batch':
                # consistent batch ids: the top level splits along x, children inherit
                top = len(spec['levels']) - 1
                b = (spec['levels'][top]['pos'][:, 0] > 0.5).long()
                batches = [None] * (top + 1)
                batches[top] = b
                for i in range(top - 1, -1, -1):
                    batches[i] = batches[i + 1][spec['levels'][i]['super_index']]
                for i in range(top + 1):
                    nag[i].batch = batches[i].clone()
            cls = ns.SampleKHopSubgraphs if kind == 'khop' else ns.SampleRadiusSubgraphs
            torch.manual_seed(17)
            res = cls(disjoint=False, **kw)(nag)
            out['subgraphs'].append({
                'nag': name, 'kind': kind, 'kw': kw, 'seed': 17, 'batches': batches,
                'out': [level_dict(ns, res[i]) for i in range(res.num_levels)]})
    path = OUT.replace('select.pt', 'sampling.pt')
    torch.save(out, path)
    print(path, {k: len(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
