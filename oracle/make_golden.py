"""Generate tests/golden/*.pt by running the REFERENCE's own source
(oracle/reference_shim.py) on seeded inputs.  Build-container only:

    python -m oracle.make_golden

Each fixture stores the inputs, the reference parameters (state_dict with the
reference key names), the reference outputs and input/parameter gradients of
sum(out * probe) in fp32 — plus the fp64 outputs of the same run as "truth".
tests/test_oracle_golden.py replays them against oracle/path.py (CPU, anywhere);
tests/test_gpu_parity.py replays them against the CUDA path.
"""
import os
import sys
import zlib

import torch

from . import reference_shim as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden')


def _graph(gen, n, e_half):
    """symmetric graph in the reference order [i<j | j>i | self-loops]"""
    i = torch.randint(0, n, (e_half * 2,), generator=gen)
    j = torch.randint(0, n, (e_half * 2,), generator=gen)
    lo, hi = torch.minimum(i, j), torch.maximum(i, j)
    keep = lo != hi
    uid = torch.unique(lo[keep] * n + hi[keep])[:e_half]
    se = torch.stack((uid // n, uid % n))
    loops = torch.arange(n)
    return torch.cat((se, se.flip(0), torch.stack((loops, loops))), dim=1), se


def _run(module, args, kwargs, wrt, probe_gen):
    """forward + backward of sum(out*probe); returns outputs, grads"""
    for t in wrt:
        t.requires_grad_(True)
    out = module(*args, **kwargs)
    outs = out if isinstance(out, (tuple, list)) else (out,)
    first = outs[0]
    probe = torch.randn(first.shape, generator=probe_gen, dtype=torch.float32).to(first.dtype)
    (first * probe).sum().backward()
    grads = [t.grad.detach().clone() if t.grad is not None else None for t in wrt]
    pgrads = {k: p.grad.detach().clone() for k, p in module.named_parameters()
              if p.grad is not None} if hasattr(module, 'named_parameters') else {}
    return [o.detach().clone() if torch.is_tensor(o) else o for o in outs], probe, grads, pgrads


def attention_cases(ref):
    cases = {}
    specs = {
        'kqv': dict(k_rpe=True, q_rpe=True, v_rpe=True),
        'share_minus': dict(k_rpe=True, q_rpe=True, qk_share_rpe=True, q_on_minus_rpe=True,
                            v_rpe=True),
        'heads_share': dict(k_rpe=True, q_rpe=True, v_rpe=True, heads_share_rpe=True),
        'no_rpe': dict(),
        'k_only_dplusg': dict(k_rpe=True, qk_scale='d+g'),
        'scale_g': dict(q_rpe=True, v_rpe=True, qk_scale='g'),
        'scale_const': dict(k_rpe=True, q_rpe=True, qk_scale=0.37),
    }
    shapes = {'c32h4': (60, 32, 4, 4, 8), 'c64h16': (48, 64, 16, 4, 32), 'c16h16d2': (40, 16, 16, 2, 16),
              'c128h4': (64, 128, 4, 4, 32)}
    for sname, (N, C, H, D, F) in shapes.items():
        for vname, kw in specs.items():
            if sname != 'c32h4' and vname not in ('kqv', 'heads_share'):
                continue
            gen = torch.Generator().manual_seed(zlib.crc32(f'{sname}/{vname}'.encode()) % 100000)
            ei, _ = _graph(gen, N, N * 4)
            E = ei.shape[1]
            x = torch.randn(N, C, generator=gen)
            ea = torch.randn(E, F, generator=gen)
            torch.manual_seed(1234)
            blk = ref.SelfAttentionBlock(C, num_heads=H, qk_dim=D, in_rpe_dim=F, out_dim=C, **kw)
            blk.apply(ref.init_weights)
            for p in blk.parameters():  # non-zero biases so they are exercised
                if p.dim() == 1:
                    p.data.normal_(0, 0.1, generator=gen)
            xin, ein = x.clone(), ea.clone()
            outs, probe, grads, pgrads = _run(blk, (xin, ei, ein), {}, [xin, ein],
                                              torch.Generator().manual_seed(7))
            blk64 = ref.SelfAttentionBlock(C, num_heads=H, qk_dim=D, in_rpe_dim=F, out_dim=C,
                                           **kw).double()
            blk64.load_state_dict({k: v.double() for k, v in blk.state_dict().items()})
            with torch.no_grad():
                out64 = blk64(x.double(), ei, ea.double())
            cases[f'{sname}_{vname}'] = dict(
                cfg=dict(dim=C, num_heads=H, qk_dim=D, in_rpe_dim=F, **kw),
                sd={k: v.detach().clone() for k, v in blk.state_dict().items()},
                x=x, edge_index=ei, edge_attr=ea, out=outs[0], out64=out64, probe=probe,
                dx=grads[0], dedge_attr=grads[1], dparams=pgrads)
    return cases


def segment_cases(ref):
    cases = {}
    gen = torch.Generator().manual_seed(11)
    Nc, Np, C = 300, 40, 32
    idx = torch.randint(0, Np - 3, (Nc,), generator=gen)  # 3 empty parents at the end
    x = torch.randn(Nc, C, generator=gen)
    for name, cls in (('max', ref.MaxPool), ('min', ref.MinPool), ('mean', ref.MeanPool),
                      ('sum', ref.SumPool)):
        xin = x.clone()
        outs, probe, grads, _ = _run(cls(), (xin, None, idx), dict(num_pool=Np), [xin],
                                     torch.Generator().manual_seed(3))
        cases[f'pool_{name}'] = dict(x=x, index=idx, num_pool=Np, out=outs[0], probe=probe,
                                     dx=grads[0])
    # unpool
    xp = torch.randn(Np, C, generator=gen)
    xin = xp.clone()
    outs, probe, grads, _ = _run(ref.IndexUnpool(), (xin, idx), {}, [xin],
                                 torch.Generator().manual_seed(4))
    cases['unpool'] = dict(x=xp, index=idx, out=outs[0], probe=probe, dx=grads[0])
    # unit sphere norm
    pos = torch.rand(Nc, 3, generator=gen) * 50
    w = torch.randint(1, 60, (Nc,), generator=gen)
    usn = ref.UnitSphereNorm()
    p1, d1 = usn(pos, idx, w=w, num_super=Np)
    p2, d2 = usn(pos, idx, w=None, num_super=Np)
    p3, d3 = usn(pos, None, w=w)
    cases['unit_sphere'] = dict(pos=pos, index=idx, w=w, num_super=Np, pos_w=p1, diam_w=d1,
                                pos_nw=p2, diam_nw=d2, pos_none=p3, diam_none=d3)
    # GraphNorm (leaf restatement run through the reference MLP glue)
    B = 3
    batch = torch.sort(torch.randint(0, B, (Nc,), generator=gen)).values
    torch.manual_seed(5)
    m = ref.MLP([C, 48, 24], norm=ref.GraphNorm)
    for p in m.parameters():
        if p.dim() == 1:
            p.data.normal_(1.0, 0.2, generator=gen)
    xin = x.clone()
    outs, probe, grads, pgrads = _run(m, (xin,), dict(batch=batch), [xin],
                                      torch.Generator().manual_seed(6))
    cases['mlp_graphnorm'] = dict(x=x, batch=batch, dims=[C, 48, 24],
                                  sd={k: v.detach().clone() for k, v in m.state_dict().items()},
                                  out=outs[0], probe=probe, dx=grads[0], dparams=pgrads)
    # same with an UNSORTED batch vector (edge-level norm index case)
    ub = batch[torch.randperm(Nc, generator=gen)]
    m.zero_grad()
    xin = x.clone()
    outs, probe, grads, pgrads = _run(m, (xin,), dict(batch=ub), [xin],
                                      torch.Generator().manual_seed(6))
    cases['mlp_graphnorm_unsorted'] = dict(
        x=x, batch=ub, dims=[C, 48, 24],
        sd={k: v.detach().clone() for k, v in m.state_dict().items()}, out=outs[0],
        probe=probe, dx=grads[0], dparams=pgrads)
    return cases


def norm_pool_cases(ref):
    """GroupNorm (reference source), LayerNorm(mode='graph') (PyG leaf), StdPool"""
    cases = {}
    gen = torch.Generator().manual_seed(21)
    N, C, B = 500, 32, 3
    x = torch.randn(N, C, generator=gen) * 2 + 1.5
    batch = torch.sort(torch.randint(0, B, (N,), generator=gen)).values
    for name, mk, kw in (
            ('groupnorm_g4', lambda: ref.GroupNorm(C, num_groups=4), dict(batch=batch)),
            ('groupnorm_g1', lambda: ref.GroupNorm(C, num_groups=1), dict(batch=batch)),
            ('groupnorm_g8_nobatch', lambda: ref.GroupNorm(C, num_groups=8), dict()),
            ('layernorm_graph', lambda: ref.LayerNorm(C, mode='graph'), dict(batch=batch)),
            ('layernorm_graph_nobatch', lambda: ref.LayerNorm(C, mode='graph'), dict())):
        m = mk()
        m.weight.data.normal_(1.0, 0.3, generator=gen)
        m.bias.data.normal_(0.0, 0.3, generator=gen)
        xin = x.clone()
        outs, probe, grads, pgrads = _run(m, (xin,), kw, [xin], torch.Generator().manual_seed(8))
        cases[name] = dict(x=x, batch=kw.get('batch'), weight=m.weight.detach().clone(),
                           bias=m.bias.detach().clone(), out=outs[0], probe=probe, dx=grads[0],
                           dparams=pgrads)
    Nc, Np = 400, 50
    idx = torch.randint(0, Np - 4, (Nc,), generator=gen)  # 4 empty parents
    idx[:3] = Np - 5                                      # ... and a parent with 3 equal rows
    xs = torch.randn(Nc, C, generator=gen)
    xs[:3] = xs[0]
    xin = xs.clone()
    outs, probe, grads, _ = _run(ref.StdPool(), (xin, None, idx), dict(num_pool=Np), [xin],
                                 torch.Generator().manual_seed(9))
    cases['pool_std'] = dict(x=xs, index=idx, num_pool=Np, out=outs[0], probe=probe, dx=grads[0])
    # TransformerBlock with the reference's code-default norm = PyG LayerNorm(mode='graph')
    # (src/nn/transformer.py:137) and with GroupNorm, on a 3-graph batch
    Nn, Cb, Fb = 260, 32, 8
    ei, _ = _graph(gen, Nn, Nn * 3)
    bidx = torch.sort(torch.randint(0, 3, (Nn,), generator=gen)).values
    xb = torch.randn(Nn, Cb, generator=gen)
    eab = torch.randn(ei.shape[1], Fb, generator=gen)
    import functools
    for name, norm in (('block_layernorm_graph', None),
                       ('block_groupnorm4', functools.partial(ref.GroupNorm, num_groups=4))):
        torch.manual_seed(77)
        kw = dict(num_heads=4, qk_dim=4, in_rpe_dim=Fb, k_rpe=True, q_rpe=True, v_rpe=True,
                  ffn_ratio=1)
        vh = ref.VersionHolder('3.0.0')
        blk = ref.TransformerBlock(Cb, version_holder=vh, **kw) if norm is None else \
            ref.TransformerBlock(Cb, norm=norm, version_holder=vh, **kw)
        blk.apply(ref.init_weights)
        for p_ in blk.parameters():
            if p_.dim() == 1:
                p_.data.normal_(0.5, 0.3, generator=gen)
        blk.eval()
        xin, ein = xb.clone(), eab.clone()
        outs, probe, grads, pgrads = _run(blk, (xin, bidx, ei, ein), {}, [xin, ein],
                                          torch.Generator().manual_seed(10))
        cases[name] = dict(x=xb, batch=bidx, edge_index=ei, edge_attr=eab, cfg=kw,
                           sd={k: v.detach().clone() for k, v in blk.state_dict().items()},
                           out=outs[0], probe=probe, dx=grads[0], dedge_attr=grads[1],
                           dparams=pgrads)
    return cases


SPT_CFG = dict(
    nano=True, segment_hf=['hf'], down_dim=[32, 32, 32],
    down_in_mlp=[[3 + 1 + 16, 32, 32], [3 + 1 + 16 + 32, 32, 32], [3 + 1 + 16 + 32, 32, 32]],
    down_num_heads=4, down_num_blocks=2, down_ffn_ratio=1, up_dim=[32, 32],
    up_in_mlp=[[20 + 32 + 32, 32, 32], [20 + 32 + 32, 32, 32]], up_num_heads=4,
    up_num_blocks=1, node_mlp=[12, 16, 16], h_edge_mlp=[18, 16, 16], qk_dim=4,
    in_rpe_dim=16, k_rpe=True, q_rpe=True, v_rpe=True, no_ffn=True,
    use_diameter_parent=True, pool='max')


def _prep_nag(levels, seed):
    """synthetic NAG -> after the on-device transforms, via the REFERENCE functions"""
    from superpoint_transformer_b200.synthetic import make_nag
    ref = R.load()
    nag = make_nag(levels, mean_degree=8, seed=seed)
    raw = {l: (nag[l].edge_index.clone(), nag[l].edge_attr.clone()) for l in nag.level_range}
    for l in nag.level_range:
        nag._list[l - nag.start_i_level] = ref.on_the_fly_horizontal_edge_features(nag[l])
    nag = ref.NAGAddSelfLoops()(nag)
    # node_size by the reference's rule (NAG.get_sub_size, src/data/nag.py:94-110)
    size = nag[1].node_size
    for l in range(2, nag.absolute_num_levels):
        size = torch.zeros(nag[l].num_nodes, dtype=torch.long).index_add_(
            0, nag[l - 1].super_index, size)
        nag[l].node_size = size
    return nag, raw


def spt_case(ref):
    nag, raw = _prep_nag([240, 48, 10], seed=5)
    torch.manual_seed(99)
    net = ref.SPT(mlp_norm=ref.GraphNorm, norm=ref.GraphNorm, **SPT_CFG)
    net.apply(ref.init_weights)
    gen = torch.Generator().manual_seed(8)
    for p in net.parameters():
        if p.dim() == 1:
            p.data.add_(torch.randn(p.shape, generator=gen) * 0.1)
    run = nag.clone()
    out = net(run)
    probe = torch.randn(out.shape, generator=torch.Generator().manual_seed(9))
    (out * probe).sum().backward()
    pgrads = {k: p.grad.detach().clone() for k, p in net.named_parameters()
              if p.grad is not None}
    net64 = ref.SPT(mlp_norm=ref.GraphNorm, norm=ref.GraphNorm, **SPT_CFG).double()
    net64.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    run64 = nag.clone()
    for d in run64:
        for k in d.keys:
            v = d[k]
            if torch.is_tensor(v) and v.is_floating_point():
                d[k] = v.double()
    out64 = net64(run64)
    (out64 * probe.double()).sum().backward()
    pgrads64 = {k: p.grad.detach().clone() for k, p in net64.named_parameters()
                if p.grad is not None}
    out64 = out64.detach()
    levels = {}
    for l in nag.level_range:
        d = nag[l]
        levels[l] = {k: d[k] for k in d.keys if torch.is_tensor(d[k])}
        if d.sub is not None:
            levels[l]['sub_pointers'] = d.sub.pointers
            levels[l]['sub_points'] = d.sub.points
        levels[l]['raw_edge_index'], levels[l]['raw_edge_attr'] = raw[l]
    return dict(cfg=SPT_CFG, sd={k: v.detach().clone() for k, v in net.state_dict().items()},
                levels=levels, start_i_level=nag.start_i_level, out=out.detach(),
                out64=out64, probe=probe, dparams=pgrads, dparams64=pgrads64)


def stage_cases(ref):
    """DownNFuseStage / UpNFuseStage with 'mean' pooling + FFN + post-norm variants"""
    nag, _ = _prep_nag([200, 40], seed=12)
    cases = {}
    gen = torch.Generator().manual_seed(21)
    C = 32
    d1, d2 = nag[1], nag[2]
    x_child = torch.randn(d1.num_nodes, C, generator=gen)
    x_parent = torch.randn(d2.num_nodes, 12, generator=gen)
    ni2 = torch.zeros(d2.num_nodes, dtype=torch.long)
    ni1 = torch.zeros(d1.num_nodes, dtype=torch.long)
    for name, kw in (('down_mean_ffn', dict(pool='mean', no_ffn=False, ffn_ratio=2)),
                     ('down_max_postnorm', dict(pool='max', no_ffn=False, ffn_ratio=1,
                                                pre_norm=False))):
        torch.manual_seed(31)
        st = ref.DownNFuseStage(
            C, num_blocks=2, num_heads=4, in_mlp=[3 + 1 + 12 + C, C, C], mlp_norm=ref.GraphNorm,
            norm=ref.GraphNorm, qk_dim=4, in_rpe_dim=18, k_rpe=True, q_rpe=True, v_rpe=True,
            use_diameter_parent=True, version_holder=ref.VersionHolder('3.0.0'), **kw)
        st.apply(ref.init_weights)
        xc = x_child.clone()
        ea = d2.edge_attr.clone()
        outs, probe, grads, pgrads = _run(
            st, (x_parent, xc, ni2, d1.super_index),
            dict(pos=d2.pos, node_size=d2.node_size, super_index=None,
                 edge_index=d2.edge_index, edge_attr=ea, num_super=d2.num_nodes),
            [xc, ea], torch.Generator().manual_seed(41))
        cases[name] = dict(
            cfg=dict(dim=C, num_heads=4, qk_dim=4, pool=kw['pool'],
                     pre_norm=kw.get('pre_norm', True)),
            sd={k: v.detach().clone() for k, v in st.state_dict().items()},
            x_parent=x_parent, x_child=x_child, norm_index=ni2, pool_index=d1.super_index,
            pos=d2.pos, node_size=d2.node_size, edge_index=d2.edge_index,
            edge_attr=d2.edge_attr, num_super=d2.num_nodes, out=outs[0], diam=outs[1],
            probe=probe, dx_child=grads[0], dedge_attr=grads[1], dparams=pgrads)
    # up stage
    torch.manual_seed(32)
    st = ref.UpNFuseStage(
        C, num_blocks=1, num_heads=4, in_mlp=[3 + 1 + C + C, C, C], mlp_norm=ref.GraphNorm,
        norm=ref.GraphNorm, qk_dim=4, in_rpe_dim=18, k_rpe=True, q_rpe=True, v_rpe=True,
        use_diameter_parent=True, no_ffn=True, version_holder=ref.VersionHolder('3.0.0'))
    st.apply(ref.init_weights)
    xp = torch.randn(d2.num_nodes, C, generator=gen)
    xc = x_child.clone()
    xpp = xp.clone()
    outs, probe, grads, pgrads = _run(
        st, (xc, xpp, ni1, d1.super_index),
        dict(pos=d1.pos, node_size=d1.node_size, super_index=d1.super_index,
             edge_index=d1.edge_index, edge_attr=d1.edge_attr),
        [xc, xpp], torch.Generator().manual_seed(42))
    cases['up'] = dict(
        cfg=dict(dim=C, num_heads=4, qk_dim=4),
        sd={k: v.detach().clone() for k, v in st.state_dict().items()},
        x_child=x_child, x_parent=xp, norm_index=ni1, unpool_index=d1.super_index, pos=d1.pos,
        node_size=d1.node_size, super_index=d1.super_index, edge_index=d1.edge_index,
        edge_attr=d1.edge_attr, out=outs[0], probe=probe, dx_child=grads[0],
        dx_parent=grads[1], dparams=pgrads)
    return cases


def edge_feature_case(ref):
    from superpoint_transformer_b200.synthetic import make_nag
    nag = make_nag([300, 50], mean_degree=8, seed=77)
    d = nag[1]
    # degenerate rows: zero mean offset (0/0 -> NaN -> 0) and coincident centroids
    d.edge_attr[0, :3] = 0
    j = int(d.edge_index[1, 1])
    d.pos[j] = d.pos[int(d.edge_index[0, 1])]
    raw = dict(edge_index=d.edge_index.clone(), edge_attr=d.edge_attr.clone(), pos=d.pos.clone(),
               normal=d.normal.clone(), log_length=d.log_length.clone(),
               log_surface=d.log_surface.clone(), log_volume=d.log_volume.clone(),
               log_size=d.log_size.clone(), num_nodes=d.num_nodes)
    out = ref.on_the_fly_horizontal_edge_features(d.clone())
    ei2, ea2 = out.edge_index, out.edge_attr
    one = type(nag)([out], 1)
    one = ref.NAGAddSelfLoops()(one)
    raw.update(sym_edge_index=ei2, sym_edge_attr=ea2, loop_edge_index=one[1].edge_index,
               loop_edge_attr=one[1].edge_attr)
    # vertical (child -> parent) features through the reference function
    nag2 = make_nag([300, 50], mean_degree=8, seed=78)
    child, parent = nag2[1], nag2[2]
    k = int(child.super_index[0])
    parent.pos[k] = child.pos[0]          # coincident centroids: 0/0 -> NaN -> 0
    vkeys = ('log_length', 'log_surface', 'log_volume', 'log_size')
    raw['v'] = dict(child_pos=child.pos.clone(), parent_pos=parent.pos.clone(),
                    child_normal=child.normal.clone(), parent_normal=parent.normal.clone(),
                    child_logs=[child[kk].clone() for kk in vkeys],
                    parent_logs=[parent[kk].clone() for kk in vkeys],
                    super_index=child.super_index.clone())
    out_child = ref.on_the_fly_vertical_edge_features(child.clone(), parent.clone())
    raw['v']['v_edge_attr'] = out_child.v_edge_attr
    return raw


def round2_cases(ref):
    """fixtures added in round 2: attentive pools (src/nn/pool.py:156-360), superedge features
    from sub-edges (src/transforms/graph.py:950-1060), key subsets of the on-the-fly
    horizontal features (:1063-1277)."""
    from superpoint_transformer_b200.data import Data
    from superpoint_transformer_b200.synthetic import make_nag
    cases = {}
    # ---- attentive pools -------------------------------------------------------------
    gen = torch.Generator().manual_seed(29)
    Nc, Np, C, Cp, H, D, F = 700, 90, 32, 20, 4, 8, 9
    idx = torch.randint(0, Np - 2, (Nc,), generator=gen)       # 2 parents without children
    xc, xp = torch.randn(Nc, C, generator=gen), torch.randn(Np, Cp, generator=gen)
    ea = torch.randn(Nc, F, generator=gen)
    specs = {
        'attpool_kq': (ref.AttentivePool, dict(dim=C, q_in_dim=Cp, num_heads=H, qk_dim=D,
                                               in_rpe_dim=F, k_rpe=True, q_rpe=True, out_dim=C)),
        'attpool_k_shared': (ref.AttentivePool, dict(dim=C, q_in_dim=Cp, num_heads=H, qk_dim=D,
                                                     in_rpe_dim=F, k_rpe=True,
                                                     heads_share_rpe=True, qk_scale='d+g')),
        'attpool_plain_inproj': (ref.AttentivePool, dict(dim=C, q_in_dim=Cp, num_heads=H,
                                                         qk_dim=D, in_dim=C, out_dim=24)),
        'attpool_learnt_kq': (ref.AttentivePoolWithLearntQueries,
                              dict(dim=C, num_heads=H, qk_dim=D, in_rpe_dim=F, k_rpe=True,
                                   q_rpe=True, out_dim=C)),
        'attpool_learnt_plain': (ref.AttentivePoolWithLearntQueries,
                                 dict(dim=C, num_heads=H, qk_dim=D)),
    }
    for name, (cls, kw) in specs.items():
        torch.manual_seed(77)
        m = cls(**kw)
        m.apply(ref.init_weights)
        for p in m.parameters():
            if p.dim() == 1:
                p.data.normal_(0, 0.2, generator=gen)
        a, b, e = xc.clone(), xp.clone(), ea.clone()
        use_ea = kw.get('k_rpe') or kw.get('q_rpe')
        wrt = [a, b, e] if use_ea else [a, b]
        outs, probe, grads, pgrads = _run(m, (a, b, idx), dict(edge_attr=e if use_ea else None,
                                                                num_pool=Np), wrt,
                                          torch.Generator().manual_seed(31))
        m64 = cls(**kw).double()
        m64.load_state_dict({k: v.double() for k, v in m.state_dict().items()})
        out64 = m64(xc.double(), xp.double(), idx, edge_attr=ea.double() if use_ea else None,
                    num_pool=Np).detach()
        cases[name] = dict(cls=cls.__name__, kw=kw, x_child=xc, x_parent=xp, index=idx,
                           edge_attr=ea if use_ea else None, num_pool=Np,
                           sd={k: v.detach().clone() for k, v in m.state_dict().items()},
                           out=outs[0], out64=out64, probe=probe, dx_child=grads[0],
                           dx_parent=grads[1], dedge_attr=grads[2] if use_ea else None,
                           dparams=pgrads)
    # ---- superedge features from sub-edges -------------------------------------------------
    gen = torch.Generator().manual_seed(41)
    N0, N1, E = 4000, 120, 260
    points = torch.rand(N0, 3, generator=gen) * 20
    i = torch.randint(0, N1, (4 * E,), generator=gen)
    j = torch.randint(0, N1, (4 * E,), generator=gen)
    lo, hi = torch.minimum(i, j), torch.maximum(i, j)
    uid = torch.unique(lo[lo != hi] * N1 + hi[lo != hi])[:E]
    se = torch.stack((uid // N1, uid % N1))
    E = se.shape[1]
    counts = torch.randint(1, 40, (E,), generator=gen)
    counts[0] = 1                          # single sub-edge: std = 0 / (1 + 1e-6)
    counts[1] = 2                          # two opposite sub-edges: zero mean offset
    counts[2] = 3                          # mean offset along (v, v, v): second fallback
    se_id = torch.repeat_interleave(torch.arange(E), counts)
    Es = se_id.numel()
    spi = torch.randint(0, N0, (2, Es), generator=gen)
    first = torch.cumsum(counts, 0) - counts
    a0, b0 = int(first[1]), int(first[1]) + 1
    spi[:, b0] = spi[:, a0].flip(0)        # reversed copy -> offsets cancel exactly
    c0 = int(first[2])
    points[spi[0, c0:c0 + 3]] = torch.tensor([[1., 1., 1.]])
    points[spi[1, c0:c0 + 3]] = torch.tensor([[3., 3., 3.]])
    perm = torch.randperm(Es, generator=gen)   # se_id unsorted, as subedges() may emit
    se_id, spi = se_id[perm], spi[:, perm]
    d = Data(edge_index=se.clone(), pos=torch.zeros(N1, 3))
    out = ref.minimalistic_horizontal_edge_features(d, points.clone(), spi, se_id)
    cases['superedge'] = dict(points=points, se_point_index=spi, se_id=se_id, edge_index=se,
                              edge_attr=out.edge_attr.clone())
    # ---- key subsets of the on-the-fly horizontal features -----------------------------------
    nag = make_nag([200, 30], mean_degree=8, seed=91)
    dd = nag[1]
    raw = dict(edge_index=dd.edge_index.clone(), edge_attr=dd.edge_attr.clone(), pos=dd.pos.clone(),
               normal=dd.normal.clone(), log_length=dd.log_length.clone(),
               log_surface=dd.log_surface.clone(), log_volume=dd.log_volume.clone(),
               log_size=dd.log_size.clone(), num_nodes=dd.num_nodes, subsets={})
    for keys in (['mean_off', 'std_off', 'log_size', 'centroid_dir'],
                 ['normal_angle', 'centroid_dist'],
                 ['angle_source', 'mean_dist', 'log_length', 'log_volume'],
                 []):
        out = ref.on_the_fly_horizontal_edge_features(dd.clone(), keys=keys or None) \
            if keys else None
        raw['subsets']['+'.join(keys) if keys else 'none'] = dict(
            keys=keys, edge_index=None if out is None else out.edge_index,
            edge_attr=None if out is None else out.edge_attr)
    del raw['subsets']['none']   # sanitize_keys(None) = default set; the empty set is a host-side branch
    cases['h_subsets'] = raw
    # ---- node-difference RPE (src/nn/attention.py:259-291) and attention dropout (:310-311) --------
    N, C, H, D, F = 70, 32, 4, 4, 12
    specs = {
        'delta_kq': dict(k_delta_rpe=True, q_delta_rpe=True),
        'delta_kq_minus_edge': dict(k_delta_rpe=True, q_delta_rpe=True, q_on_minus_rpe=True,
                                    k_rpe=True, q_rpe=True, v_rpe=True),
        'delta_k_share_minus': dict(k_delta_rpe=True, qk_share_rpe=True, q_on_minus_rpe=True,
                                    k_rpe=True),
        'delta_heads_share': dict(k_delta_rpe=True, q_delta_rpe=True, heads_share_rpe=True,
                                  qk_scale='d+g'),
        'attn_drop': dict(k_rpe=True, q_rpe=True, v_rpe=True, attn_drop=0.3),
    }
    for name, kw in specs.items():
        gen = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 100000)
        ei, _ = _graph(gen, N, N * 4)
        E = ei.shape[1]
        x = torch.randn(N, C, generator=gen)
        ea = torch.randn(E, F, generator=gen)
        torch.manual_seed(4321)
        blk = ref.SelfAttentionBlock(C, num_heads=H, qk_dim=D, in_rpe_dim=F, out_dim=C, **kw)
        blk.apply(ref.init_weights)
        for p in blk.parameters():
            if p.dim() == 1:
                p.data.normal_(0, 0.1, generator=gen)
        mask = None
        if 'attn_drop' in kw:
            # the dropout of the attention weights is the first consumer of the global RNG in
            # forward(): re-seeding reproduces the multipliers it used
            blk.train()
            torch.manual_seed(999)
            mask = torch.nn.functional.dropout(torch.ones(E, H), kw['attn_drop'], True)
            torch.manual_seed(999)
        else:
            blk.eval()
        xin, ein = x.clone(), ea.clone()
        outs, probe, grads, pgrads = _run(blk, (xin, ei, ein), {}, [xin, ein],
                                          torch.Generator().manual_seed(17))
        cases[name] = dict(cfg=dict(dim=C, num_heads=H, qk_dim=D, in_rpe_dim=F, **kw), x=x,
                           edge_index=ei, edge_attr=ea, mask=mask,
                           sd={k: v.detach().clone() for k, v in blk.state_dict().items()},
                           out=outs[0], probe=probe, dx=grads[0], dedge_attr=grads[1],
                           dparams=pgrads)
    return cases


def main():
    if not R.available():
        print('reference sources not available; cannot regenerate golden vectors')
        return 1
    ref = R.load()
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == 'round2':   # leave the round-1 fixtures untouched
        torch.save(round2_cases(ref), os.path.join(OUT, 'round2.pt'))
        print('round2.pt', os.path.getsize(os.path.join(OUT, 'round2.pt')))
        return 0
    torch.save(attention_cases(ref), os.path.join(OUT, 'attention.pt'))
    torch.save(segment_cases(ref), os.path.join(OUT, 'segment.pt'))
    torch.save(stage_cases(ref), os.path.join(OUT, 'stage.pt'))
    torch.save(spt_case(ref), os.path.join(OUT, 'spt_nano3.pt'))
    torch.save(edge_feature_case(ref), os.path.join(OUT, 'edge_features.pt'))
    torch.save(norm_pool_cases(ref), os.path.join(OUT, 'norms.pt'))
    torch.save(round2_cases(ref), os.path.join(OUT, 'round2.pt'))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    return 0


if __name__ == '__main__':
    sys.exit(main())
