"""CPU: the oracle restatement (oracle/path.py) against golden vectors produced by
the reference's own source files (oracle/make_golden.py).  fp32, tight tolerance:
both sides run the same torch CPU kernels, so they agree to rounding."""
import pytest
import torch

from oracle import path as P
from helpers import nag_from_golden, assert_close, grad_close

TOL = dict(atol=2e-6, rtol=2e-6)


def _req(*ts):
    return [t.clone().requires_grad_(True) for t in ts]


def _sd_grad(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v)
            for k, v in sd.items()}


def _attn_kw(cfg):
    return dict(num_heads=cfg['num_heads'], qk_dim=cfg['qk_dim'],
                qk_scale_mode=cfg.get('qk_scale'),
                heads_share_rpe=cfg.get('heads_share_rpe', False),
                qk_share_rpe=cfg.get('qk_share_rpe', False),
                q_on_minus_rpe=cfg.get('q_on_minus_rpe', False))


def test_attention_cases(golden):
    cases = golden('attention.pt')
    assert len(cases) >= 10
    for name, c in cases.items():
        sd = _sd_grad({'sa.' + k: v for k, v in c['sd'].items()})
        x, ea = _req(c['x'], c['edge_attr'])
        out = P.self_attention(sd, 'sa', x, c['edge_index'], ea, **_attn_kw(c['cfg']))
        torch.testing.assert_close(out, c['out'], **TOL, msg=name)
        (out * c['probe']).sum().backward()
        torch.testing.assert_close(x.grad, c['dx'], atol=1e-5, rtol=1e-5, msg=name)
        torch.testing.assert_close(ea.grad, c['dedge_attr'], atol=1e-5, rtol=1e-5, msg=name)
        for k, g in c['dparams'].items():
            torch.testing.assert_close(sd['sa.' + k].grad, g, atol=2e-5, rtol=1e-4,
                                       msg=f'{name}:{k}')
        # fp64 "truth": the fp32 reference output is within fp32 rounding of it
        out64 = P.self_attention({k: v.double() for k, v in sd.items()}, 'sa',
                                 c['x'].double(), c['edge_index'], c['edge_attr'].double(),
                                 **_attn_kw(c['cfg']))
        torch.testing.assert_close(out64, c['out64'], atol=1e-6, rtol=1e-6, msg=name)


def test_pool_unpool_cases(golden):
    cases = golden('segment.pt')
    for red in ('max', 'min', 'mean', 'sum'):
        c = cases[f'pool_{red}']
        (x,) = _req(c['x'])
        out = P.pool(x, c['index'], c['num_pool'], red)
        torch.testing.assert_close(out, c['out'], **TOL)
        (out * c['probe']).sum().backward()
        torch.testing.assert_close(x.grad, c['dx'], **TOL)
    c = cases['unpool']
    (x,) = _req(c['x'])
    out = P.unpool(x, c['index'])
    torch.testing.assert_close(out, c['out'], atol=0, rtol=0)
    (out * c['probe']).sum().backward()
    torch.testing.assert_close(x.grad, c['dx'], atol=1e-5, rtol=1e-5)


def test_unit_sphere_norm(golden):
    c = golden('segment.pt')['unit_sphere']
    p, d = P.unit_sphere_norm(c['pos'], c['index'], c['w'], c['num_super'])
    torch.testing.assert_close(p, c['pos_w'], **TOL)
    torch.testing.assert_close(d, c['diam_w'], **TOL)
    p, d = P.unit_sphere_norm(c['pos'], c['index'], None, c['num_super'])
    torch.testing.assert_close(p, c['pos_nw'], **TOL)
    torch.testing.assert_close(d, c['diam_nw'], **TOL)
    p, d = P.unit_sphere_norm(c['pos'], None, c['w'])
    torch.testing.assert_close(p, c['pos_none'], **TOL)
    torch.testing.assert_close(d, c['diam_none'], **TOL)


@pytest.mark.parametrize('name', ['mlp_graphnorm', 'mlp_graphnorm_unsorted'])
def test_mlp_graphnorm(golden, name):
    c = golden('segment.pt')[name]
    sd = _sd_grad({'m.' + k: v for k, v in c['sd'].items()})
    (x,) = _req(c['x'])
    out = P.mlp(sd, 'm', x, c['batch'])
    torch.testing.assert_close(out, c['out'], atol=1e-5, rtol=1e-5)
    (out * c['probe']).sum().backward()
    torch.testing.assert_close(x.grad, c['dx'], atol=1e-4, rtol=1e-4)
    for k, g in c['dparams'].items():
        torch.testing.assert_close(sd['m.' + k].grad, g, atol=1e-3, rtol=1e-4, msg=k)


def test_stage_cases(golden):
    cases = golden('stage.pt')
    for name in ('down_mean_ffn', 'down_max_postnorm'):
        c = cases[name]
        cfg = c['cfg']
        sd = {'st.' + k: v for k, v in c['sd'].items()}
        out, diam = P.down_stage(
            sd, 'st', c['x_parent'], c['x_child'], c['norm_index'], c['pool_index'],
            c['num_super'], cfg['pool'], pos=c['pos'], node_size=c['node_size'],
            super_index=None, edge_index=c['edge_index'], edge_attr=c['edge_attr'],
            use_diameter_parent=True,
            block_kw=dict(num_heads=cfg['num_heads'], qk_dim=cfg['qk_dim'],
                          pre_norm=cfg['pre_norm']))
        torch.testing.assert_close(out, c['out'], atol=2e-5, rtol=2e-5, msg=name)
        torch.testing.assert_close(diam, c['diam'], **TOL)
    c = cases['up']
    sd = {'st.' + k: v for k, v in c['sd'].items()}
    out, _ = P.up_stage(sd, 'st', c['x_child'], c['x_parent'], c['norm_index'],
                        c['unpool_index'], pos=c['pos'], node_size=c['node_size'],
                        super_index=c['super_index'], edge_index=c['edge_index'],
                        edge_attr=c['edge_attr'], use_diameter_parent=True,
                        block_kw=dict(num_heads=4, qk_dim=4))
    torch.testing.assert_close(out, c['out'], atol=2e-5, rtol=2e-5)


def test_spt_nano3(golden):
    c = golden('spt_nano3.pt')
    nag = nag_from_golden(c['levels'], c['start_i_level'])
    sd = _sd_grad(c['sd'])
    out = P.spt_forward(sd, nag, num_heads=4, qk_dim=4, nano=True, num_down=2, num_up=2,
                        use_diameter_parent=True, pool_reduce='max')
    torch.testing.assert_close(out, c['out'], atol=5e-5, rtol=5e-5)
    (out * c['probe']).sum().backward()
    worst = 0.0
    for k, g in c['dparams'].items():
        got = sd[k].grad
        assert got is not None, k
        scale = g.abs().max().item() + 1e-6
        worst = max(worst, (got - g).abs().max().item() / scale)
    assert worst < 5e-3, worst
    out64 = P.spt_forward({k: (v.double() if v.is_floating_point() else v)
                           for k, v in c['sd'].items()},
                          _to64(nag), num_heads=4, qk_dim=4, nano=True, num_down=2, num_up=2,
                          use_diameter_parent=True, pool_reduce='max')
    torch.testing.assert_close(out64, c['out64'], atol=1e-6, rtol=1e-6)


def _to64(nag):
    out = nag.clone()
    for d in out:
        for k in d.keys:
            v = d[k]
            if torch.is_tensor(v) and v.is_floating_point():
                d[k] = v.double()
    return out


def test_edge_features(golden):
    c = golden('edge_features.pt')
    ei, ea = P.horizontal_edge_features(
        c['edge_index'], c['edge_attr'], c['pos'], c['normal'], c['log_length'],
        c['log_surface'], c['log_volume'], c['log_size'])
    assert torch.equal(ei, c['sym_edge_index'])
    torch.testing.assert_close(ea, c['sym_edge_attr'], atol=0, rtol=0)
    ei2, ea2 = P.add_self_loops(ei, ea, c['num_nodes'])
    assert torch.equal(ei2, c['loop_edge_index'])
    torch.testing.assert_close(ea2, c['loop_edge_attr'], atol=0, rtol=0)
    assert not torch.isnan(ea).any()


def test_vertical_edge_features(golden):
    v = golden('edge_features.pt')['v']
    out = P.vertical_edge_features(v['child_pos'], v['parent_pos'], v['child_normal'],
                                   v['parent_normal'], v['child_logs'], v['parent_logs'],
                                   v['super_index'])
    torch.testing.assert_close(out, v['v_edge_attr'], atol=0, rtol=0)
    assert out.shape[1] == 9 and not torch.isnan(out).any()


@pytest.mark.parametrize('name,groups', [('groupnorm_g4', 4), ('groupnorm_g1', 1),
                                         ('groupnorm_g8_nobatch', 8)])
def test_group_norm(golden, name, groups):
    """oracle restatement of GroupNorm(mode='graph') vs the reference source run here"""
    c = golden('norms.pt')[name]
    x, w, b = _req(c['x'], c['weight'], c['bias'])
    out = P.group_norm(x, c['batch'], w, b, groups)
    torch.testing.assert_close(out, c['out'], atol=1e-5, rtol=1e-5)
    (out * c['probe']).sum().backward()
    torch.testing.assert_close(x.grad, c['dx'], atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(w.grad, c['dparams']['weight'], atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(b.grad, c['dparams']['bias'], atol=1e-4, rtol=1e-4)


def test_layer_norm_graph_is_group_norm_1(golden):
    """PyG LayerNorm(mode='graph') with a batch vector == GroupNorm(num_groups=1)"""
    c = golden('norms.pt')['layernorm_graph']
    out = P.group_norm(c['x'], c['batch'], c['weight'], c['bias'], 1)
    torch.testing.assert_close(out, c['out'], atol=1e-5, rtol=1e-5)


def test_std_pool(golden):
    c = golden('norms.pt')['pool_std']
    (x,) = _req(c['x'])
    out = P.std_pool(x, c['index'], c['num_pool'])
    torch.testing.assert_close(out, c['out'], atol=1e-6, rtol=1e-5)
    (out * c['probe']).sum().backward()
    torch.testing.assert_close(x.grad, c['dx'], atol=1e-6, rtol=1e-4)


@pytest.mark.parametrize('name,kind', [('block_layernorm_graph', 'layer_graph'),
                                       ('block_groupnorm4', ('group', 4))])
def test_transformer_block_graphwise_norms(golden, name, kind):
    """TransformerBlock with the reference's code-default norm (PyG LayerNorm, mode='graph')
    and with GroupNorm, 3-graph batch: oracle vs the reference source run here."""
    c = golden('norms.pt')[name]
    sd = _sd_grad({'b.' + k: v for k, v in c['sd'].items()})
    x, ea = _req(c['x'], c['edge_attr'])
    P.NORM_KIND['kind'] = kind
    try:
        out = P.transformer_block(sd, 'b', x, c['batch'], c['edge_index'], ea, num_heads=4, qk_dim=4)
    finally:
        P.NORM_KIND['kind'] = None
    torch.testing.assert_close(out, c['out'], atol=2e-5, rtol=1e-5)
    (out * c['probe']).sum().backward()
    torch.testing.assert_close(x.grad, c['dx'], atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(ea.grad, c['dedge_attr'], atol=1e-4, rtol=1e-4)


# --------------------------------------------------------------------------- #
#  round-2 fixtures (tests/golden/round2.pt)
# --------------------------------------------------------------------------- #
ATTPOOL = ['attpool_kq', 'attpool_k_shared', 'attpool_plain_inproj', 'attpool_learnt_kq',
           'attpool_learnt_plain']


@pytest.mark.parametrize('name', ATTPOOL)
def test_attentive_pools(golden, name):
    """oracle.path.attentive_pool == the reference AttentivePool / AttentivePoolWithLearntQueries
    (src/nn/pool.py:156-360) on the same inputs, outputs and every gradient."""
    c = golden('round2.pt')[name]
    kw = c['kw']
    sd = {'p.' + k: v.clone().requires_grad_(True) for k, v in c['sd'].items()}
    xc, xp = c['x_child'].clone().requires_grad_(True), c['x_parent'].clone().requires_grad_(True)
    ea = None if c['edge_attr'] is None else c['edge_attr'].clone().requires_grad_(True)
    out = P.attentive_pool(sd, 'p', xc, xp, c['index'], ea, c['num_pool'],
                           num_heads=kw['num_heads'], qk_dim=kw['qk_dim'],
                           qk_scale_mode=kw.get('qk_scale'),
                           heads_share_rpe=kw.get('heads_share_rpe', False),
                           learnt_queries=c['cls'] == 'AttentivePoolWithLearntQueries')
    assert_close(out, c['out'], atol=1e-5, rtol=1e-5, what=name)
    (out * c['probe']).sum().backward()
    grad_close(xc.grad, c['dx_child'], what=name + ' dx_child')
    if c['cls'] == 'AttentivePool':
        grad_close(xp.grad, c['dx_parent'], what=name + ' dx_parent')
    if ea is not None:
        grad_close(ea.grad, c['dedge_attr'], what=name + ' dedge_attr')
    for k, g in c['dparams'].items():
        grad_close(sd['p.' + k].grad, g, what=f'{name} d{k}')


def test_superedge_features(golden):
    c = golden('round2.pt')['superedge']
    out = P.minimalistic_horizontal_edge_features(c['points'], c['se_point_index'], c['se_id'],
                                                  c['edge_index'].shape[1])
    assert_close(out, c['edge_attr'], atol=1e-6, rtol=1e-6, what='superedge features')


def test_horizontal_feature_key_subsets(golden):
    c = golden('round2.pt')['h_subsets']
    for name, sub in c['subsets'].items():
        ei, ea = P.horizontal_edge_features(c['edge_index'], c['edge_attr'], c['pos'], c['normal'],
                                            c['log_length'], c['log_surface'], c['log_volume'],
                                            c['log_size'], keys=sub['keys'])
        assert torch.equal(ei, sub['edge_index'])
        assert_close(ea, sub['edge_attr'], atol=1e-6, rtol=1e-6, what=f'subset {name}')


DELTA = ['delta_kq', 'delta_kq_minus_edge', 'delta_k_share_minus', 'delta_heads_share',
         'attn_drop']


@pytest.mark.parametrize('name', DELTA)
def test_delta_rpe_and_attention_dropout(golden, name):
    """node-difference RPE (src/nn/attention.py:259-291) and attention dropout (:310-311):
    oracle == reference block, outputs and every gradient."""
    c = golden('round2.pt')[name]
    cfg = c['cfg']
    sd = {'sa.' + k: v.clone().requires_grad_(True) for k, v in c['sd'].items()}
    x, ea = _req(c['x'], c['edge_attr'])
    out = P.self_attention(sd, 'sa', x, c['edge_index'], ea, attn_drop_mask=c['mask'],
                           **_attn_kw(cfg))
    assert_close(out, c['out'], atol=2e-5, rtol=2e-5, what=name)
    (out * c['probe']).sum().backward()
    grad_close(x.grad, c['dx'], what=name + ' dx')
    grad_close(ea.grad, c['dedge_attr'], what=name + ' dedge_attr')
    for k, g in c['dparams'].items():
        grad_close(sd['sa.' + k].grad, g, what=f'{name} d{k}')
