"""Shared test helpers: rebuild NAG objects from golden fixtures, tolerances."""
import torch

from superpoint_transformer_b200.data import Data, NAG, Cluster

# parity bar of BASELINE.json north_star: 1e-4 on fp32 outputs
ATOL = 1e-4
RTOL = 1e-4


def assert_close(a, b, atol=ATOL, rtol=RTOL, what=''):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), (f'{what}: max abs err {err.max().item():.3e} '
                           f'(ref max {b.abs().max().item():.3e}), {int(bad.sum())} / {bad.numel()} off')


def grad_close(a, b, what='', rel=2e-4, floor=2e-5):
    """gradients are compared relative to the tensor's scale (sums over many edges);
    `floor` absorbs gradients that are analytically ~0 (e.g. a bias in front of a
    mean-subtracting norm).  A None reference gradient means 'input unused'."""
    if b is None:
        assert a is None or float(a.abs().max()) == 0.0, f'{what}: expected no gradient'
        return
    assert a is not None, f'{what}: gradient missing'
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    scale = max(b.abs().max().item(), 1e-6)
    err = (a - b).abs().max().item()
    assert err <= rel * scale + floor, f'{what}: max abs err {err:.3e} vs scale {scale:.3e}'


def assert_within_fp32_noise(mine, ref32, truth64, what='', factor=2.0, floor=1e-4):
    """Deep stacks: the reference's own fp32 result sits |ref32 - truth64| away from the
    fp64 truth (1.9e-3 max for the golden SPT).  The CUDA path must be as close to the
    truth as the reference is (x `factor`), in max- and mean-norm."""
    mine = mine.detach().cpu().double()
    e_ref = (ref32.double() - truth64).abs()
    e_mine = (mine - truth64).abs()
    assert e_mine.max() <= factor * e_ref.max() + floor, \
        f'{what}: max err vs fp64 truth {e_mine.max():.3e} > {factor} x reference fp32 {e_ref.max():.3e}'
    assert e_mine.mean() <= factor * e_ref.mean() + floor / 10, \
        f'{what}: mean err vs fp64 truth {e_mine.mean():.3e} vs reference fp32 {e_ref.mean():.3e}'


def nag_from_golden(levels, start_i_level, raw=False, device='cpu'):
    datas = []
    for l in sorted(levels):
        d = dict(levels[l])
        sub_ptr, sub_pts = d.pop('sub_pointers', None), d.pop('sub_points', None)
        rei, rea = d.pop('raw_edge_index'), d.pop('raw_edge_attr')
        d = {k: v for k, v in d.items() if not k.startswith('_')}
        if raw:
            d['edge_index'], d['edge_attr'] = rei, rea
        data = Data(**d)
        if sub_ptr is not None:
            data.sub = Cluster(sub_ptr, sub_pts)
        datas.append(data)
    nag = NAG(datas, start_i_level)
    return nag.to(device) if device != 'cpu' else nag
