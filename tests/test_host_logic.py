"""CPU: host-side logic — data containers, batching offsets, synthetic NAG
invariants, qk-scale parsing, drop-in state-dict layout."""
import numpy as np
import pytest
import torch

import superpoint_transformer_b200 as S
from superpoint_transformer_b200.synthetic import make_nag
from superpoint_transformer_b200.utils.nn import build_qk_scale, listify_with_reference
from helpers import nag_from_golden


def test_synthetic_nag_invariants():
    nag = make_nag([2000, 300, 40], mean_degree=10, seed=4)
    assert nag.start_i_level == 1 and nag.absolute_num_levels == 4
    for l in nag.level_range:
        d = nag[l]
        n = d.num_nodes
        ei = d.edge_index
        assert ei.dtype == torch.int64 and ei.min() >= 0 and ei.max() < n
        assert (ei[0] < ei[1]).all()                      # trimmed: i<j, no self-loops
        uid = ei[0] * n + ei[1]
        assert uid.unique().numel() == uid.numel()        # coalesced
        assert d.edge_attr.shape == (ei.shape[1], 7)
        if l < nag.end_i_level:
            sup = d.super_index
            assert sup.unique().numel() == nag[l + 1].num_nodes   # dense: every parent used
            cl = nag[l + 1].sub
            assert cl.pointers[0] == 0 and cl.pointers[-1] == n
            assert torch.equal(torch.sort(cl.points).values, torch.arange(n))
            assert torch.equal(cl.to_super_index(), sup)
            # ascending child ids inside every cluster (stable grouping)
            same = sup[cl.points][1:] == sup[cl.points][:-1]
            assert (cl.points[1:][same] > cl.points[:-1][same]).all()


def test_nag_batch_offsets():
    a = make_nag([50, 10], 4, seed=1)
    b = make_nag([70, 12], 4, seed=2)
    batch = S.NAGBatch.from_nag_list([a, b])
    assert batch[1].num_nodes == 120 and batch[2].num_nodes == 22
    assert torch.equal(batch[1].batch, torch.cat((torch.zeros(50), torch.ones(70))).long())
    # no edge crosses batch items, super_index / sub stay consistent
    ei = batch[1].edge_index
    assert ((ei[0] < 50) == (ei[1] < 50)).all()
    assert torch.equal(batch[2].sub.to_super_index(), batch[1].super_index)
    assert (batch[1].super_index[50:] >= 10).all() and batch[1].super_index.max() == 21
    assert torch.equal(batch[1].norm_index('graph'), batch[1].batch)


def test_node_size_and_super_index_chain_cpu():
    nag = make_nag([500, 60, 7], 6, seed=3)
    nag = S.transforms.NodeSize()(nag)
    total = nag[1].node_size.sum()
    assert nag[2].node_size.sum() == total and nag[3].node_size.sum() == total
    sup13 = nag.get_super_index(3, low=1)
    assert torch.equal(sup13, nag[2].super_index[nag[1].super_index])


def test_qk_scale_parsing():
    assert build_qk_scale(128, 4, None) == (S.ops.SCALE_D_TIMES_G, 32 ** -0.5)
    assert build_qk_scale(64, 16, 'd + g')[0] == S.ops.SCALE_D_PLUS_G
    assert build_qk_scale(64, 16, 'g.d')[0] == S.ops.SCALE_D_TIMES_G
    assert build_qk_scale(64, 16, 'd') == (S.ops.SCALE_D, 0.5)
    assert build_qk_scale(64, 16, 'G')[0] == S.ops.SCALE_G
    assert build_qk_scale(64, 16, 0.3) == (S.ops.SCALE_CONST, 0.3)
    with pytest.raises(ValueError):
        build_qk_scale(64, 16, 'x')


def test_listify_with_reference():
    ref, a, b = listify_with_reference([64, 64, 64], 3, [1, 2, 3])
    assert ref == [64, 64, 64] and a == [3, 3, 3] and b == [1, 2, 3]
    ref, a = listify_with_reference(None, 3)
    assert ref == [] and a == []
    ref, a = listify_with_reference(64, 'max')
    assert ref == [64] and a == ['max']


def test_spt_state_dict_matches_reference_layout(golden):
    """drop-in: same parameter names and shapes as the reference SPT instantiated
    with the same kwargs (golden state_dict comes from the reference class)"""
    c = golden('spt_nano3.pt')
    net = S.SPT(mlp_norm=S.nn.GraphNorm, norm=S.nn.GraphNorm, **c['cfg'])
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    theirs = {k: tuple(v.shape) for k, v in c['sd'].items()}
    assert mine == theirs
    net.load_state_dict(c['sd'], strict=True)
    low_lr = [k for k in mine if 'transformer_blocks' in k or 'down_pool_block' in k]
    assert len(low_lr) > 0   # differential-LR group selector of semantic.py:1253 still works


def test_attention_block_state_dict_matches_reference_layout(golden):
    for name, c in golden('attention.pt').items():
        cfg = dict(c['cfg'])
        dim = cfg.pop('dim')
        blk = S.nn.SelfAttentionBlock(dim, out_dim=dim, **cfg)
        mine = {k: tuple(v.shape) for k, v in blk.state_dict().items()}
        theirs = {k: tuple(v.shape) for k, v in c['sd'].items()}
        assert mine == theirs, name


def test_shard_indices_partition_and_balance():
    from superpoint_transformer_b200.distributed import shard_indices
    W = 8
    parts = [shard_indices(64, r, W) for r in range(W)]
    assert sorted(sum(parts, [])) == list(range(64)) and all(len(p) == 8 for p in parts)
    w = np.random.default_rng(0).integers(1, 100, size=37)
    parts = [shard_indices(37, r, W, weights=w) for r in range(W)]
    assert sorted(sum(parts, [])) == list(range(37))
    loads = [sum(w[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= w.max()
