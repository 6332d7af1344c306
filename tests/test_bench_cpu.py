"""CPU checks of bench.py's host logic: the cfg-5 tile cutter, the sharding description, the
reference arm's JSON line (tiny scene) and the gradient accumulation of FlatGradients."""
import json
import subprocess
import sys
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from superpoint_transformer_b200.synthetic import make_nag  # noqa: E402
from superpoint_transformer_b200.distributed import FlatGradients  # noqa: E402


def test_cut_tiles_partitions_the_nag_and_drops_only_cross_tile_edges():
    full = make_nag([6000, 1200, 240], seed=3, spatial=True)
    tiles, kept = bench.cut_tiles(full, 4)
    assert 0.5 < kept <= 1.0
    for l in full.level_range:
        assert sum(t[l].num_nodes for t in tiles) == full[l].num_nodes
    n_edges = 0
    for t in tiles:
        for l in t.level_range:
            d = t[l]
            assert d.edge_index.numel() == 0 or (0 <= int(d.edge_index.min())
                                                 and int(d.edge_index.max()) < d.num_nodes)
            assert d.edge_attr.shape[0] == d.edge_index.shape[1]
            if d.super_index is not None:
                # every parent of the tile keeps at least one child, ids dense
                assert torch.equal(torch.unique(d.super_index), torch.arange(t[l + 1].num_nodes))
                # the rebuilt Cluster is consistent with super_index
                assert torch.equal(t[l + 1].sub.to_super_index(), d.super_index)
        n_edges += t[1].edge_index.shape[1]
    assert n_edges <= full[1].edge_index.shape[1]
    # balanced by level-1 edges: no tile more than 25 % above the mean
    e = [t[1].edge_index.shape[1] for t in tiles]
    assert max(e) <= 1.25 * sum(e) / len(e)


def test_bench_config_is_shared_by_both_arms_and_reference_arm_runs():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--config', 'tiny', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['value'] > 0
    assert line['config'] == bench.bench_config('tiny', 1)
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['cpu_baseline']['kind'] == 'port'
    assert 'AdamW' in line['cpu_baseline']['sample']


def test_flat_gradients_accumulate_over_micro_batches():
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    flat = FlatGradients(list(lin.parameters()))
    x1, x2 = torch.randn(4, 5), torch.randn(6, 5)
    flat.release()
    lin(x1).sum().backward()
    flat.collect()
    flat.release()
    lin(x2).sum().backward()
    flat.collect(accumulate=True)
    got = flat.flat.clone()
    lin.zero_grad(set_to_none=True)
    (lin(x1).sum() + lin(x2).sum()).backward()
    want = torch.cat([p.grad.reshape(-1) for p in lin.parameters()])
    assert torch.allclose(got, want, atol=1e-6)
    assert all(p.grad.data_ptr() != 0 for p in lin.parameters())


def test_strong_scaling_shards_cover_the_step_exactly_once(monkeypatch):
    """cfg4 / cfg5 sharding (scenes / tiles LPT-assigned by edge count): the ranks' shares are
    disjoint, cover the whole step, and every rank derives the same assignment on its own."""
    small4 = dict(bench.BENCH_CONFIGS['cfg4'], levels=[600, 120, 24], scenes=6, scenes_per_batch=2)
    small5 = dict(bench.BENCH_CONFIGS['cfg5'], levels=[8000, 1600, 320], tiles=4)
    monkeypatch.setitem(bench.BENCH_CONFIGS, 'cfg4', small4)
    monkeypatch.setitem(bench.BENCH_CONFIGS, 'cfg5', small5)
    for name, per_item in (('cfg4', 600), ('cfg5', None)):
        world = 2
        shares = [bench.rank_micro_batches(name, r, world) for r in range(world)]
        total_sp = sum(m[3] for mbs, _ in shares for m in mbs)
        infos = [info for _, info in shares]
        assert infos[0]['edges_per_rank'] == infos[1]['edges_per_rank']
        assert sum(i['mine'] for i in infos) == infos[0]['items']
        if per_item:
            assert total_sp == per_item * infos[0]['items']
        else:
            assert total_sp == 8000
        assert infos[0]['imbalance_max_over_mean'] < 1.35
        one = bench.rank_micro_batches(name, 0, 1)
        assert sum(m[3] for m in one[0]) == total_sp
