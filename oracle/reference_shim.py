"""Load the REFERENCE's own hot-path source files, unmodified, from /root/reference
(TEST INFRASTRUCTURE; only usable in the build container — the GPU box has no
/root/reference, which is why the vectors it produces are committed under
tests/golden/).

`import src` fails in this image (h5py, torch_scatter, torch_geometric, hydra,
lightning ... are not installed), so the files are exec'd one by
one with importlib under a synthetic `src` package, and ONLY the missing
third-party leaves are replaced by the restatements of oracle/leaves.py:

    torch_scatter.{scatter,scatter_sum,scatter_mean,scatter_min,scatter_max,scatter_std}
    torch_geometric.utils.{softmax,degree,coalesce*,add_self_loops}
    torch_geometric.nn.aggr.{Sum,Mean,Max,Min,Std}Aggregation
    torch_geometric.nn.norm.{GraphNorm,LayerNorm,InstanceNorm*}   (* = stub, unused)
    torch_geometric.nn.inits.{ones,zeros}

Every line of reference glue (attention.py, pool.py, norm.py, mlp.py, fusion.py,
dropout.py, unpool.py, transformer.py, stage.py, utils/nn.py, utils/scatter.py,
utils/list.py, utils/version.py, models/components/spt.py) is the reference's.
"""
import importlib.util
import os
import sys
import types

import torch
from torch import nn

from . import leaves as L

REFERENCE_ROOT = os.environ.get('SPT_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'src', 'nn'))


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _package(name):
    m = _module(name)
    m.__path__ = []
    return m


def _load(name, relpath):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


# ----------------------------------------------------------------------------
# leaf shims with the third-party call signatures
# ----------------------------------------------------------------------------
class _Aggregation(nn.Module):
    reduce = None

    def forward(self, x, index=None, ptr=None, dim_size=None, dim=-2):
        return L.aggregate(x, index, dim_size=dim_size, reduce=self.reduce)


def _aggr(name, red):
    return type(name, (_Aggregation,), {'reduce': red})


class _StdAggregation(_Aggregation):
    def forward(self, x, index=None, ptr=None, dim_size=None, dim=-2):
        mean = L.scatter_mean(x, index, 0, dim_size)
        mean2 = L.scatter_mean(x * x, index, 0, dim_size)
        var = mean2 - mean * mean
        out = var.clamp(min=1e-5).sqrt()
        return out.masked_fill(out <= (1e-5) ** 0.5, 0.0)


class _GraphNorm(nn.Module):
    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.in_channels, self.eps = in_channels, eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))
        self.mean_scale = nn.Parameter(torch.ones(in_channels))

    def forward(self, x, batch=None, batch_size=None):
        return L.graph_norm(x, batch, self.weight, self.bias, self.mean_scale, self.eps,
                            batch_size)


class _LayerNorm(nn.Module):
    def __init__(self, in_channels, eps=1e-5, affine=True, mode='graph'):
        super().__init__()
        self.in_channels, self.eps, self.mode = in_channels, eps, mode
        if affine:
            self.weight = nn.Parameter(torch.ones(in_channels))
            self.bias = nn.Parameter(torch.zeros(in_channels))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)

    def forward(self, x, batch=None, batch_size=None):
        if self.mode == 'graph':
            return L.layer_norm_graph(x, batch, self.weight, self.bias, self.eps, batch_size)
        return torch.nn.functional.layer_norm(x, (self.in_channels,), self.weight, self.bias,
                                              self.eps)


class _InstanceNorm(nn.Module):  # imported by src/nn/norm.py:5, never instantiated here
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError


def _ones(t):
    if t is not None:
        t.data.fill_(1.0)


def _zeros(t):
    if t is not None:
        t.data.fill_(0.0)


_LOADED = None


def load():
    """Returns a namespace with the reference classes: SelfAttentionBlock,
    TransformerBlock, Stage, DownNFuseStage, UpNFuseStage, MaxPool, ..., GraphNorm,
    UnitSphereNorm, MLP, SPT, init_weights."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT}")

    # third-party leaves -------------------------------------------------------
    _module('torch_scatter', scatter=L.scatter, scatter_sum=L.scatter_sum,
            scatter_mean=L.scatter_mean, scatter_min=L.scatter_min,
            scatter_max=L.scatter_max, scatter_std=L.scatter_std)
    _package('torch_geometric')
    _module('torch_geometric.utils', softmax=lambda src, index=None, ptr=None,
            num_nodes=None, dim=0: L.segment_softmax(src, index, num_nodes),
            degree=L.degree, coalesce=None, remove_self_loops=None,
            add_self_loops=L.add_self_loops)
    _package('torch_geometric.nn')
    _module('torch_geometric.nn.aggr', SumAggregation=_aggr('SumAggregation', 'sum'),
            MeanAggregation=_aggr('MeanAggregation', 'mean'),
            MaxAggregation=_aggr('MaxAggregation', 'max'),
            MinAggregation=_aggr('MinAggregation', 'min'), StdAggregation=_StdAggregation)
    _module('torch_geometric.nn.norm', LayerNorm=_LayerNorm, InstanceNorm=_InstanceNorm,
            GraphNorm=_GraphNorm)
    _module('torch_geometric.nn.inits', ones=_ones, zeros=_zeros)
    # tooling deps of src/utils/{version,list}.py
    _module('git', Repo=None)
    _module('omegaconf', ListConfig=type('ListConfig', (list,), {}))

    # synthetic `src` package ----------------------------------------------------
    src = _package('src')
    src.__version__ = '3.0.0'          # src/__init__.py:13
    src.is_debug_enabled = lambda: False
    utils = _package('src.utils')
    _load('src.utils.parameter', 'src/utils/parameter.py')
    unn = _load('src.utils.nn', 'src/utils/nn.py')
    _module('src.utils.edge', edge_wise_points=None)   # imported, unused on this path
    usc = _load('src.utils.scatter', 'src/utils/scatter.py')
    uver = _load('src.utils.version', 'src/utils/version.py')
    ulist = _load('src.utils.list', 'src/utils/list.py')
    utils.scatter_mean_weighted = usc.scatter_mean_weighted
    utils.VersionHolder = uver.VersionHolder
    utils.listify_with_reference = ulist.listify_with_reference
    utils.init_weights = unn.init_weights

    nnpkg = _package('src.nn')
    ns = types.SimpleNamespace()
    # same order as src/nn/__init__.py:1-9 (position_encoding / instance skipped:
    # other model families)
    for name in ('norm', 'mlp', 'pool', 'unpool', 'attention', 'fusion', 'dropout',
                 'transformer', 'stage'):
        mod = _load(f'src.nn.{name}', f'src/nn/{name}.py')
        for k in getattr(mod, '__all__', []):
            setattr(nnpkg, k, getattr(mod, k))
            setattr(ns, k, getattr(mod, k))
        if name == 'fusion':
            nnpkg.fusion_factory = mod.fusion_factory

    # SPT: isinstance(nag, NAG) asserts use the product's plain containers
    from superpoint_transformer_b200.data import Data, NAG
    _module('src.data', Data=Data, NAG=NAG)
    _package('src.models')
    _package('src.models.components')
    spt = _load('src.models.components.spt', 'src/models/components/spt.py')
    ns.SPT = spt.SPT
    ns.init_weights = unn.init_weights
    ns.VersionHolder = uver.VersionHolder
    ns.build_qk_scale_func = unn.build_qk_scale_func
    ns.scatter_mean_weighted = usc.scatter_mean_weighted

    # on-the-fly transforms (src/transforms/graph.py): the module imports a long
    # list of preprocessing utilities at the top; only those the on-the-fly
    # functions touch are real, the rest are None placeholders.
    keys = _load('src.utils.keys', 'src/utils/keys.py')
    for k in keys.__all__:
        setattr(utils, k, getattr(keys, k))
    for k in ('print_tensor_info', 'isolated_nodes', 'edge_to_superedge', 'subedges',
              'to_trimmed', 'cluster_radius_nn_graph',
              'scatter_mean_orientation', 'geometric_features', 'arange_interleave',
              'csr_to_dense'):
        setattr(utils, k, None)
    # src/utils/geometry.py (base_vectors_3d, used by _minimalistic_horizontal_edge_features):
    # its own third-party imports (pgeof) and the neighbour helpers are stubs, the function
    # itself is the reference's
    _module('pgeof')
    _module('src.utils.neighbors', neighbors_dense_to_csr=None)
    if not hasattr(usc, 'scatter_pca'):
        usc.scatter_pca = None
    geom = _load('src.utils.geometry', 'src/utils/geometry.py')
    utils.base_vectors_3d = geom.base_vectors_3d

    def is_trimmed(edge_index, return_trimmed=False):
        # stand-in for src/utils/graph.py:505-521 (needs PyG coalesce): i<j, no dups
        i, j = edge_index[0], edge_index[1]
        uid = i * (int(edge_index.max()) + 1 if edge_index.numel() else 1) + j
        return bool((i < j).all()) and uid.unique().numel() == uid.numel()

    utils.is_trimmed = is_trimmed

    class Transform:  # minimal src/transforms/transforms.py:12-49
        def __call__(self, x):
            return self._process(x)

    _module('src.transforms', Transform=Transform)
    graph = _load('src.transforms.graph', 'src/transforms/graph.py')
    ns.on_the_fly_horizontal_edge_features = graph._on_the_fly_horizontal_edge_features
    ns.NAGAddSelfLoops = graph.NAGAddSelfLoops
    ns.on_the_fly_vertical_edge_features = graph._on_the_fly_vertical_edge_features
    ns.minimalistic_horizontal_edge_features = graph._minimalistic_horizontal_edge_features
    ns.base_vectors_3d = geom.base_vectors_3d
    _LOADED = ns
    return ns
