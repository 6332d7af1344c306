"""Load the REFERENCE's own data containers (src/data/{csr,cluster,instance,data,nag}.py),
unmodified, from /root/reference — TEST INFRASTRUCTURE, build container only (the GPU
box has no /root/reference; the vectors this produces are committed under tests/golden/
by oracle/make_golden_select.py).

Same idea as oracle/reference_shim.py, for the indexing path SURVEY.md §8(f) rank 1
names (`NAG.select` nag.py:306-399, `Data.select` data.py:286-470, `Cluster.select`
cluster.py:79-140, `CSRData.__getitem__` csr.py:328-393): the files are exec'd under a
synthetic `src` package and only the THIRD-PARTY pieces that are absent from this image are
restated here:

    torch_geometric.data.Data / Batch        minimal attribute store (PyG 2.3 semantics of
                                             `keys`, iteration, `num_nodes`, `clone`)
    torch_geometric.data.storage.recursive_apply(_)
    torch_geometric.nn.pool.consecutive.consecutive_cluster
    torch_scatter.*                          oracle/leaves.py
    h5py, numba, omegaconf, torch_cluster    import-only stubs (no I/O, no jit on this path)

`src/transforms/sampling.py` is loaded too (SampleSubNodes, SampleSegments, through
`sparse_sample` of src/utils/sparse.py).  Everything under `src.data` and the helpers it calls (`src/utils/{tensor,sparse,dict,list,
memory}.py`) is the reference's code.  The synthetic modules are removed from sys.modules
again after loading so that oracle/reference_shim.py (which registers its own `src`) can be
used in the same process.
"""
import copy
import importlib.util
import os
import sys
import types
from collections.abc import Mapping, Sequence

import torch

from . import leaves as L

REFERENCE_ROOT = os.environ.get('SPT_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'src', 'data', 'nag.py'))


# ----------------------------------------------------------------------------
# third-party restatements
# ----------------------------------------------------------------------------
def consecutive_cluster(src):
    """torch_geometric/nn/pool/consecutive.py (PyG 2.3.0): relabel `src` to consecutive
    ids in sorted order; `perm[j]` = a position holding the j-th distinct value."""
    unique, inv = torch.unique(src, sorted=True, return_inverse=True)
    perm = torch.arange(inv.size(0), dtype=inv.dtype, device=inv.device)
    perm = inv.new_empty(unique.size(0)).scatter_(0, inv, perm)
    return inv, perm


def recursive_apply(data, func):
    """torch_geometric/data/storage.py recursive_apply (PyG 2.3.0)."""
    if isinstance(data, torch.Tensor):
        return func(data)
    if isinstance(data, tuple) and hasattr(data, '_fields'):
        return type(data)(*(recursive_apply(d, func) for d in data))
    if isinstance(data, Sequence) and not isinstance(data, str):
        return [recursive_apply(d, func) for d in data]
    if isinstance(data, Mapping):
        return {k: recursive_apply(v, func) for k, v in data.items()}
    try:
        return func(data)
    except Exception:
        return data


def recursive_apply_(data, func):
    if isinstance(data, torch.Tensor):
        func(data)
    elif isinstance(data, tuple) and hasattr(data, '_fields'):
        for d in data:
            recursive_apply_(d, func)
    elif isinstance(data, Sequence) and not isinstance(data, str):
        for d in data:
            recursive_apply_(d, func)
    elif isinstance(data, Mapping):
        for d in data.values():
            recursive_apply_(d, func)
    else:
        try:
            func(data)
        except Exception:
            pass


_N_KEYS = {'x', 'feat', 'pos', 'batch', 'node_type', 'n_id'}


class PyGData:
    """The part of torch_geometric.data.Data (2.3.0) the reference's `Data` builds on:
    an ordered attribute store, `keys` as a property, `(key, value)` iteration,
    `num_nodes` inferred from the node-level keys unless set, shallow copy + tensor
    clone."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.__dict__['_store'] = {}
        for k, v in (('x', x), ('edge_index', edge_index), ('edge_attr', edge_attr),
                     ('y', y), ('pos', pos)):
            if v is not None:
                self._store[k] = v
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __getattr__(self, key):
        if key == '_store':
            raise AttributeError(key)
        try:
            return self.__dict__['_store'][key]
        except KeyError:
            raise AttributeError(
                f"'{self.__class__.__name__}' object has no attribute '{key}'") from None

    def __setattr__(self, key, value):
        prop = getattr(self.__class__, key, None)
        if prop is not None and getattr(prop, 'fset', None) is not None:
            prop.fset(self, value)
        elif key.startswith('__') or key in ('_slice_dict', '_inc_dict', '_num_graphs'):
            self.__dict__[key] = value
        elif value is None and key in self._store:
            del self._store[key]
        elif value is not None:
            self._store[key] = value

    def __delattr__(self, key):
        del self._store[key]

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self._store

    def __iter__(self):
        for k, v in list(self._store.items()):
            yield k, v

    def __copy__(self):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = v
        out.__dict__['_store'] = dict(self._store)
        return out

    @property
    def keys(self):
        return list(self._store.keys())

    x = property(lambda self: self._store.get('x'))
    edge_index = property(lambda self: self._store.get('edge_index'))
    edge_attr = property(lambda self: self._store.get('edge_attr'))
    y = property(lambda self: self._store.get('y'))

    @property
    def num_nodes(self):
        if 'num_nodes' in self._store:
            return self._store['num_nodes']
        for k, v in self._store.items():
            if isinstance(v, torch.Tensor) and (k in _N_KEYS or 'node' in k):
                return v.size(self.__cat_dim__(k, v))
        if 'edge_index' in self._store and self._store['edge_index'].numel() > 0:
            return int(self._store['edge_index'].max()) + 1
        return None

    @property
    def num_edges(self):
        ei = self._store.get('edge_index')
        return 0 if ei is None else ei.size(-1)

    def __cat_dim__(self, key, value, *args, **kwargs):
        return -1 if 'index' in key and key != 'batch' else 0

    def __inc__(self, key, value, *args, **kwargs):
        return self.num_nodes if 'index' in key and key != 'batch' else 0

    def to_dict(self):
        return dict(self._store)

    def apply(self, func, *args):
        for k in (args or list(self._store.keys())):
            self._store[k] = recursive_apply(self._store[k], func)
        return self

    def clone(self, *args):
        return copy.copy(self).apply(lambda x: x.clone(), *args)

    def to(self, device, *args, **kwargs):
        return self.apply(lambda x: x.to(device, **kwargs), *args)

    def validate(self, raise_on_error=True):
        return True


class PyGBatch(PyGData):
    """Placeholder base of the reference's `Batch` (collation is not on this path)."""

    @classmethod
    def from_data_list(cls, *a, **k):
        raise NotImplementedError('PyG collation is not restated; see Batch.from_data_list '
                                  'in superpoint_transformer_b200/data/data.py')


def to_undirected(edge_index, *args, **kwargs):
    """torch_geometric.utils.to_undirected (PyG 2.3.0) without attributes: both directions,
    sorted, duplicates removed."""
    row = torch.cat((edge_index[0], edge_index[1]))
    col = torch.cat((edge_index[1], edge_index[0]))
    n = int(max(row.max(), col.max())) + 1 if row.numel() else 1
    key = (row * n + col).unique()
    return torch.stack((key // n, key % n))


def k_hop_subgraph(node_idx, num_hops, edge_index, relabel_nodes=False, num_nodes=None,
                   flow='source_to_target'):
    """torch_geometric.utils.k_hop_subgraph (PyG 2.3.0): (subset, edge_index, inv, edge_mask);
    the reference only reads `subset` (src/transforms/sampling.py:1086-1090)."""
    col, row = edge_index if flow == 'source_to_target' else (edge_index[1], edge_index[0])
    node_idx = node_idx.view(-1)
    subsets = [node_idx]
    for _ in range(num_hops):
        node_mask = torch.zeros(num_nodes, dtype=torch.bool, device=row.device)
        node_mask[subsets[-1]] = True
        subsets.append(col[node_mask[row]])
    subset, inv = torch.cat(subsets).unique(return_inverse=True)
    inv = inv[:node_idx.numel()]
    node_mask = torch.zeros(num_nodes, dtype=torch.bool, device=row.device)
    node_mask[subset] = True
    edge_mask = node_mask[row] & node_mask[col]
    return subset, edge_index[:, edge_mask], inv, edge_mask


def _njit(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
    return lambda f: f


_LOADED = None
_STUB_NAMES = ('h5py', 'numba', 'omegaconf', 'torch_cluster', 'torch_geometric.utils',
               'torch_geometric.nn.pool',
               'torch_geometric.transforms', 'torch_scatter', 'torch_geometric', 'torch_geometric.data',
               'torch_geometric.data.storage', 'torch_geometric.nn', 'torch_geometric.nn.pool',
               'torch_geometric.nn.pool.consecutive')


def load_data():
    """Namespace with the reference's CSRData, CSRBatch, Cluster, InstanceData, Data, NAG,
    NAGBatch and the index helpers they use."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f'reference sources not found under {REFERENCE_ROOT}')

    saved = {k: v for k, v in sys.modules.items()
             if k == 'src' or k.startswith('src.') or k in _STUB_NAMES}
    for k in saved:
        del sys.modules[k]

    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def package(name):
        m = module(name)
        m.__path__ = []
        return m

    def load(name, relpath):
        spec = importlib.util.spec_from_file_location(
            name, os.path.join(REFERENCE_ROOT, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    try:
        module('h5py', File=type('File', (), {}), Group=type('Group', (), {}),
               Dataset=type('Dataset', (), {}))
        module('numba', njit=_njit)
        module('omegaconf', ListConfig=type('ListConfig', (list,), {}))
        module('torch_scatter', scatter=L.scatter, scatter_sum=L.scatter_sum,
               scatter_mean=L.scatter_mean, scatter_min=L.scatter_min,
               scatter_max=L.scatter_max, scatter_std=L.scatter_std)
        package('torch_geometric')
        package('torch_geometric.data').__dict__.update(Data=PyGData, Batch=PyGBatch)
        module('torch_geometric.data.storage', recursive_apply=recursive_apply,
               recursive_apply_=recursive_apply_)
        package('torch_geometric.nn')
        package('torch_geometric.nn.pool')
        module('torch_geometric.nn.pool.consecutive', consecutive_cluster=consecutive_cluster)

        src = package('src')
        src.is_debug_enabled = lambda: False
        utils = package('src.utils')
        for name in ('dict', 'list', 'memory', 'tensor', 'sparse'):
            mod = load(f'src.utils.{name}', f'src/utils/{name}.py')
            public = getattr(mod, '__all__', None) or [
                k for k, v in vars(mod).items()
                if callable(v) and getattr(v, '__module__', None) == mod.__name__]
            for k in public:
                setattr(utils, k, getattr(mod, k))
        # imported by the data modules, never called on the indexing path
        for k in ('save_tensor', 'load_tensor', 'save_tensor_dict', 'load_tensor_dict',
                  'save_dense_to_csr', 'load_csr_to_dense', 'isolated_nodes', 'knn_2',
                  'to_trimmed', 'to_float_rgb', 'to_byte_rgb'):
            setattr(utils, k, None)
        placeholder = type('MetricResults', (), {})
        module('src.metrics', SemanticMetricResults=placeholder,
               PanopticMetricResults=placeholder, InstanceMetricResults=placeholder)

        data = package('src.data')
        ns = types.SimpleNamespace()
        for name in ('tensor_holder', 'csr', 'cluster', 'instance', 'data', 'nag'):
            mod = load(f'src.data.{name}', f'src/data/{name}.py')
            for k in getattr(mod, '__all__', []):
                setattr(data, k, getattr(mod, k))
                setattr(ns, k, getattr(mod, k))
        # sampling transforms (src/transforms/sampling.py): SampleSubNodes / SampleSegments only
        # touch sparse_sample, NAG.get_sampling, NAG.select and torch.multinomial; the voxel /
        # k-hop / radius helpers the module imports at the top are placeholders
        sys.modules['torch_geometric.nn.pool'].voxel_grid = None
        module('torch_geometric.utils', k_hop_subgraph=k_hop_subgraph,
               to_undirected=to_undirected, coalesce=None)
        module('src.dependencies')
        module('src.dependencies.FRNN', frnn=None)
        module('src.utils.scatter', scatter_nearest_neighbor=None)
        nb = load('src.utils.neighbors', 'src/utils/neighbors.py')
        utils.knn_brute_force, utils.knn_2 = nb.knn_brute_force, nb.knn_2
        module('torch_cluster', grid_cluster=None)
        module('torch_geometric.transforms', BaseTransform=type('BaseTransform', (), {}))
        for k in ('scatter_pca', 'sanitize_keys', 'knn_brute_force', 'split_histogram'):
            if not hasattr(utils, k):
                setattr(utils, k, None)
        module('src.utils.histogram', atomic_to_histogram=None)
        transforms = package('src.transforms')
        tr = load('src.transforms.transforms', 'src/transforms/transforms.py')
        transforms.Transform = tr.Transform
        sampling = load('src.transforms.sampling', 'src/transforms/sampling.py')
        ns.SampleSubNodes = sampling.SampleSubNodes
        ns.SampleSegments = sampling.SampleSegments
        ns.SampleEdges = sampling.SampleEdges
        ns.SampleRadiusSubgraphs = sampling.SampleRadiusSubgraphs
        ns.SampleKHopSubgraphs = sampling.SampleKHopSubgraphs
        ns.NAGRestrictSize = sampling.NAGRestrictSize
        ns.sparse_sample = utils.sparse_sample
        ns.consecutive_cluster = consecutive_cluster
        ns.index_select_pointers = ns.CSRData.index_select_pointers
        ns.sizes_to_pointers = utils.sizes_to_pointers
        ns.indices_to_pointers = utils.indices_to_pointers
    finally:
        for k in [k for k in sys.modules
                  if k == 'src' or k.startswith('src.') or k in _STUB_NAMES]:
            del sys.modules[k]
        sys.modules.update(saved)
    _LOADED = ns
    return ns
