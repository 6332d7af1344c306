"""Loading the reference's on-disk NAG format (read side): `NAG.load` ->
`Data.load` -> `Cluster.load` / `load_csr_to_dense` / `load_tensor` (reference
src/data/nag.py:436-570, src/data/data.py:736-940, src/data/csr.py:495-610,
src/utils/io.py:70-299), on top of the HDF5 subset reader of io/h5lite.py (no h5py here).

File layout (written by `NAG.save`): root attribute `start_i_level`; one group `level_<i>` per
partition level holding one dataset per tensor attribute, `_csr_/<key>/{pointers, columns,
values, shape}` for sparse-saved dense matrices (label histograms `y`), `_cluster_/<key>/
{pointers, value_0, is_index_value}` for `Cluster` objects (`sub`), `_not_indexable_` (names);
integers are stored in the smallest dtype that holds them.

Host-side I/O (CPU tensors come back; `.cuda()` them).  Row selection at load time (`idx`) is
not offered: load, then `NAG.select` on the device."""
import torch

from .h5lite import H5File, H5Group
from .h5write import write_h5
from ..data.cluster import Cluster
from ..data.data import Data
from ..data.nag import NAG

__all__ = ['load_tensor', 'load_csr_to_dense', 'load_cluster', 'load_data', 'load_nag',
           'save_nag', 'data_to_tree', 'cluster_to_tree']

LEVEL_PREFIX = 'level_'              # NAG._data_serialization_prefix
START_KEY = 'start_i_level'          # NAG._start_i_level_serialization_key


def load_tensor(dataset, non_fp_to_long=False):
    """reference src/utils/io.py:70-120: the stored array; integers back to int64 on request."""
    x = torch.from_numpy(dataset.read())
    if not x.is_floating_point() and non_fp_to_long:
        x = x.long()
    return x


def load_csr_to_dense(group, non_fp_to_long=False):
    """reference src/utils/io.py:205-253 + csr_to_dense (src/utils/sparse.py:63-87): a dense
    [n, m] matrix saved as CSR (pointers, columns, values, shape), zeros elsewhere."""
    assert all(k in group for k in ('pointers', 'columns', 'values', 'shape'))
    pointers = load_tensor(group['pointers'], True)
    columns = load_tensor(group['columns'], True)
    values = load_tensor(group['values'], non_fp_to_long)
    shape = load_tensor(group['shape'], True).tolist()
    n = max(shape[0], pointers.shape[0] - 1)
    m = max(shape[1], int(columns.max()) + 1 if columns.numel() else 0)
    out = torch.zeros((n, m), dtype=values.dtype)
    rows = torch.arange(pointers.shape[0] - 1).repeat_interleave(pointers[1:] - pointers[:-1])
    out[rows, columns] = values
    return out


def load_cluster(group, non_fp_to_long=False):
    """reference src/data/csr.py:495-575 (no-indexing branch) for a Cluster: pointers +
    `value_0` (the point ids)."""
    assert 'pointers' in group and 'value_0' in group and 'is_index_value' in group
    if 'value_1' in group:
        raise NotImplementedError(f'{group.name}: CSRData with several value tensors')
    return Cluster(load_tensor(group['pointers'], non_fp_to_long),
                   load_tensor(group['value_0'], non_fp_to_long))


def load_data(group, keys=None, non_fp_to_long=False, rgb_to_float=False):
    """One level (reference src/data/data.py:736-940, no-indexing branch).  `keys`: attributes
    to read (default: all).  `_instance_data_` (instance labels) is outside this package's scope
    and is skipped."""
    special = ('_not_indexable_', '_csr_', '_cluster_', '_instance_data_', '_slice_dict',
               '_inc_dict', '_num_graphs')
    names = group.keys()
    csr_keys = group['_csr_'].keys() if '_csr_' in group else []
    cluster_keys = group['_cluster_'].keys() if '_cluster_' in group else []
    if keys is None:
        keys = [k for k in names if k not in special] + csr_keys + cluster_keys
    out = {}
    for k in names:
        if k in special or k not in keys:
            continue
        obj = group[k]
        if isinstance(obj, H5Group):
            raise NotImplementedError(f'{obj.name}: nested group')
        out[k] = load_tensor(obj, non_fp_to_long)
    for k in csr_keys:
        if k in keys:
            out[k] = load_csr_to_dense(group['_csr_'][k], non_fp_to_long)
    for k in cluster_keys:
        if k in keys:
            out[k] = load_cluster(group['_cluster_'][k], non_fp_to_long)
    for k in ('rgb', 'mean_rgb'):       # src/utils/color.py:17-29
        if k in out:
            rgb = out[k]
            if rgb_to_float:
                rgb = rgb.float()
                rgb = (rgb / 255 if rgb.numel() and rgb.max() > 1 else rgb).clamp(min=0, max=1)
            else:
                if rgb.is_floating_point() and rgb.max() <= 1:
                    rgb = rgb * 255
                rgb = rgb.clamp(min=0, max=255).byte()
            out[k] = rgb
    return Data(**out)


def load_nag(path, low=0, high=-1, idx=None, keys_low=None, keys=None, non_fp_to_long=False,
             rgb_to_float=False):
    """reference src/data/nag.py:436-570: levels `low`..`high` of the file as a NAG whose
    `start_i_level` is `low`."""
    if idx is not None:
        raise NotImplementedError("row selection at load time: load, then NAG.select")
    keys_low = keys if keys_low is None and keys is not None else keys_low
    with H5File(path) as f:
        saved_start = int(f.attrs.get(START_KEY, 0))
        assert low >= saved_start, "Trying to load low levels that are not saved in the file"
        if not any(LEVEL_PREFIX in k for k in f.keys()):
            return load_data(f, keys=keys_low, non_fp_to_long=non_fp_to_long,
                             rgb_to_float=rgb_to_float)
        high = saved_start + len(f) - 1 if high < 0 else high
        assert high <= saved_start + len(f) - 1, \
            "Trying to load high levels that are not saved in the file"
        levels = []
        for i in range(low, high + 1):
            assert f'{LEVEL_PREFIX}{i}' in f, f'level {i} missing from {path}'
            levels.append(load_data(f[f'{LEVEL_PREFIX}{i}'], keys=keys_low if i == low else keys,
                                    non_fp_to_long=non_fp_to_long, rgb_to_float=rgb_to_float))
    return NAG(levels, start_i_level=low)


# ------------------------------------------------------------------------------------- save
def _smallest_int(a):
    """reference src/utils/tensor.py:223-242: the smallest of uint8 / int16 / int32 / int64 that
    holds the values."""
    if a.numel() == 0:
        return a.byte()
    lo, hi = int(a.min()), int(a.max())
    for dtype in (torch.uint8, torch.int16, torch.int32, torch.int64):
        info = torch.iinfo(dtype)
        if info.min <= lo and hi <= info.max:
            return a.to(dtype)
    raise ValueError(f"Could not cast dtype={a.dtype} to integer.")


def _array(x, fp_dtype=torch.float):
    """`cast_numpyfy` (reference src/utils/tensor.py:268-285): floats to `fp_dtype`, integers to
    the smallest integer dtype."""
    x = x.detach().cpu()
    return (x.to(fp_dtype) if x.is_floating_point() else _smallest_int(x)).contiguous().numpy()


def cluster_to_tree(cluster, fp_dtype=torch.float):
    """`CSRData.save` (reference src/data/csr.py:456-493) of a Cluster."""
    return {'pointers': _array(cluster.pointers, fp_dtype),
            'is_index_value': _array(torch.tensor([True]), fp_dtype),
            'value_0': _array(cluster.points, fp_dtype)}


def data_to_tree(data, y_to_csr=True, pos_dtype=torch.float, fp_dtype=torch.float,
                 rgb_to_byte=True):
    """`Data.save` (reference src/data/data.py:663-734): the datasets and sub-groups of one
    level as a nested dict."""
    tree, num_nodes, not_indexable = {}, data.num_nodes, []
    for k in data.keys:
        if k.startswith('_'):
            continue
        val = data[k]
        node_level = torch.is_tensor(val) and val.dim() > 0 and val.shape[0] == num_nodes
        if not node_level or k in ('edge_index', 'edge_attr'):
            not_indexable.append(k)
        if k == 'pos_offset':
            tree[k] = _array(val, torch.double)
        elif k == 'pos':
            tree[k] = _array(val, pos_dtype)
        elif k == 'y' and val.dim() > 1 and y_to_csr:       # save_dense_to_csr, io.py:168-203
            rows, columns = val.nonzero(as_tuple=True)
            pointers = torch.zeros(val.shape[0] + 1, dtype=torch.long)
            pointers[1:] = torch.bincount(rows, minlength=val.shape[0]).cumsum(0)
            tree.setdefault('_csr_', {})[k] = {
                'pointers': _array(pointers, fp_dtype), 'columns': _array(columns, fp_dtype),
                'values': _array(val[rows, columns], fp_dtype),
                'shape': torch.tensor(val.shape).numpy()}
        elif k in ('rgb', 'mean_rgb') and rgb_to_byte:
            tree[k] = _array((val * 255).byte() if val.is_floating_point() else val.byte(),
                             fp_dtype)
        elif isinstance(val, Cluster):
            tree.setdefault('_cluster_', {})[k] = cluster_to_tree(val, fp_dtype)
        elif torch.is_tensor(val):
            tree[k] = _array(val, fp_dtype)
        else:
            raise NotImplementedError(
                f"Cannot save attribute {k} with unsupported type {type(val)}")
    tree['_not_indexable_'] = not_indexable
    return tree


def save_nag(nag, path, y_to_csr=True, pos_dtype=torch.float, fp_dtype=torch.float,
             rgb_to_byte=True):
    """`NAG.save` (reference src/data/nag.py:401-434): one `level_<i>` group per level, the
    start level as a root attribute."""
    tree = {f'{LEVEL_PREFIX}{i}': data_to_tree(nag[i], y_to_csr=y_to_csr, pos_dtype=pos_dtype,
                                              fp_dtype=fp_dtype, rgb_to_byte=rgb_to_byte)
            for i in nag.level_range}
    write_h5(path, tree, {START_KEY: nag.start_i_level})
