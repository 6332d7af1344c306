"""Data: one NAG level (PyG-free holder with the attribute names and helpers of
reference src/data/data.py that the hot path touches: x, pos, edge_index,
edge_attr, super_index, sub, node_size, v_edge_attr, batch, norm_index(),
num_nodes, add_keys_to())."""
import copy

import torch

from .cluster import Cluster, CSRData
from ..utils.tensor import tensor_idx, is_arange

__all__ = ['Data', 'Batch']


class Data:
    def __init__(self, **kwargs):
        object.__setattr__(self, '_store', {})
        for k, v in kwargs.items():
            self[k] = v

    # attribute / item access ------------------------------------------------
    def __getattr__(self, key):
        if key.startswith('__'):
            raise AttributeError(key)
        store = object.__getattribute__(self, '_store')
        if key in store:
            return store[key]
        raise AttributeError(f"'Data' object has no attribute '{key}'")

    def __setattr__(self, key, value):
        if value is None:
            self._store.pop(key, None)
        else:
            self._store[key] = value

    def __delattr__(self, key):
        self._store.pop(key, None)

    def __getitem__(self, key):
        return self._store.get(key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self._store

    @property
    def keys(self):
        return list(self._store.keys())

    # well-known fields default to None ----------------------------------------
    def _get(self, key):
        return self._store.get(key)

    pos = property(lambda self: self._get('pos'))
    x = property(lambda self: self._get('x'))
    edge_index = property(lambda self: self._get('edge_index'))
    edge_attr = property(lambda self: self._get('edge_attr'))
    super_index = property(lambda self: self._get('super_index'))
    sub = property(lambda self: self._get('sub'))
    node_size = property(lambda self: self._get('node_size'))
    v_edge_attr = property(lambda self: self._get('v_edge_attr'))
    batch = property(lambda self: self._get('batch'))
    diameter = property(lambda self: self._get('diameter'))
    normal = property(lambda self: self._get('normal'))

    for _k in ('pos', 'x', 'edge_index', 'edge_attr', 'super_index', 'sub', 'node_size',
               'v_edge_attr', 'batch', 'diameter', 'normal'):
        locals()[_k] = locals()[_k].setter(
            (lambda k: lambda self, v: Data.__setattr__(self, k, v))(_k))
    del _k

    @property
    def edge_keys(self):
        """Keys starting with `edge_` other than edge_index / edge_attr
        (reference src/data/data.py:158-164)."""
        return [k for k in self.keys
                if k.startswith('edge_') and k not in ('edge_index', 'edge_attr')]

    def raise_if_edge_keys(self):
        if len(self.edge_keys) > 0:
            raise NotImplementedError(
                f"Edge keys are not supported, stack them in `edge_attr`: {self.edge_keys}")

    # derived ------------------------------------------------------------------
    @property
    def v_edge_keys(self):
        """reference src/data/data.py:178-180"""
        return [k for k in self.keys if k.startswith('v_edge_')]

    @property
    def num_nodes(self):
        if '_num_nodes' in self._store:
            return self._store['_num_nodes']
        for key in ('pos', 'x', 'super_index', 'node_size', 'batch'):
            t = self._get(key)
            if t is not None:
                return int(t.shape[0])
        if self.sub is not None:
            return self.sub.num_clusters
        if self.edge_index is not None and self.edge_index.numel() > 0:
            return int(self.edge_index.max()) + 1
        return 0

    @property
    def num_points(self):
        return self.num_nodes

    @property
    def num_edges(self):
        return 0 if self.edge_index is None else int(self.edge_index.shape[1])

    @property
    def has_edges(self):
        return self.edge_index is not None and self.edge_index.shape[1] > 0

    @property
    def is_super(self):
        return self.sub is not None

    @property
    def is_sub(self):
        return self.super_index is not None

    @property
    def num_super(self):
        return int(self.super_index.max()) + 1 if self.is_sub else 0

    @property
    def device(self):
        for v in self._store.values():
            if isinstance(v, (torch.Tensor, Cluster)):
                return v.device
        return torch.device('cpu')

    def norm_index(self, mode='graph'):
        """Index used by the index-based norms (reference src/data/data.py:103-130)."""
        n, dev = self.num_nodes, self.device
        batch = self.batch if self.batch is not None else \
            self._cached_zeros(n, dev)
        if mode == 'graph':
            return batch
        if mode == 'node':
            return torch.arange(n, device=dev)
        if mode == 'segment':
            sup = self.super_index if self.super_index is not None else \
                torch.zeros(n, dtype=torch.long, device=dev)
            return sup * (batch.max() + 1) + batch
        raise NotImplementedError(f"Unknown mode='{mode}'")

    def _cached_zeros(self, n, dev):
        z = self._store.get('_zero_batch')
        if z is None or z.shape[0] != n or z.device != dev:
            z = torch.zeros(n, dtype=torch.long, device=dev)
            self._store['_zero_batch'] = z
            if z.is_cuda:
                # single graph: tell the norm kernels the segment count up front (no
                # batch.max() host sync -> the forward stays CUDA-graph capturable)
                from .. import ops
                ops.register_num_segments(z, 1)
        return z

    def add_keys_to(self, keys, to='x', strict=True, delete_after=False):
        """Concatenate attributes `keys` into `to` (reference src/data/data.py:1097-1141)."""
        if keys is None or len(keys) == 0:
            return
        prev = self._get(to)
        feats = [prev] if prev is not None else []
        for key in keys:
            feat = self._get(key)
            if feat is None:
                if strict:
                    raise Exception(f"Data should contain the attribute '{key}'")
                continue
            if delete_after:
                delattr(self, key)
            if prev is not None and prev.shape[0] != feat.shape[0]:
                raise Exception(f"The tensors '{to}' and '{key}' can't be concatenated")
            feats.append(feat.unsqueeze(-1) if feat.dim() == 1 else feat)
        setattr(self, to, torch.cat(feats, dim=1))

    def to(self, device, non_blocking=False):
        out = self.__class__()
        for k, v in self._store.items():
            if isinstance(v, torch.Tensor):
                out._store[k] = v.to(device, non_blocking=non_blocking)
            elif isinstance(v, Cluster):
                out._store[k] = v.to(device, non_blocking=non_blocking)
            else:
                out._store[k] = v
        out._store.pop('_zero_batch', None)
        b, ng = out._store.get('batch'), out._store.get('_num_graphs')
        if b is not None and ng is not None and b.is_cuda:
            from .. import ops
            ops.register_num_segments(b, ng)
        return out

    def cuda(self, non_blocking=False):
        return self.to('cuda', non_blocking=non_blocking)

    def cpu(self):
        return self.to('cpu')

    def clone(self):
        out = self.__class__()
        out._store.update(self._store)
        return out

    @classmethod
    def load(cls, f, keys=None, non_fp_to_long=False, rgb_to_float=False, **kwargs):
        """One level from a file / group written by the reference's `Data.save` (reference
        src/data/data.py:736-940); see io/nag_io.py."""
        from ..io import load_data, H5File
        if isinstance(f, str):
            f = H5File(f)
        return load_data(f, keys=keys, non_fp_to_long=non_fp_to_long, rgb_to_float=rgb_to_float)

    def save(self, path, y_to_csr=True, pos_dtype=torch.float, fp_dtype=torch.float,
             rgb_to_byte=True):
        """This level alone as a file of the reference's format (reference
        src/data/data.py:663-734: datasets at the root of the file)."""
        from ..io import data_to_tree, write_h5
        write_h5(path, data_to_tree(self, y_to_csr=y_to_csr, pos_dtype=pos_dtype,
                                    fp_dtype=fp_dtype, rgb_to_byte=rgb_to_byte))

    def select(self, idx, update_sub=True, update_super=True, _num_super=None,
               _skip_sub=False, _skip_super=False):
        """Nodes `idx` (duplicate-free) of this level, with edges re-indexed and restricted to
        the selection, `sub` restricted and — with `update_sub` / `update_super` — the ids of
        the neighbouring levels made dense again (reference src/data/data.py:286-470).
        Returns data, (idx_sub, sub_super), (idx_super, super_sub) exactly like the reference:
        `idx_sub` / `idx_super` select the level below / above, `sub_super` replaces the lower
        level's `super_index`, `super_sub` the upper level's `sub`.

        CUDA tensors only: the integer work runs in csrc/select.cu (edge compaction and CSR
        selection by scan, relabels by bitmap + scan instead of `consecutive_cluster`'s sort).
        Inside a cluster of `super_sub` the points are in ascending order (the reference's
        unstable torch.sort leaves that order unspecified, src/utils/sparse.py:31-33).
        `_num_super`, `_skip_sub`, `_skip_super`: NAG.select's shortcuts (parent count known;
        `sub` / `super_index` are about to be replaced by the neighbouring level's result)."""
        device = self.device
        idx = tensor_idx(idx, device=device)
        num_nodes = self.num_nodes
        if idx is None or is_arange(idx, num_nodes):
            return self.clone(), (None, None), (None, None)
        from .. import ops
        num_sel = idx.shape[0]
        has_edges, num_edges = self.has_edges, self.num_edges
        use_sub = self.is_super and not _skip_sub
        if use_sub and not isinstance(self.sub, CSRData):
            raise NotImplementedError("Data.select needs `sub` as a Cluster")
        gather_super = self.is_sub and not _skip_super
        relabel_super = self.is_sub and update_super
        if relabel_super and not gather_super:
            raise ValueError("update_super needs the level's own super_index")
        num_super = None
        if relabel_super:
            num_super = self.num_super if _num_super is None else int(_num_super)

        # which attribute follows the nodes, which the edges, which is copied (data.py:420-463)
        skip_keys = ('edge_index', 'sub', 'super_index', 'neighbor_index',
                     'neighbor_distance')
        edge_keys = ['edge_attr'] + self.edge_keys
        v_edge_keys = self.v_edge_keys
        order, node_items, edge_items, csr_items, copied = [], [], [], [], []
        for key, item in list(self._store.items()):
            if key in skip_keys or key.startswith('_'):
                continue
            order.append(key)
            is_tensor = torch.is_tensor(item)
            is_node_size = is_tensor and item.dim() > 0 and item.shape[0] == num_nodes
            is_edge_size = is_tensor and item.dim() > 0 and item.shape[0] == num_edges
            if isinstance(item, CSRData):
                csr_items.append(key)
            elif is_node_size and key in v_edge_keys:
                node_items.append(key)
            elif has_edges and is_edge_size and key in edge_keys:
                edge_items.append(key)
            elif is_node_size:
                node_items.append(key)
            else:
                copied.append(key)

        data = Data()
        out_sub, out_super = (None, None), (None, None)
        values = {}
        if idx.is_cuda and ops.SELECT_FUSED:
            # one native call for the level (csrc/select.cu: spt_data_select)
            res = ops.data_select(
                num_nodes, idx, edge_index=self.edge_index if has_edges else None,
                sub=(self.sub.pointers, self.sub.points) if use_sub else None,
                update_sub=update_sub,
                super_index=self.super_index if gather_super else None,
                num_super=num_super if relabel_super else 0, update_super=relabel_super,
                node_rows=[self._store[k] for k in node_items],
                edge_rows=[self._store[k] for k in edge_items])
            if has_edges:
                data.edge_index = res['edge_index']
            if use_sub:
                data.sub = Cluster(*res['sub'])
                if update_sub:
                    out_sub = (res['idx_sub'], res['sub_super'])
            if gather_super:
                data.super_index = res['super_index']
                if relabel_super:
                    out_super = (res['idx_super'], Cluster(*res['super_sub']))
            values.update(zip(node_items, res['node_rows']))
            values.update(zip(edge_items, res['edge_rows']))
        else:
            # the same steps through the primitives, one by one
            idx_edge = None
            if has_edges:   # re-index, drop edges that lose an end point; validates idx too
                data.edge_index, idx_edge = ops.select_edges(self.edge_index, idx, num_nodes)
            else:
                ops.select_edges(None, idx, num_nodes)
            if use_sub:
                data.sub, out_sub = self.sub.select(idx, update_sub=update_sub)
            if gather_super:
                data.super_index = ops.take_rows(self.super_index, idx)
            if relabel_super:
                data.super_index, idx_super = ops.relabel_consecutive(data.super_index,
                                                                      num_super)
                super_sub = Cluster.from_super_index(data.super_index, idx_super.shape[0])
                out_super = (idx_super, super_sub)
            for keys, index in ((node_items, idx), (edge_items, idx_edge)):
                values.update(zip(keys, ops.take_rows_multi([self._store[k] for k in keys],
                                                            index)))
        for key in order:                      # the reference's attribute order
            item = self._store[key]
            if key in values:
                data[key] = values[key]
            elif key in csr_items:
                data[key] = item.select(idx)
            else:
                data[key] = item.clone() if torch.is_tensor(item) else copy.deepcopy(item)

        if data.num_nodes != num_sel:
            data._store['_num_nodes'] = num_sel
        return data, out_sub, out_super

    def __repr__(self):
        parts = []
        for k, v in self._store.items():
            if k.startswith('_'):
                continue
            parts.append(f"{k}={list(v.shape)}" if isinstance(v, torch.Tensor) else f"{k}={v}")
        return f"Data({', '.join(parts)})"


class Batch(Data):
    """Disjoint union of Data objects of one level.  Node-indexed integer
    attributes are offset like reference Data.__inc__ (src/data/data.py:267-274):
    edge_index by the node count, super_index by the parent count, Cluster.points
    by the child count; `batch` records the item of every node."""

    @classmethod
    def from_data_list(cls, data_list, num_super_list=None):
        out = cls()
        keys = [k for k in data_list[0].keys if not k.startswith('_')]
        n_nodes = [d.num_nodes for d in data_list]
        node_off = [0]
        for n in n_nodes:
            node_off.append(node_off[-1] + n)
        if num_super_list is None and data_list[0].super_index is not None:
            num_super_list = [int(d.super_index.max()) + 1 for d in data_list]
        sup_off = [0]
        for n in (num_super_list or []):
            sup_off.append(sup_off[-1] + n)
        on_dev = data_list[0].device.type == 'cuda'
        if on_dev:
            from .. import ops
        for k in keys:
            vals = [d[k] for d in data_list]
            if k == 'edge_index':
                if on_dev:   # one offset-concat launch per row of the [2, E] index
                    out[k] = torch.stack((
                        ops.concat_offset([v[0] for v in vals], node_off[:-1]),
                        ops.concat_offset([v[1] for v in vals], node_off[:-1])))
                else:
                    out[k] = torch.cat([v + node_off[i] for i, v in enumerate(vals)], dim=1)
            elif k == 'super_index':
                out[k] = ops.concat_offset(vals, sup_off[:-1]) if on_dev else \
                    torch.cat([v + sup_off[i] for i, v in enumerate(vals)])
            elif k == 'sub':
                if on_dev:
                    child_off = [0]
                    for c in vals:
                        child_off.append(child_off[-1] + c.num_points)
                    out[k] = Cluster(
                        ops.concat_offset([c.pointers for c in vals], child_off[:-1],
                                          skip_first=True),
                        ops.concat_offset([c.points for c in vals], child_off[:-1]))
                    continue
                child_off, ptrs, pts = 0, [], []
                ptr_off = 0
                for i, c in enumerate(vals):
                    p = c.pointers + ptr_off
                    ptrs.append(p if i == 0 else p[1:])
                    pts.append(c.points + child_off)
                    ptr_off += c.num_points
                    child_off += c.num_points
                out[k] = Cluster(torch.cat(ptrs), torch.cat(pts))
            elif isinstance(vals[0], torch.Tensor):
                out[k] = torch.cat(vals, dim=0)
            else:
                out[k] = vals[0]
        dev = out.device
        if on_dev:
            out['batch'] = ops.segment_ids(n_nodes, dev)
        else:
            out['batch'] = torch.repeat_interleave(
                torch.arange(len(data_list), device=dev),
                torch.tensor(n_nodes, device=dev))
        out['_num_graphs'] = len(data_list)
        if out['batch'].is_cuda:
            from .. import ops
            ops.register_num_segments(out['batch'], len(data_list))
        return out

    @property
    def num_graphs(self):
        return self._store.get('_num_graphs', 1)
