"""Cluster: CSR child lists of a partition level (API surface of reference
src/data/cluster.py:19-77 / src/data/csr.py:48-248 that the hot path reads).

`nag[l].sub` is exactly the CSR used to pool level l-1 into level l:
pointers [Np+1], points [Nc] (a permutation of the children, grouped by parent,
ascending inside a group).  `Cluster.from_super_index` builds it on the GPU with
the same stable grouping kernel the attention CSR uses (bit-exact w.r.t.
torch.sort(stable=True))."""
import torch

from ..utils.tensor import tensor_idx, is_arange, indices_to_pointers

__all__ = ['CSRData', 'Cluster']


class CSRData:
    """pointers + one value tensor per item (minimal: a single `values[0]`)."""

    def __init__(self, pointers, *values):
        self.pointers = pointers
        self.values = list(values)

    @property
    def device(self):
        return self.pointers.device

    @property
    def num_groups(self):
        return self.pointers.shape[0] - 1

    @property
    def num_items(self):
        return int(self.values[0].shape[0]) if self.values else 0

    @property
    def sizes(self):
        return self.pointers[1:] - self.pointers[:-1]

    @property
    def indices(self):
        """Group id of every item, in storage order."""
        return torch.repeat_interleave(
            torch.arange(self.num_groups, device=self.device), self.sizes.long())

    def to(self, device, **kwargs):
        out = self.__class__.__new__(self.__class__)
        out.pointers = self.pointers.to(device, **kwargs)
        out.values = [v.to(device, **kwargs) for v in self.values]
        return out

    def cuda(self, **kwargs):
        return self.to('cuda', **kwargs)

    def cpu(self):
        return self.to('cpu')

    def __len__(self):
        return self.num_groups

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        out.pointers = self.pointers.clone()
        out.values = [v.clone() for v in self.values]
        return out

    @staticmethod
    def index_select_pointers(pointers, indices):
        """(pointers_new, val_idx): pointers of the groups `indices` and the positions of their
        items in the value tensors (reference src/data/csr.py:328-356)."""
        from .. import ops
        items = torch.arange(int(pointers[-1]), device=pointers.device)
        return ops.csr_select(pointers, items, indices)

    def __getitem__(self, idx):
        """Copy of self restricted to the groups `idx` (reference src/data/csr.py:358-393).
        An `idx` equal to arange(num_groups) returns a plain copy (the reference returns an
        EMPTY object on that branch, csr.py:371-378 — unreachable from Data.select / NAG.select,
        which return a clone before getting here)."""
        idx = tensor_idx(idx, device=self.device)
        if idx is None or is_arange(idx, self.num_groups):
            return self.clone()
        from .. import ops
        out = self.__class__.__new__(self.__class__)
        if len(self.values) == 1 and self.values[0].dtype == torch.int64 \
                and self.values[0].dim() == 1:
            out.pointers, v = ops.csr_select(self.pointers, self.values[0], idx)
            out.values = [v]
        else:
            out.pointers, val_idx = self.index_select_pointers(self.pointers, idx)
            out.values = [ops.take_rows(v, val_idx) for v in self.values]
        return out

    def select(self, idx, **kwargs):
        """reference src/data/csr.py:395-408"""
        return self[idx]


class Cluster(CSRData):
    def __init__(self, pointers, points, dense=False, **kwargs):
        """dense=True: `pointers` holds the cluster id of every point (reference
        src/data/csr.py:83-85) and is converted to CSR (stable order inside a cluster)."""
        if dense:
            pointers, order = indices_to_pointers(pointers)
            points = points[order]
        super().__init__(pointers, points)

    @property
    def points(self):
        return self.values[0]

    @points.setter
    def points(self, points):
        self.values[0] = points

    @property
    def num_clusters(self):
        return self.num_groups

    @property
    def num_points(self):
        return self.num_items

    def to_super_index(self):
        """Inverse view: parent id of every child (reference cluster.py:67-77)."""
        out = torch.empty(self.num_items, dtype=torch.long, device=self.device)
        out[self.points.long()] = self.indices
        return out

    def select(self, idx, update_sub=True, num_sub=None, **kwargs):
        """Clusters `idx` (duplicate-free) with the point ids made dense again (reference
        src/data/cluster.py:79-140).  Returns cluster, (idx_sub, sub_super): `idx_sub` selects
        the surviving points on the level below (ascending old ids), `sub_super` is that
        level's new `super_index`.

        Device path (csrc/select.cu): CSR group selection, then a bitmap + scan relabel in
        place of the reference's `consecutive_cluster` sort (its "bottleneck" note at
        cluster.py:128-130).  `num_sub` = number of points of the level below (default: the
        points of `self` are a permutation of [0, num_points))."""
        idx = tensor_idx(idx, device=self.device)
        if idx is None or is_arange(idx, self.num_clusters):
            return self.clone(), (None, None)
        from .. import ops
        if not update_sub:
            return Cluster(*ops.csr_select(self.pointers, self.points, idx)), (None, None)
        pointers, points, group = ops.csr_select(self.pointers, self.points, idx,
                                                 want_group=True)
        num_sub = self.num_points if num_sub is None else int(num_sub)
        new_points, idx_sub, sub_super = ops.relabel_consecutive(points, num_sub, payload=group)
        return Cluster(pointers, new_points), (idx_sub, sub_super)

    @classmethod
    def load(cls, f, non_fp_to_long=False, **kwargs):
        """From a `_cluster_/<key>` group written by the reference's `Cluster.save` (reference
        src/data/csr.py:495-575); returns cluster, (None, None) like the reference's
        no-indexing branch (src/data/cluster.py:204-219)."""
        from ..io import load_cluster
        return load_cluster(f, non_fp_to_long=non_fp_to_long), (None, None)

    @classmethod
    def from_super_index(cls, super_index, num_super):
        """CSR of `super_index` (int64 pointers/points like the reference).  CUDA
        tensors go through libspt_b200's stable grouping kernel; CPU tensors (data
        preparation only) through a stable sort."""
        if super_index.is_cuda:
            from .. import ops
            seg = ops.segment_index(super_index, num_super)
            return cls(seg.ptr.long(), seg.perm.long())
        order = torch.sort(super_index, stable=True).indices
        counts = torch.bincount(super_index, minlength=num_super)
        pointers = torch.zeros(num_super + 1, dtype=torch.long)
        pointers[1:] = counts.cumsum(0)
        return cls(pointers, order)

    def __repr__(self):
        return f'Cluster(num_clusters={self.num_clusters}, num_points={self.num_points})'
