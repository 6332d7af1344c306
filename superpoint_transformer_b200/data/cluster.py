"""Cluster: CSR child lists of a partition level (API surface of reference
src/data/cluster.py:19-77 / src/data/csr.py:48-248 that the hot path reads).

`nag[l].sub` is exactly the CSR used to pool level l-1 into level l:
pointers [Np+1], points [Nc] (a permutation of the children, grouped by parent,
ascending inside a group).  `Cluster.from_super_index` builds it on the GPU with
the same stable grouping kernel the attention CSR uses (bit-exact w.r.t.
torch.sort(stable=True))."""
import torch

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


class Cluster(CSRData):
    def __init__(self, pointers, points, dense=False, **kwargs):
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
