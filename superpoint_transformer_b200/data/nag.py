"""NAG: nested partition hierarchy (API surface of reference src/data/nag.py
used by SPT.forward: absolute-level indexing, start/end levels, add_keys_to,
get_sub_size / get_super_index)."""
import torch

from .data import Data, Batch
from ..utils.tensor import tensor_idx, is_arange

__all__ = ['NAG', 'NAGBatch']


def _fill_levels(level, default, value, length, start):
    """'all' / 'i+' / 'i-' / int level selectors (reference src/utils/list.py:46-95)."""
    out = [default] * length
    if isinstance(level, int):
        out[level] = value
    elif level == 'all':
        out[start:] = [value] * (length - start)
    elif level[-1] == '+':
        i = int(level[:-1])
        out[i:] = [value] * (length - i)
    elif level[-1] == '-':
        i = int(level[:-1])
        out[:i] = [value] * i
    else:
        raise ValueError(f'Unsupported level={level}')
    return out


class NAG:
    def __init__(self, data_list, start_i_level=0):
        assert len(data_list) > 0, "The NAG must have at least 1 level"
        self._list = list(data_list)
        self.start_i_level = start_i_level

    # level bookkeeping (reference nag.py:745-775) ------------------------------
    @property
    def num_levels(self):
        return len(self._list)

    @property
    def end_i_level(self):
        return self.start_i_level + self.num_levels - 1

    @property
    def absolute_num_levels(self):
        return self.num_levels + self.start_i_level

    @property
    def level_range(self):
        return range(self.start_i_level, self.end_i_level + 1)

    @property
    def has_atoms(self):
        return self.start_i_level == 0

    @property
    def device(self):
        return self._list[0].device

    @property
    def num_points(self):
        return [d.num_nodes for d in self._list]

    def __len__(self):
        return self.num_levels

    def __iter__(self):
        for i in self.level_range:
            yield self[i]

    def __getitem__(self, idx):
        if isinstance(idx, int):
            if idx < 0:
                idx = idx % self.num_levels + self.start_i_level
            assert idx in self.level_range, \
                f"Level {idx} is out of range. NAG has levels {self.level_range}"
            return self._list[idx - self.start_i_level]
        levels = list(range(self.absolute_num_levels))[idx]
        assert all(l in self.level_range for l in levels)
        return NAG([self[l] for l in levels], levels[0])

    # transforms / helpers ---------------------------------------------------
    def to(self, device, non_blocking=False):
        out = self.__class__.__new__(self.__class__)
        out._list = [d.to(device, non_blocking=non_blocking) for d in self._list]
        out.start_i_level = self.start_i_level
        return out

    def cuda(self, non_blocking=False):
        return self.to('cuda', non_blocking=non_blocking)

    def cpu(self):
        return self.to('cpu')

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        out._list = [d.clone() for d in self._list]
        out.start_i_level = self.start_i_level
        return out

    def add_keys_to(self, level, keys, to='x', strict=True, delete_after=False):
        """reference nag.py:834-868"""
        per_level = _fill_levels(level, [], keys, self.absolute_num_levels,
                                 self.start_i_level)
        for i_level, ks in enumerate(per_level):
            if ks is None or len(ks) == 0 or i_level not in self.level_range:
                continue
            self[i_level].add_keys_to(keys=ks, to=to, strict=strict,
                                      delete_after=delete_after)

    def select(self, i_level, idx):
        """New NAG holding the nodes `idx` (duplicate-free) of level `i_level`, everything below
        them and every ancestor that keeps a child, with all cross-level indices consistent
        (reference src/data/nag.py:306-399).  Runs on the device (csrc/select.cu)."""
        assert isinstance(i_level, int)
        assert i_level in self.level_range, \
            f"Level {i_level} is out of range. NAG has levels {self.level_range}"
        idx = tensor_idx(idx, device=self.device)
        if idx is None or is_arange(idx, self[i_level].num_nodes):
            return self.clone()
        for i in self.level_range:
            for k in ('obj', 'obj_pred'):
                if k in self[i]:
                    raise NotImplementedError(
                        f"NAG.select: instance labels ('{k}') are outside this package's scope")

        def num_parents(i):
            return self[i + 1].num_nodes if i + 1 <= self.end_i_level else None

        data_list = [None] * self.absolute_num_levels
        data_list[i_level], out_sub, out_super = self[i_level].select(
            idx, update_sub=True, update_super=True, _num_super=num_parents(i_level))

        # lower levels: points selected by the level above, `super_index` handed down
        for i in range(i_level - 1, self.start_i_level - 1, -1):
            idx_sub, sub_super = out_sub
            data_list[i], out_sub, _ = self[i].select(
                idx_sub, update_sub=True, update_super=False, _skip_super=True)
            # (no selection on this level -> nothing handed down: the level keeps its own
            # `super_index`; the reference overwrites it with None there, nag.py:368-370)
            if sub_super is not None:
                data_list[i].super_index = sub_super

        # higher levels: surviving parents, `sub` handed up
        for i in range(i_level + 1, self.absolute_num_levels):
            idx_super, super_sub = out_super
            data_list[i], _, out_super = self[i].select(
                idx_super, update_sub=False, update_super=True, _num_super=num_parents(i),
                _skip_sub=True)
            if super_sub is not None:
                data_list[i].sub = super_sub

        return NAG(data_list[self.start_i_level:], start_i_level=self.start_i_level)

    def get_sub_size(self, high, low=0):
        """Number of `low`-level elements under each `high`-level node, bottom-up
        (reference nag.py:59-110).  Exact int64; on CUDA runs spt_segment_sum_i64."""
        assert self.start_i_level - 1 <= low < high < self.absolute_num_levels
        from .. import ops

        def seg_sum(values, index, num):
            if index.is_cuda:
                return ops.node_size(index, num, child_size=values)
            out = torch.zeros(num, dtype=torch.long, device=index.device)
            src = values if values is not None else torch.ones_like(index)
            return out.index_add_(0, index, src)

        start = self.start_i_level
        if low >= start and self[low].node_size is not None:
            sizes = seg_sum(self[low].node_size, self[low].super_index,
                            self[low + 1].num_nodes)
            cur = low + 1
        elif self[low + 1].sub is not None:
            sizes = self[low + 1].sub.sizes.long()
            cur = low + 1
        elif low >= start:
            sizes = seg_sum(None, self[low].super_index, self[low + 1].num_nodes)
            cur = low + 1
        elif self[start].node_size is not None:
            # nano NAGs (level `low` not loaded) whose first-level sizes were
            # computed at preprocessing time
            sizes = self[start].node_size
            cur = start
        else:
            raise ValueError(f"Cannot infer the size of level {low=} element sizes")
        for i in range(cur, high):
            sizes = seg_sum(sizes, self[i].super_index, self[i + 1].num_nodes)
        return sizes

    def get_super_index(self, high, low=0):
        """reference nag.py:112-138"""
        assert self.start_i_level - 1 <= low < high <= self.end_i_level + 1
        idx = self[0].sub.to_super_index() if low < 0 else self[low].super_index
        for i in range(max(low, -1) + 1, high):
            idx = self[i].super_index[idx]
        return idx

    @classmethod
    def load(cls, path, low=0, high=-1, idx=None, keys_low=None, keys=None,
             non_fp_to_long=False, rgb_to_float=False, **kwargs):
        """Read levels `low`..`high` of a file written by the reference's `NAG.save` (reference
        src/data/nag.py:436-570); see io/nag_io.py.  CPU tensors; integers stay in their
        stored (smallest) dtype unless `non_fp_to_long`."""
        from ..io import load_nag
        return load_nag(path, low=low, high=high, idx=idx, keys_low=keys_low, keys=keys,
                        non_fp_to_long=non_fp_to_long, rgb_to_float=rgb_to_float)

    def save(self, path, y_to_csr=True, pos_dtype=torch.float, fp_dtype=torch.float,
             rgb_to_byte=True):
        """Write the file format of the reference's `NAG.save` (reference
        src/data/nag.py:401-434; uncompressed HDF5, see io/h5write.py)."""
        from ..io import save_nag
        save_nag(self, path, y_to_csr=y_to_csr, pos_dtype=pos_dtype, fp_dtype=fp_dtype,
                 rgb_to_byte=rgb_to_byte)

    def get_sampling(self, high=1, low=0, n_max=32, n_min=1, mask=None,
                     return_pointers=False, seed=None):
        """Indices sampling `low`-level elements by the `high`-level segment they belong to:
        at least `n_min`, at most `n_max` per segment, without replacement (reference
        src/data/nag.py:662-711 -> `sparse_sample`).  On the device (csrc/sample.cu)."""
        from .. import ops
        super_index = self.get_super_index(high, low=low)
        return ops.sparse_sample(super_index, n_max=n_max, n_min=n_min, mask=mask,
                                 return_pointers=return_pointers,
                                 num_segments=self[high].num_nodes, seed=seed)

    def __repr__(self):
        return (f"{self.__class__.__name__}(num_levels={self.num_levels}, "
                f"start_i_level={self.start_i_level}, num_points={self.num_points})")


class NAGBatch(NAG):
    """Batch of NAGs: level-wise disjoint union (reference nag.py:870-898)."""

    @classmethod
    def from_nag_list(cls, nag_list):
        start = nag_list[0].start_i_level
        n_levels = nag_list[0].num_levels
        levels = []
        for rel in range(n_levels):
            datas = [n._list[rel] for n in nag_list]
            num_super = None
            if rel + 1 < n_levels:
                num_super = [n._list[rel + 1].num_nodes for n in nag_list]
            levels.append(Batch.from_data_list(datas, num_super_list=num_super))
        out = cls(levels, start)
        return out
